"""-m gpu: the edge cases the reference's arithmetic has, driven through the CUDA path and
compared with the CPU oracle on the same hand-made index:

* delta == 0 (constant residual -> 1/0 = inf -> 0 * inf = NaN -> `_mm256_cvtps_epi32` gives INT_MIN, src/simd.rs:214-215);
* NaN / +-inf `error_bound` (src/rabitq.rs:225-226) through the strict filter `rough < thr` (src/rerank.rs:84);
* exact duplicates of a base vector at the heap boundary (strict `<`, src/rerank.rs:92-100);
* clusters of size 0 and 1, cluster sizes that are not multiples of 32 / 128, `count < topk`;
* a hypothesis property test over (dim, clusters, cluster-size skew, probe, topk).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(base, cent, P=None, seed=3, mutate=None):
    """(oracle index, GPU index) over the same arrays; `mutate(arrays)` may edit the built arrays before both adopt them."""
    import rabitq_b200 as rb
    from oracle import oracle as orc

    o = orc.OracleIndex.from_arrays(base, cent, P=P, seed=seed, nthreads=4)
    a = o.arrays()
    if mutate is not None:
        mutate(a)
        o.close()
        o = orc.OracleIndex.from_built(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"], a["factors"])
    g = rb.RaBitQ.from_arrays(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"], a["factors"], device=0)
    return o, g, a


def _bits_equal_nan_aware(x, y):
    x, y = np.asarray(x, np.float32), np.asarray(y, np.float32)
    nx, ny = np.isnan(x), np.isnan(y)
    return np.array_equal(nx, ny) and np.array_equal(x[~nx].view(np.uint32), y[~ny].view(np.uint32))


def _check_all(o, g, a, q, probe, topk, heuristic=False, stages=True):
    """Stage-level (lo, delta, sum, planes, abdp, rough) and end-to-end (distances, ids up to ties, counters) equality."""
    q = np.ascontiguousarray(q, np.float32)
    if stages:
        lo, delta, s, planes = g.stage_quantize(q, probe)
        rough, abdp, start = g.stage_scan(q, probe, pair_capacity=max(1, a["base"].shape[0] * q.shape[0]))
        for i in range(q.shape[0]):
            tr = o.trace(q[i], probe, topk)
            assert _bits_equal_nan_aware(lo[i], tr["lo"]), f"lo, query {i}"
            assert _bits_equal_nan_aware(delta[i], tr["delta"]), f"delta, query {i}"
            assert np.array_equal(s[i], tr["sum"]), f"sum q_u, query {i}"
            assert np.array_equal(planes[i], tr["planes"]), f"planes, query {i}"
            b, e = int(start[i]), int(start[i + 1])
            assert e - b == tr["pairs"]
            assert np.array_equal(abdp[b:e], tr["abdp"]), f"abdp, query {i}"
            assert _bits_equal_nan_aware(rough[b:e], tr["rough"]), f"rough, query {i}"
    g.metrics_reset()
    gd, gi, gc = g.query_batch(q, probe, topk, heuristic_rank=heuristic)
    r = o.query_batch(q, probe, topk, heuristic_rank=heuristic)
    assert np.array_equal(gc, r["count"]), (gc, r["count"])
    for i in range(q.shape[0]):
        c = int(gc[i])
        assert np.array_equal(np.sort(gd[i, :c]).view(np.uint32), np.sort(r["dist"][i, :c]).view(np.uint32)), f"distances, query {i}"
        sa, sb = set(gi[i, :c].tolist()), set(r["ids"][i, :c].tolist())
        if sa != sb:  # only exact-distance ties may differ
            da = sorted(float(gd[i, j]) for j in range(c) if int(gi[i, j]) not in sb)
            db = sorted(float(r["dist"][i, j]) for j in range(c) if int(r["ids"][i, j]) not in sa)
            assert da == db, f"ids differ beyond ties, query {i}"
    m = g.metrics()
    assert m["rough"] == r["rough"] and m["precise"] == r["precise"], (m, r["rough"], r["precise"])
    return gd, gi, gc


def _clustered(rng, dim, sizes, spread=0.05):
    """Gaussian blobs with EXACTLY the given cluster populations around well-separated centroids."""
    k = len(sizes)
    cent = rng.normal(size=(k, dim)).astype(np.float32) * 4.0
    rows = [cent[c] + spread * rng.normal(size=(n, dim)).astype(np.float32) for c, n in enumerate(sizes) if n > 0]
    base = np.concatenate(rows).astype(np.float32) if rows else np.zeros((0, dim), np.float32)
    return base, cent


def test_delta_zero_constant_residual():
    """y - c constant over the dimensions: lo == hi, delta = 0, inv = inf, (r - lo) * inv = NaN -> code INT_MIN -> byte 0."""
    rng = np.random.default_rng(5)
    dim = 128
    base, cent = _clustered(rng, dim, [300, 257, 64, 190, 33, 1, 127, 129])
    cent = (np.round(cent * 64.0) / 64.0).astype(np.float32)   # multiples of 2^-6: c + 0.5 and (c + 0.5) - c are exact in fp32
    o, g, a = _pair(base, cent, P=np.eye(dim, dtype=np.float32))
    q = np.stack([cent[0] + np.float32(0.5),        # constant residual against its nearest centroid
                  cent[3].copy(),                    # exactly a centroid: residual identically 0
                  cent[1] + np.float32(-0.25),
                  base[5]]).astype(np.float32)
    lo, delta, s, planes = g.stage_quantize(q, 3)
    assert delta[0, 0] == 0.0 and delta[1, 0] == 0.0 and delta[2, 0] == 0.0   # the nearest probe of the crafted queries
    assert not planes[0, 0].any()                                                # INT_MIN & 15 == 0 in every dimension
    _check_all(o, g, a, q, probe=3, topk=10)
    _check_all(o, g, a, q, probe=8, topk=5, stages=False)


@pytest.mark.parametrize("heuristic", [False, True])
def test_nan_and_infinite_error_bounds(heuristic):
    """NaN error_bound -> NaN rough -> `rough < thr` is false on both sides; +inf -> rough = -inf (always passes the filter);
    -inf -> rough = +inf (never passes once the threshold is finite)."""
    rng = np.random.default_rng(6)
    dim = 192
    base, cent = _clustered(rng, dim, [140, 260, 97, 31, 64, 200], spread=0.3)

    def mutate(a):
        f = a["factors"]
        n = f.shape[0]
        idx = rng.permutation(n)
        f[idx[: n // 10], 2] = np.nan
        f[idx[n // 10: n // 10 + n // 20], 2] = np.inf
        f[idx[n // 10 + n // 20: n // 10 + n // 10], 2] = -np.inf
        f[idx[-7:], 3] = np.nan   # center_distance_square NaN as well

    o, g, a = _pair(base, cent, mutate=mutate)
    q = (base[rng.integers(0, base.shape[0], 12)] + 0.1 * rng.normal(size=(12, dim))).astype(np.float32)
    _check_all(o, g, a, q, probe=4, topk=10, heuristic=heuristic, stages=not heuristic)


def test_duplicate_vectors_at_the_heap_boundary():
    """40 exact copies of one vector next to the query: exact distances tie across the k-th place; strict `<` keeps the earlier one."""
    rng = np.random.default_rng(7)
    dim = 64
    base, cent = _clustered(rng, dim, [200, 150, 90], spread=0.2)
    dup = base[17].copy()
    base[40:80] = dup
    base[250:260] = dup          # copies that land in another cluster's neighbourhood still quantise against their own centroid
    o, g, a = _pair(base, cent)
    q = np.stack([dup, dup + np.float32(1e-3), base[3], cent[1]]).astype(np.float32)
    for topk in (1, 10, 45, 64):
        _check_all(o, g, a, q, probe=3, topk=topk, stages=(topk == 10))


def test_empty_and_single_vector_clusters_and_short_results():
    """Clusters of size 0 and 1 are probed first; cluster sizes around the 32- and 128-vector block boundaries; count < topk."""
    rng = np.random.default_rng(8)
    dim = 128
    sizes = [0, 1, 3, 31, 32, 33, 127, 128, 129, 0, 255, 2]
    base, cent = _clustered(rng, dim, sizes)
    o, g, a = _pair(base, cent)
    off = a["offsets"]
    assert sorted(np.diff(off).tolist()) == sorted(sizes)
    q = np.concatenate([cent + np.float32(0.01), cent[[0, 9]] + np.float32(0.3)]).astype(np.float32)
    gd, gi, gc = _check_all(o, g, a, q, probe=1, topk=10)
    assert int(gc[0]) == 0 and int(gc[1]) == 1 and int(gc[2]) == 3          # count < topk (and 0) where the probed cluster is small
    _check_all(o, g, a, q, probe=2, topk=40, stages=False)
    _check_all(o, g, a, q, probe=len(sizes), topk=1000, stages=False)       # topk > n: every vector comes back
    _check_all(o, g, a, q, probe=5, topk=7, heuristic=True, stages=False)


def test_large_probe_and_topk_caps_lifted():
    """probe.min(k) (src/rabitq.rs:294) and BinaryHeap::with_capacity(topk) (src/rerank.rs:69-78) have no caps in the reference."""
    rng = np.random.default_rng(9)
    dim = 64
    sizes = [int(x) for x in rng.integers(0, 9, size=5000)]
    base, cent = _clustered(rng, dim, sizes, spread=0.5)
    o, g, a = _pair(base, cent)
    q = (base[rng.integers(0, base.shape[0], 6)] + 0.2 * rng.normal(size=(6, dim))).astype(np.float32)
    _check_all(o, g, a, q, probe=5000, topk=10, stages=False)     # probe > 4096
    _check_all(o, g, a, q, probe=4500, topk=1500, stages=False)   # topk > 1024


def test_property_random_shapes():
    """hypothesis over (dim, clusters, cluster-size skew, probe, topk, heuristic): CUDA == oracle, every time."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=16, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(dim=st.sampled_from([40, 64, 100, 128, 192, 320, 704]), k=st.integers(1, 40), skew=st.floats(0.0, 2.5),
           n=st.integers(1, 3000), probe=st.integers(1, 48), topk=st.sampled_from([1, 3, 10, 33, 100]), heur=st.booleans(),
           seed=st.integers(0, 2 ** 16))
    def run(dim, k, skew, n, probe, topk, heur, seed):
        rng = np.random.default_rng(seed)
        w = np.exp(skew * rng.normal(size=k))
        sizes = rng.multinomial(n, w / w.sum()).tolist()
        base, cent = _clustered(rng, dim, sizes, spread=0.4)
        if base.shape[0] == 0:
            return
        o, g, a = _pair(base, cent, seed=seed % 97 + 1)
        nq = 5
        q = (base[rng.integers(0, base.shape[0], nq)] + 0.3 * rng.normal(size=(nq, dim))).astype(np.float32)
        _check_all(o, g, a, q, probe=probe, topk=topk, heuristic=heur, stages=not heur)
        g.close()
        o.close()

    run()


def test_speculative_slot_sizing_and_its_overflow_path():
    """The survivor slots of a batch are sized from earlier batches (no host round trip in the middle of the batch); a batch that
    needs more is neutralised on the device and repeated with the exact sizes.  Results never depend on which path ran."""
    from tools import synth

    base, queries, cent = synth.make_numpy(6000, 128, 48, 24, "sift", seed=21)
    o, g, a = _pair(base, cent, seed=5)
    ref = o.query_batch(queries, 6, 10)

    def same(res):
        d, _, c = res
        for i in range(queries.shape[0]):
            n = int(ref["count"][i])
            assert int(c[i]) == n
            assert np.array_equal(np.sort(d[i, :n]).view(np.uint32), np.sort(ref["dist"][i, :n]).view(np.uint32)), i

    same(g.query_batch(queries, 6, 10))           # first batch: exact sizes (nothing to speculate from)
    same(g.query_batch(queries, 6, 10))           # second batch: speculative, fits
    g.set_option("spec_words_per_query_milli", 1)  # pretend earlier batches were tiny: the next one overflows and is repeated
    g.metrics_reset()
    same(g.query_batch(queries, 6, 10))
    assert g.metrics()["precise"] == ref["precise"] and g.metrics()["rough"] == ref["rough"]
    same(g.query_batch(queries, 6, 10))           # high-water mark now real again: speculative, fits
    g.set_option("speculative_sizing", 0)
    same(g.query_batch(queries, 6, 10))
    g.close()
    o.close()
