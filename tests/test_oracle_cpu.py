"""CPU tests of the oracle (no GPU): the reference's own AVX2-vs-scalar twins, known answers of the
intrinsics the parity depends on, the float64 model, IO round trips and index-builder invariants."""
import numpy as np
import pytest

from tests import model_f64 as m64

import ctypes as C


def _p(a, t):
    return a.ctypes.data_as(t)


f32p, u32p, u64p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)


def test_consts(oracle_lib):
    # src/consts.rs:10  SCALAR = 1/15 as f32
    assert np.float32(oracle_lib.lib().orc_scalar_const()) == np.float32(1.0) / np.float32(15.0)
    assert abs(float(oracle_lib.lib().orc_scalar_const()) - 0.0666666701) < 1e-9


@pytest.mark.parametrize("x", [0.0, -0.0, 1.0, -1.0, 3.5e38, -3.5e38, 1e-40, -1e-40, float("inf"), float("-inf")])
def test_ord32_monotone_roundtrip(oracle_lib, x):
    L = oracle_lib.lib()
    k = L.orc_ord32_from_f32(x)
    assert np.float32(L.orc_ord32_to_f32(k)).view(np.uint32) == np.float32(x).view(np.uint32)


def test_ord32_order(oracle_lib):
    L = oracle_lib.lib()
    xs = np.array([-np.inf, -3.0, -1e-30, -0.0, 0.0, 1e-30, 2.5, np.inf], np.float32)
    ks = [L.orc_ord32_from_f32(float(x)) for x in xs]
    assert ks == sorted(ks) and len(set(ks)) == len(ks)


def test_cvtps_round_half_even_known_answers(oracle_lib):
    """_mm256_cvtps_epi32 gives 0.5->0, 1.5->2, 2.5->2, 14.5->14, 7.5->8; the bias is unused."""
    L = oracle_lib.lib()
    v = np.array([0.5, 1.5, 2.5, 14.5, 7.5, 3.49, 3.51, 15.0], np.float32)
    q = np.zeros(8, np.uint8)
    s = L.orc_scalar_quantize(_p(q, u8p), _p(v, f32p), 8, 0.0, 1.0)
    assert q.tolist() == [0, 2, 2, 14, 8, 3, 4, 15]
    assert s == sum(q.tolist())


def test_quantize_degenerate_delta_zero(oracle_lib):
    """constant residual -> delta = 0 -> inv = inf -> 0*inf = NaN -> cvtps = INT_MIN ->
    low byte 0 and the i32 sums wrap to 0 because D is a multiple of 64."""
    L = oracle_lib.lib()
    D = 128
    v = np.full(D, 2.5, np.float32)
    q = np.full(D, 77, np.uint8)
    with np.errstate(all="ignore"):
        s = L.orc_scalar_quantize(_p(q, u8p), _p(v, f32p), D, 2.5, float("inf"))
    assert q.tolist() == [0] * D and s == 0


@pytest.mark.parametrize("dim", [64, 128, 192, 256, 960, 1536])
def test_twin_binarize_and_popcount_bit_exact(oracle_lib, dim):
    """Integer twins must agree bit-exactly: src/simd.rs:83-107 vs src/utils.rs:90-97, src/simd.rs:326-384 vs
    src/utils.rs:101-107 (AVX2 LUT path only when dim >= 256)."""
    L = oracle_lib.lib()
    rng = np.random.default_rng(dim)
    W = dim // 64
    for _ in range(20):
        q = rng.integers(0, 16, dim).astype(np.uint8)
        a = np.zeros(4 * W, np.uint64)
        b = np.zeros(4 * W, np.uint64)
        L.orc_vector_binarize_query(_p(q, u8p), dim, _p(a, u64p))
        L.orc_vector_binarize_query_raw(_p(q, u8p), dim, _p(b, u64p))
        assert np.array_equal(a, b)
        assert np.array_equal(a, m64.planes_from_q(q, dim))
        x = rng.integers(0, 2**63, W).astype(np.uint64) | (rng.integers(0, 2, W).astype(np.uint64) << np.uint64(63))
        assert L.orc_binary_dot_product(_p(x, u64p), _p(a, u64p), W) == L.orc_binary_dot_product_raw(_p(x, u64p), _p(a, u64p), W)
        ab = L.orc_asymmetric_binary_dot_product(_p(x, u64p), _p(a, u64p), W)
        assert ab == m64.abdp(m64.unpack_bits(x, dim), q.astype(np.int64))


@pytest.mark.parametrize("n", [64, 128, 960, 1536, 100, 7])
def test_twin_minmax_exact_and_l2_dot_close(oracle_lib, n):
    L = oracle_lib.lib()
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    r1 = np.zeros(n, np.float32); r2 = np.zeros(n, np.float32)
    mn1, mx1, mn2, mx2 = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    L.orc_min_max_residual(_p(r1, f32p), _p(x, f32p), _p(y, f32p), n, C.byref(mn1), C.byref(mx1))
    L.orc_min_max_raw(_p(r2, f32p), _p(x, f32p), _p(y, f32p), n, C.byref(mn2), C.byref(mx2))
    assert np.array_equal(r1, r2) and mn1.value == mn2.value and mx1.value == mx2.value
    assert np.array_equal(r1, x - y)
    l2 = L.orc_l2_squared_distance(_p(x, f32p), _p(y, f32p), n)
    dp = L.orc_vector_dot_product(_p(x, f32p), _p(y, f32p), n)
    assert abs(l2 - float(((x.astype(np.float64) - y) ** 2).sum())) <= 1e-5 * max(1.0, l2)
    assert abs(dp - float((x.astype(np.float64) * y).sum())) <= 1e-4 * max(1.0, abs(dp))


def test_l2_lane_order_is_the_avx_one(oracle_lib):
    """The 8-lane association: lane v sums elements v, v+8, ... with fma, then ((s0+s4)+(s1+s5))+((s2+s6)+(s3+s7))."""
    L = oracle_lib.lib()
    rng = np.random.default_rng(5)
    n = 128
    x = (rng.standard_normal(n) * 100).astype(np.float32)
    y = (rng.standard_normal(n) * 100).astype(np.float32)
    import math

    s = [np.float32(0)] * 8
    for i in range(n):
        d = np.float32(x[i] - y[i])
        # fma with a single rounding == exact product in float64 (24+24 bits fit) then one rounding of the sum
        s[i % 8] = np.float32(np.float64(d) * np.float64(d) + np.float64(s[i % 8]))
    ref = np.float32(np.float32(np.float32(s[0] + s[4]) + np.float32(s[1] + s[5])) + np.float32(np.float32(s[2] + s[6]) + np.float32(s[3] + s[7])))
    got = np.float32(L.orc_l2_squared_distance(_p(x, f32p), _p(y, f32p), n))
    assert got.view(np.uint32) == ref.view(np.uint32)


def test_heap_replay_is_topk_multiset(oracle_lib):
    L = oracle_lib.lib()
    rng = np.random.default_rng(3)
    for n, k in [(5, 10), (10, 10), (100, 10), (1000, 7), (50, 1)]:
        a = rng.integers(0, 40, n).astype(np.float32)  # many exact ties
        ids = np.arange(n, dtype=np.uint32)
        od = np.zeros(k + 1, np.float32); oi = np.zeros(k + 1, np.uint32)
        c = L.orc_heap_replay(_p(a, f32p), _p(ids, u32p), n, k, _p(od, f32p), _p(oi, u32p))
        assert c == min(n, k)
        assert sorted(od[:c].tolist()) == sorted(a.tolist())[:c]
        for d, i in zip(od[:c], oi[:c]):
            assert a[i] == d


def test_builder_invariants_and_f64_model(case_d128):
    a = case_d128["arrays"]
    D = a["dim"]
    n = a["base"].shape[0]
    off = a["offsets"]
    assert off[0] == 0 and off[-1] == n and np.all(np.diff(off.astype(np.int64)) >= 0)
    assert sorted(a["map_ids"].tolist()) == list(range(n))
    P = a["orthogonal"].astype(np.float64)
    assert np.allclose(P @ P.T, np.eye(D), atol=1e-5)
    # base is the permuted, unrotated input (src/rabitq.rs:245-247)
    assert np.array_equal(a["base"], case_d128["base"][a["map_ids"]])
    # per-vector factors from the definitions (src/rabitq.rs:218-229), cluster order ascending by centroid distance
    rng = np.random.default_rng(0)
    cl = np.searchsorted(off, np.arange(n), side="right") - 1
    for j in rng.integers(0, n, 40):
        c = cl[j]
        xr = a["base"][j].astype(np.float64) @ P
        r = xr - a["centroids"][c].astype(np.float64)
        bits = m64.unpack_bits(a["codes"][j], D)
        assert np.array_equal(bits, (r > 0).astype(np.int64)) or np.min(np.abs(r)) < 1e-4
        s = 2.0 * bits - 1.0
        nr = np.linalg.norm(r)
        xdp = (r * s).sum() / (nr * np.sqrt(D))
        t = nr / xdp
        ip, ppc, err, cds = a["factors"][j]
        assert np.isclose(cds, nr * nr, rtol=1e-4)
        assert np.isclose(ip, -2.0 / np.sqrt(D) * t, rtol=1e-4)
        assert np.isclose(ppc, ip * s.sum(), rtol=1e-4, atol=1e-3)
        assert np.isclose(err, 2 * 1.9 / np.sqrt(D - 1) * np.sqrt(max(t * t - nr * nr, 0)), rtol=1e-3)
    for c in rng.integers(0, len(off) - 1, 10):
        rows = np.arange(off[c], off[c + 1])
        if len(rows) > 1:
            assert np.all(np.diff(a["factors"][rows, 3]) >= -1e-3 * np.abs(a["factors"][rows[1:], 3]))


@pytest.mark.parametrize("case_name", ["case_d128", "case_d96", "case_d960"])
def test_trace_against_f64_model(request, case_name):
    case = request.getfixturevalue(case_name)
    a = case["arrays"]
    D = a["dim"]
    ix = case["oracle"]
    q = case["queries"][0]
    tr = ix.trace(q, 6, 10)
    qp = np.zeros(D, np.float32); qp[: len(q)] = q
    y = qp.astype(np.float64) @ a["orthogonal"].astype(np.float64)
    assert np.allclose(tr["y"], y, rtol=1e-4, atol=1e-3 * np.abs(y).max())
    cd = ((a["centroids"].astype(np.float64) - y[None, :]) ** 2).sum(1)
    assert np.allclose(tr["centroid_dist"], cd, rtol=1e-4)
    assert np.array_equal(np.sort(tr["probe_dist"]), tr["probe_dist"])
    assert set(tr["probe_ids"].tolist()) == set(np.argsort(cd, kind="stable")[:6].tolist())
    t = 0
    for p, c in enumerate(tr["probe_ids"]):
        r = tr["y"] - a["centroids"][c]
        lo, delta, qv = m64.quantize_rne(r)
        assert lo == tr["lo"][p] and delta == tr["delta"][p]
        assert np.array_equal(qv.astype(np.uint8), tr["quantized"][p])
        assert int(qv.sum()) == int(tr["sum"][p])
        assert np.array_equal(m64.planes_from_q(qv, D), tr["planes"][p])
        for j in range(a["offsets"][c], a["offsets"][c + 1]):
            if (j - a["offsets"][c]) % 17 == 0:
                bits = m64.unpack_bits(a["codes"][j], D)
                ab = m64.abdp(bits, qv)
                assert ab == tr["abdp"][t]
                r1 = m64.rough_f64(a["factors"][j], tr["probe_dist"][p], lo, delta, qv.sum(), ab)
                r2 = m64.rough_identity_f64(a["factors"][j], tr["probe_dist"][p], lo, delta, qv, bits)
                scale = abs(float(a["factors"][j][3])) + abs(float(tr["probe_dist"][p])) + 1.0
                assert abs(r1 - float(tr["rough"][t])) <= 2e-5 * scale
                assert abs(r2 - r1) <= 1e-6 * scale
            assert tr["pair_pos"][t] == j
            t += 1
    assert t == tr["pairs"]
    # rerank bookkeeping: exact distances are true squared L2 to the unrotated base vectors
    m = tr["action"] > 0
    ex = ((a["base"][tr["pair_pos"][m]].astype(np.float64) - qp[None, :]) ** 2).sum(1)
    assert np.allclose(tr["exact"][m], ex, rtol=1e-5)
    assert tr["precise"] == int(m.sum())
    res_ids = sorted(i for _, i in tr["result"])
    pushed = tr["action"] == 2
    best = np.argsort(tr["exact"][pushed], kind="stable")[:10]
    assert sorted(np.sort(tr["exact"][pushed])[:10].tolist()) == sorted(d for d, _ in tr["result"])
    assert len(res_ids) == min(10, int(pushed.sum()))


def test_recall_vs_brute_force(case_d128):
    from tools import synth

    truth = synth.brute_force_topk_numpy(case_d128["base"], case_d128["queries"], 10)
    r = case_d128["oracle"].query_batch(case_d128["queries"], 64, 10)  # all 64 clusters probed
    rec = np.mean([len(set(r["ids"][i].tolist()) & set(truth[i].tolist())) / 10 for i in range(truth.shape[0])])
    assert rec >= 0.97
    assert r["rough"] == case_d128["base"].shape[0] * case_d128["queries"].shape[0]


def test_dump_load_round_trip_and_layout(case_d96, tmp_path):
    """Six-file layout (src/rabitq.rs:128-156): record headers, centroids stored dim x k."""
    from oracle import oracle as orc

    ix = case_d96["oracle"]
    d = tmp_path / "idx"
    ix.dump_to_dir(str(d))
    a = ix.arrays()
    D, n, k = a["dim"], a["base"].shape[0], a["centroids"].shape[0]
    import os, struct

    assert os.path.getsize(d / "base.fvecs") == n * (4 + 4 * D)
    assert os.path.getsize(d / "orthogonal.fvecs") == D * (4 + 4 * D)
    assert os.path.getsize(d / "centroids.fvecs") == D * (4 + 4 * k)
    assert os.path.getsize(d / "offsets_ids.ivecs") == 4 + 4 * (k + 1) + 4 + 4 * n
    assert os.path.getsize(d / "factors.fvecs") == 4 + 16 * n
    assert os.path.getsize(d / "x_binary_vec.u64vecs") == 4 + 8 * n * (D // 64)
    with open(d / "centroids.fvecs", "rb") as f:
        assert struct.unpack("<I", f.read(4))[0] == k
    ix2 = orc.OracleIndex.load_from_dir(str(d))
    b = ix2.arrays()
    for key in a:
        assert np.array_equal(a[key], b[key]), key
    q = case_d96["queries"][:5]
    r1 = ix.query_batch(q, 8, 10); r2 = ix2.query_batch(q, 8, 10)
    assert np.array_equal(r1["ids"], r2["ids"]) and np.array_equal(r1["dist"], r2["dist"])


def test_query_dim_assert(case_d128):
    with pytest.raises(RuntimeError):  # src/rabitq.rs:275
        case_d128["oracle"].query(np.zeros(64, np.float32), 4, 10)


def test_heuristic_reranker_runs(case_d128):
    q = case_d128["queries"][:8]
    r = case_d128["oracle"].query_batch(q, 16, 10, heuristic_rank=True)
    assert np.all(r["count"] == 10)


def test_raw_bias_switch_changes_only_the_quantiser(oracle_lib):
    """orc_set_raw_bias: the query path takes scalar_quantize's non-AVX2 branch (src/utils.rs:194-209); with bias = 0.5 the
    truncation rounds half UP, so it differs from the AVX2 branch exactly where (r - lo) / delta sits on a .5 tie or where
    round-half-even and truncate(+0.5) disagree -- and nowhere else."""
    from tools import synth

    base, queries, cent = synth.make_numpy(3000, 64, 8, 12, "sift", 77)
    ix = oracle_lib.OracleIndex.from_arrays(base, cent, seed=3, nthreads=2)
    a = ix.trace(queries[0], 4, 5)
    ix.set_raw_bias(np.full(ix.dim, 0.5, np.float32))
    b = ix.trace(queries[0], 4, 5)
    ix.set_raw_bias(None)
    c = ix.trace(queries[0], 4, 5)
    assert np.array_equal(a["planes"], c["planes"]) and np.array_equal(a["sum"], c["sum"])
    assert np.array_equal(a["lo"], b["lo"]) and np.array_equal(a["delta"], b["delta"])
    qa, qb = a["quantized"].astype(np.int32), b["quantized"].astype(np.int32)
    assert np.all(np.abs(qa - qb) <= 1)   # half-even vs half-up: at most one step apart, and only on ties
    # the raw twin itself, element by element
    v = np.array([0.0, 0.49, 0.5, 1.5, 2.5, 14.99, 15.0, 300.0, -3.0, np.nan], np.float32)
    out = np.zeros(len(v), np.uint8)
    s = oracle_lib.lib().orc_scalar_quantize_raw(out.ctypes.data_as(oracle_lib.c_u8p), v.ctypes.data_as(oracle_lib.c_f32p),
                                                 np.full(len(v), 0.5, np.float32).ctypes.data_as(oracle_lib.c_f32p), len(v), 0.0, 1.0)
    assert out.tolist() == [0, 0, 1, 2, 3, 15, 15, 255, 0, 0] and s == sum(out.tolist())


# ---- the sequential rerankers, restated a second time in plain Python ---------------------------------------------------------
def _py_rerank(rough, pos, exact_of, map_ids, topk, heuristic):
    """src/rerank.rs:81-113 (HeapReRanker) and :142-176 (HeuristicReRanker), statement by statement, on Python floats holding
    fp32 values.  Independent of the oracle's C++ heap / window code: only the visit order and the exact distances are shared."""
    import heapq

    FLT_MAX, FLT_MIN = float(np.finfo(np.float32).max), float(np.finfo(np.float32).min)
    thr, precise = FLT_MAX, 0
    if not heuristic:
        heap = []  # BinaryHeap<(Ord32, AlwaysEqual<u32>)> = max-heap on the distance: heapq on the negated key
        for r, u in zip(rough, pos):
            if r < thr:                                     # :84
                acc = exact_of(u)
                precise += 1
                if acc < thr:                               # :92
                    heapq.heappush(heap, (-acc, int(map_ids[u])))
                    if len(heap) > topk:                    # :95
                        heapq.heappop(heap)
                    if len(heap) == topk:                   # :98
                        thr = -heap[0][0]
        return sorted(-d for d, _ in heap), precise
    arr, count, recent = [], 0, FLT_MIN
    for r, u in zip(rough, pos):
        if r < thr:                                         # :145
            acc = exact_of(u)
            precise += 1
            if acc < thr:                                   # :153
                arr.append((acc, int(map_ids[u])))
                count += 1
                recent = max(recent, acc)
                if count >= 12:                             # WINDOW_SIZE, src/consts.rs:12
                    thr, count, recent = recent, 0, FLT_MIN
    return sorted(d for d, _ in arr)[: min(topk, len(arr))], precise       # get_result: the topk smallest


@pytest.mark.parametrize("case_name", ["case_d128", "case_d96", "case_d960"])
@pytest.mark.parametrize("heuristic", [False, True])
def test_rerankers_against_python_restatement(request, case_name, heuristic):
    from oracle import oracle as orc

    case = request.getfixturevalue(case_name)
    a, o = case["arrays"], case["oracle"]
    D = a["dim"]
    L = orc.lib()
    for probe, topk in [(4, 10), (24, 10), (12, 1), (16, 40)]:
        for qi in range(0, case["queries"].shape[0], 5):
            q = case["queries"][qi]
            tr = o.trace(q, probe, topk, heuristic_rank=heuristic)
            qpad = np.zeros(D, np.float32)
            qpad[: q.shape[0]] = q

            def exact_of(u):
                row = np.ascontiguousarray(a["base"][u], dtype=np.float32)
                return float(np.float32(L.orc_l2_squared_distance(_p(row, f32p), _p(qpad, f32p), D)))

            dists, precise = _py_rerank([float(x) for x in tr["rough"]], [int(x) for x in tr["pair_pos"]], exact_of, a["map_ids"], topk, heuristic)
            assert precise == tr["precise"], (case_name, probe, topk, qi)
            got = sorted(d for d, _ in tr["result"])
            assert np.array_equal(np.array(dists, np.float32).view(np.uint32), np.array(got, np.float32).view(np.uint32)), (case_name, probe, topk, qi)


# ---- bit-exact numpy emulations of the fp32 evaluation ORDER (not just the value) ---------------------------------------------
def _fma32(a, b, c):
    """fp32 fused multiply-add: the product of two fp32 values is exact in fp64 (24 + 24 bits), one rounding of the sum."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _reduce8(acc):
    """reduce_f32_256 (src/simd.rs:292-303): ((s0+s4)+(s1+s5)) + ((s2+s6)+(s3+s7)), every sum rounded to fp32; acc: [8, ...]."""
    c = (acc[:4] + acc[4:]).astype(np.float32)
    return ((c[0] + c[1]).astype(np.float32) + (c[2] + c[3]).astype(np.float32)).astype(np.float32)


@pytest.mark.parametrize("case_name", ["case_d128", "case_d96"])
def test_rotation_and_estimator_order_bit_exact(request, case_name):
    """`project` = `vector_dot_product(vec, P.col(i))` (src/utils.rs:237-258, src/simd.rs:257-314): lane l accumulates
    fma(q[r], P[r, i], acc_l) over r = l, l+8, ... ascending, then reduce_f32_256; and `calculate_rough_distance`
    (src/rabitq.rs:352-363) left to right in fp32.  Emulated in numpy with explicit roundings; must equal the oracle's bits."""
    case = request.getfixturevalue(case_name)
    a, o = case["arrays"], case["oracle"]
    D = a["dim"]
    P = a["orthogonal"]
    for qi in (0, 3):
        q = case["queries"][qi]
        tr = o.trace(q, 5, 10)
        qp = np.zeros(D, np.float32)
        qp[: len(q)] = q
        acc = np.zeros((8, D), np.float32)
        for r in range(D):  # sequential over rows, vectorised over the D output columns
            acc[r % 8] = _fma32(np.full(D, qp[r], np.float32), P[r, :], acc[r % 8])
        y = _reduce8(acc)
        assert np.array_equal(y.view(np.uint32), tr["y"].view(np.uint32))
        # centroid distances: l2_squared_distance(centroid, y) (src/rabitq.rs:283-293, src/simd.rs:14-73): diff = c - y, fma(diff, diff, acc)
        cent = a["centroids"]
        acc = np.zeros((8, cent.shape[0]), np.float32)
        for d in range(D):
            diff = (cent[:, d] - y[d]).astype(np.float32)
            acc[d % 8] = _fma32(diff, diff, acc[d % 8])
        cd = _reduce8(acc)
        assert np.array_equal(cd.view(np.uint32), tr["centroid_dist"].view(np.uint32))
        # the estimator, term by term in fp32
        t = 0
        f32 = np.float32
        for p, c in enumerate(tr["probe_ids"]):
            ycd, lo, delta, ssum = f32(tr["probe_dist"][p]), f32(tr["lo"][p]), f32(tr["delta"][p]), f32(tr["sum"][p])
            sq = np.sqrt(ycd, dtype=np.float32)
            for j in range(a["offsets"][c], a["offsets"][c + 1]):
                ip, ppc, err, cds = (f32(x) for x in a["factors"][j])
                ab = f32(tr["abdp"][t])
                e = f32(f32(cds + ycd) + f32(lo * ppc))
                e = f32(e + f32(f32(f32(f32(f32(2.0) * ab) - ssum) * ip) * delta))
                e = f32(e - f32(err * sq))
                assert e.view(np.uint32) == f32(tr["rough"][t]).view(np.uint32), (qi, p, j)
                t += 1
