"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same seeded index and queries.

Bars (BASELINE.json north_star): quantised codes / bit-planes / popcount sums bit-exact; estimated distances
within 1e-5 relative (we assert bit-exact); final top-k id lists identical up to exact-distance ties; the
reference's `rough` and `precise` counters equal.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu_index(case):
    import rabitq_b200 as rb

    a = case["arrays"]
    return rb.RaBitQ.from_arrays(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"],
                                 a["factors"], device=0)


CASES = ["case_d128", "case_d96", "case_d960"]


@pytest.fixture(params=CASES)
def pair(request):
    case = request.getfixturevalue(request.param)
    if "gpu" not in case:
        case["gpu"] = _gpu_index(case)
    return case


def test_rotate_bit_exact(pair):
    q = pair["queries"]
    y = pair["gpu"].stage_rotate(q)
    for i in range(q.shape[0]):
        tr = pair["oracle"].trace(q[i], 4, 10)
        assert np.array_equal(y[i].view(np.uint32), tr["y"].view(np.uint32)), f"query {i}"


def test_probe_lists_bit_exact(pair):
    q = pair["queries"]
    probe = 12
    cd, pid, pd = pair["gpu"].stage_probe(q, probe)
    for i in range(q.shape[0]):
        tr = pair["oracle"].trace(q[i], probe, 10)
        assert np.array_equal(cd[i].view(np.uint32), tr["centroid_dist"].view(np.uint32))
        assert np.array_equal(pid[i], tr["probe_ids"])
        assert np.array_equal(pd[i].view(np.uint32), tr["probe_dist"].view(np.uint32))


def test_probe_larger_than_k(pair):
    q = pair["queries"][:4]
    K = pair["gpu"].num_clusters
    _, pid, _ = pair["gpu"].stage_probe(q, K + 50)
    assert pid.shape[1] == K
    for i in range(4):
        assert sorted(pid[i].tolist()) == list(range(K))


def test_quantize_planes_bit_exact(pair):
    q = pair["queries"]
    probe = 12
    lo, delta, s, planes = pair["gpu"].stage_quantize(q, probe)
    for i in range(q.shape[0]):
        tr = pair["oracle"].trace(q[i], probe, 10)
        assert np.array_equal(lo[i].view(np.uint32), tr["lo"].view(np.uint32))
        assert np.array_equal(delta[i].view(np.uint32), tr["delta"].view(np.uint32))
        assert np.array_equal(s[i], tr["sum"])
        assert np.array_equal(planes[i], tr["planes"])


def test_scan_abdp_and_rough_bit_exact(pair):
    q = pair["queries"]
    probe = 12
    n = pair["arrays"]["base"].shape[0]
    rough, abdp, start = pair["gpu"].stage_scan(q, probe, pair_capacity=n * q.shape[0])
    for i in range(q.shape[0]):
        tr = pair["oracle"].trace(q[i], probe, 10)
        s, e = int(start[i]), int(start[i + 1])
        assert e - s == tr["pairs"]
        assert np.array_equal(abdp[s:e], tr["abdp"])
        assert np.array_equal(rough[s:e].view(np.uint32), tr["rough"].view(np.uint32))


def _same_up_to_ties(case, qi, gd, gi, od, oi):
    """North-star bar: "final top-k id lists identical up to exact-distance ties".  The distance multisets must be
    bit-identical; where ids differ, the differing ids must sit on a distance that is tied (inside the list or at
    its boundary), which we verify by recomputing the exact distance of every returned id with the oracle's L2."""
    from oracle import oracle as orc
    import ctypes as C

    go = np.lexsort((gi, gd))
    oo = np.lexsort((oi, od))
    gd, gi, od, oi = gd[go], gi[go], od[oo], oi[oo]
    if not np.array_equal(gd.view(np.uint32), od.view(np.uint32)):
        return False
    if len(set(gi.tolist())) != len(gi):
        return False
    if np.array_equal(gi, oi):
        return True
    D = case["arrays"]["dim"]
    q = np.zeros(D, np.float32)
    q[: case["queries"].shape[1]] = case["queries"][qi]
    f32p = C.POINTER(C.c_float)
    for d, i in zip(gd, gi):
        if i in oi:
            continue
        x = np.zeros(D, np.float32)
        x[: case["base"].shape[1]] = case["base"][i]
        ex = np.float32(orc.lib().orc_l2_squared_distance(x.ctypes.data_as(f32p), q.ctypes.data_as(f32p), D))
        if ex.view(np.uint32) != np.float32(d).view(np.uint32):
            return False
    return True


@pytest.mark.parametrize("probe,topk", [(1, 10), (8, 10), (32, 1), (64, 100)])
def test_query_batch_identical_topk_and_counters(pair, probe, topk):
    q = pair["queries"]
    g = pair["gpu"]
    g.metrics_reset()
    gd, gi, gc = g.query_batch(q, probe, topk)
    o = pair["oracle"].query_batch(q, probe, topk)
    assert np.array_equal(gc, o["count"])
    for i in range(q.shape[0]):
        c = int(gc[i])
        assert _same_up_to_ties(pair, i, gd[i, :c], gi[i, :c], o["dist"][i, :c], o["ids"][i, :c]), f"query {i}"
        assert np.all(np.diff(gd[i, :c]) >= 0)
    m = g.metrics()
    assert m["query"] == q.shape[0]
    assert m["rough"] == o["rough"]
    assert m["precise"] == o["precise"]


@pytest.mark.parametrize("rounds", [[0], [0, 1], [0, 1, 4, 16], [0, 2, 3, 5, 40]])
def test_rounds_do_not_change_results(pair, rounds):
    q = pair["queries"]
    g = pair["gpu"]
    g.set_rounds(rounds)
    try:
        g.metrics_reset()
        gd, gi, gc = g.query_batch(q, 48, 10)
        o = pair["oracle"].query_batch(q, 48, 10)
        for i in range(q.shape[0]):
            c = int(gc[i])
            assert _same_up_to_ties(pair, i, gd[i, :c], gi[i, :c], o["dist"][i, :c], o["ids"][i, :c])
        assert g.metrics()["precise"] == o["precise"]
    finally:
        g.set_rounds([0])


@pytest.mark.parametrize("opt,val", [("first_chunks", 0), ("first_chunks", 3), ("scan_mode", 0), ("scan_mode", 1), ("scan_mode", 2), ("scan_mode", 3),
                                     ("rerank_rows", 1), ("rerank_rows", 3), ("rerank_rows", 32), ("scan_slices", 2), ("scan_slices", 7)])
def test_tuning_knobs_do_not_change_results(pair, opt, val):
    q = pair["queries"]
    g = pair["gpu"]
    default = {"first_chunks": 1, "scan_mode": -1, "rerank_rows": 0, "scan_slices": 1}[opt]
    g.set_option(opt, val)
    try:
        g.metrics_reset()
        gd, gi, gc = g.query_batch(q, 40, 10)
        o = pair["oracle"].query_batch(q, 40, 10)
        for i in range(q.shape[0]):
            c = int(gc[i])
            assert _same_up_to_ties(pair, i, gd[i, :c], gi[i, :c], o["dist"][i, :c], o["ids"][i, :c])
        assert g.metrics()["precise"] == o["precise"] and g.metrics()["rough"] == o["rough"]
    finally:
        g.set_option(opt, default)


def test_single_query_matches_batch(pair):
    q = pair["queries"]
    g = pair["gpu"]
    gd, gi, gc = g.query_batch(q[:5], 16, 10)
    for i in range(5):
        r = g.query(q[i], 16, 10)
        assert [x[1] for x in r] == gi[i, : int(gc[i])].tolist()
        assert np.array_equal(np.array([x[0] for x in r], np.float32).view(np.uint32), gd[i, : int(gc[i])].view(np.uint32))


def test_load_from_dir_round_trip(pair, tmp_path):
    import rabitq_b200 as rb

    d = tmp_path / "idx"
    pair["oracle"].dump_to_dir(str(d))
    g2 = rb.RaBitQ.load_from_dir(str(d), device=0)
    q = pair["queries"][:8]
    a = pair["gpu"].query_batch(q, 16, 10)
    b = g2.query_batch(q, 16, 10)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    g2.close()


def test_errors_match_reference_asserts(pair):
    import rabitq_b200 as rb

    g = pair["gpu"]
    D = g.dim
    with pytest.raises(rb.RabitqError):  # rabitq.rs:275
        g.query(np.zeros(D + 64, np.float32), 4, 10)
    with pytest.raises(rb.RabitqError):
        g.query(np.zeros(D, np.float32), 4, 0)
    with pytest.raises(rb.RabitqError):
        rb.RaBitQ.load_from_dir("/nonexistent/dir")


def test_sharded_union_equals_full(pair):
    """cluster-range shards: per-shard top-k merged on the device equals the unsharded top-k distances."""
    import torch

    import rabitq_b200 as rb

    a = pair["arrays"]
    q = pair["queries"]
    topk, probe, S = 10, 32, 3
    full_d, full_i, _ = pair["gpu"].query_batch(q, probe, topk)
    ds, is_ = [], []
    tot = 0
    for r in range(S):
        g = rb.RaBitQ.from_arrays(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"],
                                  a["factors"], device=0, shard_rank=r, shard_count=S)
        tot += g.num_vectors
        d, i, _ = g.query_batch(q, probe, topk)
        ds.append(d); is_.append(i)
        g.close()
    assert tot == a["base"].shape[0]
    D = torch.from_numpy(np.stack(ds)).cuda()
    I = torch.from_numpy(np.stack(is_).astype(np.int32)).cuda()
    od = torch.empty((q.shape[0], topk), dtype=torch.float32, device="cuda")
    oi = torch.empty((q.shape[0], topk), dtype=torch.int32, device="cuda")
    oc = torch.empty((q.shape[0],), dtype=torch.int32, device="cuda")
    import ctypes as C
    rc = rb.lib().rabitq_merge_topk_device(0, C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), S, q.shape[0], topk,
                                           C.c_void_p(od.data_ptr()), C.c_void_p(oi.data_ptr()), C.c_void_p(oc.data_ptr()),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    md = od.cpu().numpy()
    # Shard-local thresholds are looser than the global one, so the merged list can only be equal or better.
    assert np.all(md <= full_d + 0)
    same = np.mean([np.array_equal(md[i].view(np.uint32), full_d[i].view(np.uint32)) for i in range(q.shape[0])])
    assert same >= 0.9


def test_probe_select_pivot_path_and_ties(oracle_lib):
    """K >= 2048 and K >= 16 * probe takes the sampling-pivot fast path; duplicated centroids give exactly equal
    distances, which must resolve to the smaller centroid id exactly as in the oracle."""
    import rabitq_b200 as rb
    from tools import synth

    base, queries, cent = synth.make_numpy(6000, 64, 24, 2304, "sift", 41)
    cent[1000:1200] = cent[0:200]  # exact duplicates -> exact distance ties
    ix = oracle_lib.OracleIndex.from_arrays(base, cent, seed=7, nthreads=8)
    a = ix.arrays()
    g = rb.RaBitQ.from_arrays(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"], a["factors"], device=0)
    for probe in (1, 7, 64, 144):  # 144 * 16 = 2304 -> still the pivot path; 145+ would fall back to radix
        cd, pid, pd = g.stage_probe(queries, probe)              # classic path (the exact matrix is requested)
        _, pid2, pd2 = g.stage_probe(queries, probe, want_all=False)  # tensor-core prefilter: duplicates = tied keys
        assert np.array_equal(pid, pid2) and np.array_equal(pd.view(np.uint32), pd2.view(np.uint32))
        for i in range(queries.shape[0]):
            tr = ix.trace(queries[i], probe, 5)
            assert np.array_equal(pid[i], tr["probe_ids"]), (probe, i)
            assert np.array_equal(pd[i].view(np.uint32), tr["probe_dist"].view(np.uint32))
    for probe in (200, 2304):  # radix path
        cd, pid, pd = g.stage_probe(queries[:6], probe)
        for i in range(6):
            tr = ix.trace(queries[i], probe, 5)
            assert np.array_equal(pid[i], tr["probe_ids"]), (probe, i)
    gd, gi, gc = g.query_batch(queries, 64, 10)
    o = ix.query_batch(queries, 64, 10)
    assert np.array_equal(np.sort(gd, 1).view(np.uint32), np.sort(o["dist"], 1).view(np.uint32))
    g.close()


@pytest.mark.parametrize("probe,topk", [(4, 10), (32, 10), (64, 50)])
def test_heuristic_reranker_matches_oracle(pair, probe, topk):
    """-h / heuristic_rank=True: HeuristicReRanker (src/rerank.rs:117-176), same sequential replay, windowed threshold."""
    q = pair["queries"]
    g = pair["gpu"]
    g.metrics_reset()
    gd, gi, gc = g.query_batch(q, probe, topk, heuristic_rank=True)
    o = pair["oracle"].query_batch(q, probe, topk, heuristic_rank=True)
    assert np.array_equal(gc, o["count"])
    for i in range(q.shape[0]):
        c = int(gc[i])
        assert _same_up_to_ties(pair, i, gd[i, :c], gi[i, :c], o["dist"][i, :c], o["ids"][i, :c]), f"query {i}"
    m = g.metrics()
    assert m["rough"] == o["rough"] and m["precise"] == o["precise"]


def test_device_index_build_and_dump(case_d96, tmp_path, oracle_lib):
    """rabitq_build / rabitq_dump_to_dir (RaBitQ::from_path + dump_to_dir on the device): the directory it writes is
    read back by the oracle's load_from_dir, and both sides then agree on every query; given the same P the device
    builder reproduces the oracle builder's clustering and codes up to fp32 near-ties."""
    import rabitq_b200 as rb

    case = case_d96
    a = case["arrays"]
    g = rb.RaBitQ.build(case["base"], case["centroids"], orthogonal=a["orthogonal"], device=0)
    d = tmp_path / "built"
    g.dump_to_dir(str(d))
    o = oracle_lib.OracleIndex.load_from_dir(str(d))
    b = o.arrays()
    n = case["base"].shape[0]
    assert b["dim"] == a["dim"] and b["base"].shape == a["base"].shape
    assert sorted(b["map_ids"].tolist()) == list(range(n))
    assert np.array_equal(b["base"][:, : case["base"].shape[1]], case["base"][b["map_ids"]])
    assert b["offsets"][0] == 0 and b["offsets"][-1] == n
    # same P -> same assignment and codes except for fp32 near-ties
    assert np.mean(b["offsets"] == a["offsets"]) > 0.95
    inv_a = np.empty(n, np.int64); inv_a[a["map_ids"]] = np.arange(n)
    inv_b = np.empty(n, np.int64); inv_b[b["map_ids"]] = np.arange(n)
    assert np.mean(np.all(a["codes"][inv_a] == b["codes"][inv_b], axis=1)) > 0.999
    fa, fb = a["factors"][inv_a], b["factors"][inv_b]
    ok = np.isclose(fa, fb, rtol=1e-3, atol=1e-4).all(axis=1)
    assert ok.mean() > 0.999
    # distances sorted inside clusters
    for c in range(0, len(b["offsets"]) - 1, 5):
        rows = np.arange(b["offsets"][c], b["offsets"][c + 1])
        if len(rows) > 1:
            assert np.all(np.diff(b["factors"][rows, 3]) >= -1e-4 * np.abs(b["factors"][rows[1:], 3]) - 1e-6)
    # query parity on the device-built index
    q = case["queries"]
    gd, gi, gc = g.query_batch(q, 16, 10)
    r = o.query_batch(q, 16, 10)
    assert np.array_equal(np.sort(gd, 1).view(np.uint32), np.sort(r["dist"], 1).view(np.uint32))
    assert g.metrics()["precise"] == r["precise"]
    # self-generated rotation: orthogonal, and recall holds
    g2 = rb.RaBitQ.build(case["base"], case["centroids"], seed=5, device=0)
    d2 = tmp_path / "built2"
    g2.dump_to_dir(str(d2))
    P = oracle_lib.OracleIndex.load_from_dir(str(d2)).arrays()["orthogonal"].astype(np.float64)
    assert np.allclose(P @ P.T, np.eye(P.shape[0]), atol=2e-6)
    from tools import synth
    truth = synth.brute_force_topk_numpy(case["base"], q, 10)
    _, ids, _ = g2.query_batch(q, 48, 10)
    rec = np.mean([len(set(ids[i].tolist()) & set(truth[i].tolist())) / 10 for i in range(q.shape[0])])
    assert rec >= 0.95
    g.close(); g2.close()


def test_cli_twin_trains_saves_loads(case_d128, tmp_path):
    """rabitq_cli with the reference's flags: trains when -s does not exist, then reloads; prints QPS/recall/Metrics."""
    import os
    import subprocess

    from rabitq_b200 import build as bld
    from tools import synth

    case = case_d128
    def write_vecs(path, arr, dtype):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        with open(path, "wb") as f:
            for row in arr:
                f.write(np.uint32(len(row)).tobytes()); f.write(row.tobytes())
    write_vecs(tmp_path / "base.fvecs", case["base"], np.float32)
    write_vecs(tmp_path / "cent.fvecs", case["centroids"], np.float32)
    write_vecs(tmp_path / "q.fvecs", case["queries"], np.float32)
    truth = synth.brute_force_topk_numpy(case["base"], case["queries"], 10)
    write_vecs(tmp_path / "truth.ivecs", truth, np.int32)
    cmd = [bld.CLI, "-b", str(tmp_path / "base.fvecs"), "-c", str(tmp_path / "cent.fvecs"), "-q", str(tmp_path / "q.fvecs"),
           "-t", str(tmp_path / "truth.ivecs"), "-p", "32", "-k", "10", "-s", str(tmp_path / "saved")]
    out1 = subprocess.run(cmd, capture_output=True, text=True)
    assert out1.returncode == 0, out1.stderr
    assert "training..." in out1.stderr and "QPS:" in out1.stderr and "Metrics [query: 64," in out1.stderr
    assert os.path.exists(tmp_path / "saved" / "x_binary_vec.u64vecs")
    out2 = subprocess.run(cmd + ["--single"], capture_output=True, text=True)
    assert out2.returncode == 0, out2.stderr
    assert "loading from" in out2.stderr
    rec = float(out2.stderr.split("recall: ")[1].split()[0])
    assert rec >= 0.9


def test_raw_quantiser_with_explicit_bias(pair):
    """the reference's non-AVX2 quantiser (`scalar_quantize_raw`, src/utils.rs:194-209: truncate + rand_bias)
    with the bias handed to both sides: planes, sums, rough distances and final results equal the oracle's raw branch."""
    q = pair["queries"]
    g, o = pair["gpu"], pair["oracle"]
    rng = np.random.default_rng(123)
    bias = rng.random(g.dim, dtype=np.float32)  # Uniform[0, 1), gen_random_bias (src/utils.rs:36-40)
    probe = 12
    avx_planes = g.stage_quantize(q, probe)[3]
    g.set_quantize_bias(bias)
    o.set_raw_bias(bias)
    try:
        lo, delta, s, planes = g.stage_quantize(q, probe)
        assert not np.array_equal(planes, avx_planes)  # it really is a different rounding
        for i in range(q.shape[0]):
            tr = o.trace(q[i], probe, 10)
            assert np.array_equal(s[i], tr["sum"])
            assert np.array_equal(planes[i], tr["planes"])
            assert np.array_equal(lo[i].view(np.uint32), tr["lo"].view(np.uint32))
        g.metrics_reset()
        gd, gi, gc = g.query_batch(q, 24, 10)
        r = o.query_batch(q, 24, 10)
        for i in range(q.shape[0]):
            c = int(gc[i])
            assert _same_up_to_ties(pair, i, gd[i, :c], gi[i, :c], r["dist"][i, :c], r["ids"][i, :c]), f"query {i}"
        assert g.metrics()["precise"] == r["precise"]
    finally:
        g.set_quantize_bias(None)
        o.set_raw_bias(None)
    # back on the AVX2 semantics
    assert np.array_equal(g.stage_quantize(q, probe)[3], avx_planes)


# ---- tensor-core prefilter of the centroid scan (prefilter.cuh): probe lists must stay bit-identical ------------------------
_PF_CASES = {}


def _pf_case(name):
    from tests.conftest import make_case

    if name not in _PF_CASES:
        spec = {"k1024_d128": (20000, 128, 32, 1024, "sift", 21), "k600_d96": (9000, 96, 24, 600, "deep", 22),
                "k512_d960": (5000, 960, 16, 512, "gist", 23)}[name]
        _PF_CASES[name] = make_case(name, *spec)
    return _PF_CASES[name]


@pytest.mark.parametrize("name", ["k1024_d128", "k600_d96", "k512_d960"])
def test_prefilter_probe_lists_bit_exact(oracle_lib, name):
    """K >= 512 and probe <= K/8: approximate TF32 keys on the tensor cores, exact recheck of the candidates.  The probe
    ids and distances must equal the oracle's (and the classic path's), the fallback path as well."""
    case = _pf_case(name)
    if "gpu" not in case:
        case["gpu"] = _gpu_index(case)
    g, q = case["gpu"], case["queries"]
    K = g.num_clusters
    for probe in (1, 7, 32, K // 8):
        tr = [case["oracle"].trace(q[i], probe, 5) for i in range(q.shape[0])]
        for opt, val in (("prefilter", 1), ("prefilter_mode", 3), ("prefilter_cap", 1), ("prefilter", 0)):
            g.set_option(opt, val)   # plain TF32 keys, the 3xTF32 split, the device-side fallback, the classic path
            try:
                _, pid, pd = g.stage_probe(q, probe, want_all=False)
            finally:
                g.set_option("prefilter", 1)
                g.set_option("prefilter_cap", 1024)
                g.set_option("prefilter_mode", 1)
            for i in range(q.shape[0]):
                assert np.array_equal(pid[i], tr[i]["probe_ids"]), (name, probe, opt, val, i)
                assert np.array_equal(pd[i].view(np.uint32), tr[i]["probe_dist"].view(np.uint32)), (name, probe, opt, val, i)


def test_prefilter_end_to_end_and_counters(oracle_lib):
    case = _pf_case("k1024_d128")
    if "gpu" not in case:
        case["gpu"] = _gpu_index(case)
    g, q = case["gpu"], case["queries"]
    g.metrics_reset()
    gd, gi, gc = g.query_batch(q, 64, 10)
    o = case["oracle"].query_batch(q, 64, 10)
    for i in range(q.shape[0]):
        c = int(gc[i])
        assert _same_up_to_ties(case, i, gd[i, :c], gi[i, :c], o["dist"][i, :c], o["ids"][i, :c]), f"query {i}"
    m = g.metrics()
    assert m["rough"] == o["rough"] and m["precise"] == o["precise"]
