"""-m gpu: every kernel family of the path at the other supported dimensions (register-resident quantiser W32 in {2,6,8}, the
generic one, D = 1536 of config 4), top-k above one warp (k > 32: the shuffle-tree heap maximum), and the serving-loop entry
(`query_batch_device_into`), against the CPU oracle -- same bars as tests/test_gpu_parity.py."""
import numpy as np
import pytest

from tests.conftest import make_case
from tests.test_gpu_parity import _gpu_index, _same_up_to_ties

pytestmark = pytest.mark.gpu

DIM_CASES = {
    "d64": (5000, 64, 24, 32, "sift", 21),      # W32 = 2, one rotation column tile
    "d192": (5000, 192, 24, 32, "sift", 22),    # W32 = 6 (64-bit code loads in the scan)
    "d250": (4000, 250, 16, 24, "deep", 23),    # 250 -> padded to 256, W32 = 8
    "d448": (3000, 448, 16, 16, "gist", 24),    # generic quantiser, W32 = 14
    "d1536": (2500, 1536, 12, 16, "embed", 25),  # config 4's dimension
    "d704": (2500, 704, 12, 16, "gist", 26),     # a dim % 64 == 0 no template ever named (W32 = 22: odd count of k-step quads)
    "d3072": (1500, 3072, 8, 12, "embed", 27),   # text-embedding-3-large; 4 record tiles no longer fit next to the chunk's codes
}


@pytest.fixture(params=sorted(DIM_CASES))
def dcase(request, oracle_lib):
    n, dim, nq, k, flavour, seed = DIM_CASES[request.param]
    case = make_case(request.param, n, dim, nq, k, flavour, seed)
    if "gpu" not in case:
        case["gpu"] = _gpu_index(case)
    return case


def test_stages_bit_exact(dcase):
    q, g, o = dcase["queries"], dcase["gpu"], dcase["oracle"]
    probe = 6
    y = g.stage_rotate(q)
    cd, pid, pd = g.stage_probe(q, probe)
    lo, delta, s, planes = g.stage_quantize(q, probe)
    for i in range(q.shape[0]):
        tr = o.trace(q[i], probe, 10)
        assert np.array_equal(y[i].view(np.uint32), tr["y"].view(np.uint32)), f"y, query {i}"
        assert np.array_equal(pid[i], tr["probe_ids"])
        assert np.array_equal(pd[i].view(np.uint32), tr["probe_dist"].view(np.uint32))
        assert np.array_equal(lo[i].view(np.uint32), tr["lo"].view(np.uint32))
        assert np.array_equal(delta[i].view(np.uint32), tr["delta"].view(np.uint32))
        assert np.array_equal(s[i], tr["sum"])
        assert np.array_equal(planes[i], tr["planes"])


@pytest.mark.parametrize("probe,topk", [(5, 10), (16, 40)])
def test_end_to_end(dcase, probe, topk):
    q, g = dcase["queries"], dcase["gpu"]
    g.metrics_reset()
    gd, gi, gc = g.query_batch(q, probe, topk)
    o = dcase["oracle"].query_batch(q, probe, topk)
    assert np.array_equal(gc, o["count"])
    for i in range(q.shape[0]):
        c = int(gc[i])
        assert _same_up_to_ties(dcase, i, gd[i, :c], gi[i, :c], o["dist"][i, :c], o["ids"][i, :c]), f"query {i}"
    m = g.metrics()
    assert m["rough"] == o["rough"] and m["precise"] == o["precise"]


def test_device_entry_with_caller_owned_outputs(case_d128):
    """`query_batch_device_into` (no allocation, no device-wide sync, K5 writes the caller's tensors) == the host-pointer call;
    stage timers are resolved lazily and the per-query rerank statistics are consistent with the counters."""
    import torch

    if "gpu" not in case_d128:
        case_d128["gpu"] = _gpu_index(case_d128)
    g, q = case_d128["gpu"], case_d128["queries"]
    hd, hi, hc = g.query_batch(q, 16, 10)
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        g.set_stream(st.cuda_stream)
        qd = torch.from_numpy(q).to(dev)
        d = torch.empty((q.shape[0], 10), dtype=torch.float32, device=dev)
        i = torch.empty((q.shape[0], 10), dtype=torch.int32, device=dev)
        c = torch.empty((q.shape[0],), dtype=torch.int32, device=dev)
        g.set_option("debug_rerank", 1)
        g.query_batch_device_into(qd, 16, 10, d, i, c)
        st.synchronize()
    g.set_stream(None)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), hd.view(np.uint32))
    assert np.array_equal(i.cpu().numpy().view(np.uint32), hi)
    assert np.array_equal(c.cpu().numpy().view(np.uint32), hc)
    t = g.last_timings()
    assert t["ms_total"] > 0 and t["ms_scan"] > 0 and t["kernel_launches"] > 0
    stats = g.debug_rerank_stats(q.shape[0])
    assert int(stats[:, :, 2].sum()) == t["exact_computed"]   # speculative exact distances, both rounds
    assert np.all(stats[:, :, 0] > 0)
    g.set_option("debug_rerank", 0)
