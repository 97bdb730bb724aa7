"""Golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle; regression pins, the
reference itself has no golden vectors).  CPU: the oracle still reproduces them bit-for-bit.  GPU: the CUDA path
reproduces them through the C ABI with no oracle in the loop."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_golden(oracle_lib, path):
    g = np.load(path)
    ix = oracle_lib.OracleIndex.from_built(int(g["dim"]), g["in_base"], g["in_orthogonal"], g["in_centroids"], g["in_offsets"],
                                           g["in_map_ids"], g["in_codes"], g["in_factors"])
    probe, topk = int(g["probe"]), int(g["topk"])
    for i, q in enumerate(g["queries"]):
        tr = ix.trace(q, probe, topk)
        s, e = int(g["pair_start"][i]), int(g["pair_start"][i + 1])
        assert np.array_equal(_bits(tr["y"]), _bits(g["y"][i]))
        assert np.array_equal(_bits(tr["centroid_dist"]), _bits(g["centroid_dist"][i]))
        assert np.array_equal(tr["probe_ids"], g["probe_ids"][i])
        assert np.array_equal(_bits(tr["lo"]), _bits(g["lo"][i])) and np.array_equal(_bits(tr["delta"]), _bits(g["delta"][i]))
        assert np.array_equal(tr["sum"], g["sum"][i]) and np.array_equal(tr["planes"], g["planes"][i])
        assert np.array_equal(tr["abdp"], g["abdp"][s:e]) and np.array_equal(_bits(tr["rough"]), _bits(g["rough"][s:e]))
        r = sorted(tr["result"])
        assert np.array_equal(_bits(np.array([x[0] for x in r], np.float32)), _bits(g["result_dist"][i]))
        assert tr["precise"] == int(g["precise"][i])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_cuda_reproduces_golden(path):
    import rabitq_b200 as rb

    g = np.load(path)
    ix = rb.RaBitQ.from_arrays(int(g["dim"]), g["in_base"], g["in_orthogonal"], g["in_centroids"], g["in_offsets"], g["in_map_ids"],
                               g["in_codes"], g["in_factors"], device=0)
    probe, topk = int(g["probe"]), int(g["topk"])
    q = g["queries"]
    assert np.array_equal(_bits(ix.stage_rotate(q)), _bits(g["y"]))
    cd, pid, pd = ix.stage_probe(q, probe)
    assert np.array_equal(_bits(cd), _bits(g["centroid_dist"])) and np.array_equal(pid, g["probe_ids"])
    assert np.array_equal(_bits(pd), _bits(g["probe_dist"]))
    lo, delta, s, planes = ix.stage_quantize(q, probe)
    assert np.array_equal(_bits(lo), _bits(g["lo"])) and np.array_equal(_bits(delta), _bits(g["delta"]))
    assert np.array_equal(s, g["sum"]) and np.array_equal(planes, g["planes"])
    rough, abdp, start = ix.stage_scan(q, probe, pair_capacity=len(g["rough"]))
    assert np.array_equal(start, g["pair_start"]) and np.array_equal(abdp, g["abdp"]) and np.array_equal(_bits(rough), _bits(g["rough"]))
    ix.metrics_reset()
    d, i, c = ix.query_batch(q, probe, topk)
    assert np.array_equal(_bits(d), _bits(g["result_dist"]))
    assert ix.metrics()["precise"] == int(g["precise"].sum())
    ix.close()
