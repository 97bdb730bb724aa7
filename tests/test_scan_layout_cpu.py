"""CPU: the index algebra of the tensor-core code scan (K4) and of the harness around it, without a GPU.

* `rec_pos` (the byte order K3 writes and K4's B fragments read) is a permutation of every k-step pair, and the library's
  own function agrees with the formula the simulation uses;
* a lane-accurate simulation of one warp of `scan_mma_kernel` -- PTX fragment layouts of mma.m16n8k32.u8, A fragments built
  from packed code words with one AND, records in `rec_pos` order, ballots rotated into bitmap words -- reproduces
  8 * sum_d bit_d * q_d for every accumulator element and vector-ordered bitmaps (tools/sim_mma_scan.py);
* bench.py: result comparison helper, and the round-1 crash guard (arrays are never exported from a resharded handle).
"""
import importlib.util
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rec_pos(d):
    j, r = d >> 5, d & 31
    return 64 * (j >> 1) + 16 * (r & 3) + 8 * (j & 1) + 4 * ((r >> 2) & 1) + (r >> 3)


def test_rec_pos_is_a_permutation_and_matches_the_library():
    import rabitq_b200 as rb

    L = rb.lib()
    for D in (64, 128, 192, 960, 3072, 8192):
        pos = [_rec_pos(d) for d in range(D)]
        assert sorted(pos) == list(range(D))
        assert all(L.rabitq_debug_rec_pos(d) == pos[d] for d in range(D))
        for d in range(D):  # a k-step pair (64 dimensions) stays inside its own 64-byte group: one 128-bit load per lane covers two k-steps
            assert pos[d] // 64 == d // 64


def test_warp_simulation_of_the_mma_scan():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sim_mma_scan.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "sim ok" in out.stdout


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    src = open(os.path.join(ROOT, "bench.py")).read()
    # bench.py redirects fd 1 at import time (one JSON line on stdout); load only the pure helper
    ns = {}
    m = re.search(r"def compare_results\(.*?\n(?=\n\ndef )", src, re.S)
    exec("import numpy as np\n" + m.group(0), ns)
    return src, ns["compare_results"]


def test_bench_compare_results_and_shard_export_guard():
    src, cmp_ = _bench()
    d = np.array([[1.0, 2.0, 3.0], [1.0, 2.0, 2.0]], np.float32)
    i = np.array([[7, 8, 9], [1, 2, 3]], np.uint32)
    c = np.array([3, 3], np.uint32)
    r = cmp_(d, i, c, d.copy(), i.copy(), c.copy())
    assert r["dist_bit_identical"] and r["ids_identical_up_to_ties"] and r["ids_identical_strict"] == 2
    i2 = i.copy(); i2[1, 2] = 99           # differs on a distance that is tied inside the list: allowed
    r = cmp_(d, i, c, d.copy(), i2, c.copy())
    assert r["dist_bit_identical"] and r["ids_identical_up_to_ties"] and r["ids_identical_strict"] == 1
    i3 = i.copy(); i3[0, 0] = 99           # differs on an untied distance... which is still the same distance value on both sides
    d3 = d.copy(); d3[0, 0] = 0.5
    r = cmp_(d, i, c, d3, i3, c.copy())
    assert not r["dist_bit_identical"] and r["mismatching_queries"] == [0]
    # the crash of round 1: cpu_baseline exported arrays from the RESHARDED handle.  The oracle is built from wl["single"] / arrays only.
    assert 'oracle_from_index(wl["handle"])' not in src
    assert "keep_single" in src and 'cpu_baseline(wl, single' in src
    # both arms print the same workload string
    assert src.count("workload_string(wl, probe)") >= 2
