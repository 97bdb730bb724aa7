"""Generates tests/golden/small_d64.npz and small_d192.npz: self-contained input + every intermediate of the query path.

PARITY UNPINNED: the reference (Rust) cannot run in this image and ships no golden vectors, so these are outputs of
oracle/ (the C++ restatement), committed as REGRESSION pins: (1) the oracle must keep reproducing them bit-for-bit,
(2) the CUDA path must reproduce them without the oracle being present.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from tools import synth  # noqa: E402


def make(name, n, dim, nq, k, flavour, seed, probe, topk):
    base, queries, cent = synth.make_numpy(n, dim, nq, k, flavour, seed)
    ix = orc.OracleIndex.from_arrays(base, cent, seed=seed + 1, nthreads=1)
    a = ix.arrays()
    out = {"in_" + key: val for key, val in a.items() if key != "dim"}
    out["dim"] = np.int64(a["dim"])
    out["queries"] = queries
    out["probe"], out["topk"] = np.int64(probe), np.int64(topk)
    ys, cds, pids, pds, los, deltas, sums, planes, roughs, abdps, starts, res_d, res_i, precise = ([] for _ in range(14))
    t = 0
    for i in range(nq):
        tr = ix.trace(queries[i], probe, topk)
        ys.append(tr["y"]); cds.append(tr["centroid_dist"]); pids.append(tr["probe_ids"]); pds.append(tr["probe_dist"])
        los.append(tr["lo"]); deltas.append(tr["delta"]); sums.append(tr["sum"]); planes.append(tr["planes"])
        roughs.append(tr["rough"]); abdps.append(tr["abdp"]); starts.append(t); t += tr["pairs"]
        r = sorted(tr["result"])
        res_d.append([x[0] for x in r]); res_i.append([x[1] for x in r]); precise.append(tr["precise"])
    starts.append(t)
    out.update(y=np.stack(ys), centroid_dist=np.stack(cds), probe_ids=np.stack(pids), probe_dist=np.stack(pds), lo=np.stack(los),
               delta=np.stack(deltas), sum=np.stack(sums), planes=np.stack(planes), rough=np.concatenate(roughs),
               abdp=np.concatenate(abdps), pair_start=np.array(starts, np.uint64), result_dist=np.array(res_d, np.float32),
               result_ids=np.array(res_i, np.uint32), precise=np.array(precise, np.uint64))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", "pairs", t)


if __name__ == "__main__":
    make("small_d64", 700, 64, 12, 8, "sift", 21, 5, 10)
    make("small_d192", 500, 130, 8, 6, "deep", 22, 4, 5)   # 130 dims -> padded to 192
