"""Per-query statistics of the rerank kernel (K5) on a bench workload: where the tail of the launch comes from.
    python tools/rerank_stats.py c2 [key=value ...]      # needs a GPU; options go to RaBitQ.set_option"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    wl_name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    opts = dict(kv.split("=") for kv in sys.argv[2:])
    args = argparse.Namespace(shape=None, workload=wl_name, nq=0, truth_queries=1, builder="native")
    dev = torch.device("cuda", 0)
    wl = bench.build_workload(args, dev, 0, 1)
    g, q = wl["handle"], wl["queries"]
    probe = {"c1": 64, "c2": 16, "c3": 64, "c4": 32, "c5": 64}[wl_name]
    for k, v in opts.items():
        g.set_option(k, int(v))
    g.set_option("debug_rerank", 1)
    for _ in range(3):
        g.query_batch(q, probe, 10)
    t = g.last_timings()
    st = g.debug_rerank_stats(q.shape[0]).astype(np.float64)
    print(f"{wl_name} {opts}: rerank {t['ms_rerank']:.4f} ms scan {t['ms_scan']:.4f} total {t['ms_total']:.4f} exact {t['exact_computed']} precise {t['precise']}")
    pct = [50, 90, 99, 100]
    for r in range(2):
        for i, name in enumerate(["cycles", "waves", "computed", "enq|pblock", "wait|cwait", "l2|cbusy", "replay|rwait", "stage|ptotal"]):
            v = st[:, r, i]
            print(f"  round {r + 1} {name:13s} mean {v.mean():10.1f}  " + "  ".join(f"p{p}={np.percentile(v, p):9.0f}" for p in pct))
        c, w = st[:, r, 0], np.maximum(st[:, r, 1], 1)
        print(f"  round {r + 1} cycles/wave: mean {np.mean(c / w):.0f}; corr(cycles, waves) {np.corrcoef(c, st[:, r, 1])[0, 1]:.3f}")


if __name__ == "__main__":
    main()
