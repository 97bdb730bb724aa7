"""Cross-check of the query path at a bench workload's scale on ONE GPU: K5 in its CTA form vs its warp form vs the distributed
pipeline driven as N virtual ranks (the collectives become device copies).   python tools/scale_check.py c5 65536 8"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rabitq_b200 import distributed as D

wl_name = sys.argv[1] if len(sys.argv) > 1 else "c5"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
world = int(sys.argv[3]) if len(sys.argv) > 3 else 2
args = argparse.Namespace(shape=None, workload=wl_name, nq=nq, truth_queries=1, builder="native")
dev = torch.device("cuda", 0)
wl = bench.build_workload(args, dev, 0, 1)
g, q = wl["handle"], wl["queries"]
probe = {"c1": 64, "c2": 16, "c3": 64, "c4": 32, "c5": 64}[wl_name]
res = {}
for mode in (1, 0):
    g.set_option("rerank_mode", mode)
    g.metrics_reset()
    d, i, c = g.query_batch(q, probe, 10)
    torch.cuda.synchronize()
    res[mode] = (d.cpu().numpy(), i.cpu().numpy(), c.cpu().numpy(), g.metrics()["precise"])
    print(f"mode {mode}: precise {res[mode][3]}", flush=True)
a, b = res[1], res[0]
bad = np.where((a[0].view(np.uint32) != b[0].view(np.uint32)).any(axis=1))[0]
print(f"new-vs-old K5: {len(bad)} of {nq} queries differ; first {bad[:10].tolist()}", flush=True)
g.set_option("rerank_mode", 1)
# distributed pipeline, virtual ranks on this GPU
shards = [g.reshard(r, world) for r in range(world)]
st = torch.cuda.Stream(dev)
with torch.cuda.stream(st):
    dd, di, dc, _ = D.run_virtual_ranks(shards, q, probe, 10)
    st.synchronize()
dd = dd.cpu().numpy()
prec = sum(s.metrics()["precise"] for s in shards)
for name, ref in (("new", a), ("old", b)):
    bad = np.where((dd.view(np.uint32) != ref[0].view(np.uint32)).any(axis=1))[0]
    print(f"distributed (virtual x{world}) vs {name} K5: {len(bad)} of {nq} queries differ; first {bad[:10].tolist()}", flush=True)
bad = np.where((dd.view(np.uint32) != b[0].view(np.uint32)).any(axis=1))[0]
print("bad queries home ranks:", np.bincount(bad // (nq // world), minlength=world).tolist())
for qq in bad[:3]:
    print("query", qq, "\n dist  ", dd[qq], "\n single", b[0][qq], "\n ids dist", di[qq].cpu().numpy(), "\n ids single", b[1][qq])
