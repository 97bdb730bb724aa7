"""Dev probe: prefilter behaviour at K = 65536 (timings per batch, mode changes with RABITQ_TRACE=1)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import rabitq_b200 as rb
from tools import synth
dev = torch.device('cuda', 0)
n, dim, nq, k = 3_000_000, 128, 2048, 65536
mix = synth.Mixture(dim, k, 'sift', 5003, dev)
base = mix.draw(n, 5001); q = mix.draw(nq, 5002); cent = mix.centroids()
g = rb.RaBitQ.build(base.contiguous(), cent.contiguous(), seed=7, device=0)
ref = None
for mode in (1, 1, 3, 0):
    g.set_option("prefilter_mode", mode) if mode else g.set_option("prefilter", 0)
    d, i, c = g.query_batch(q.contiguous(), 64, 10)
    t = g.last_timings()
    same = True if ref is None else bool(torch.equal(d, ref[0]) and torch.equal(i, ref[1]))
    ref = ref or (d.clone(), i.clone())
    print(mode, same, {k2: round(v, 3) for k2, v in t.items() if k2.startswith('ms_') and v > 0.01}, flush=True)
