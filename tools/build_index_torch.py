"""Index build on the device with torch ops (harness for the bench sizes; the product is the QUERY path).

Restates `RaBitQ::from_path` (reference src/rabitq.rs:159-265, formulas in SURVEY.md section 3.4) with batched tensor
ops so that a 1M x 960 index is built in seconds on the GPU box instead of minutes on its CPU.  The reference draws P
from an unseeded RNG and multiplies with faer, so no bit-parity target exists for the builder: both the CUDA path and
the oracle consume the SAME arrays this function returns, which is what parity needs.
"""
from __future__ import annotations

import math

import torch

EPSILON = 1.9              # src/consts.rs:6
DEFAULT_X_DOT_PRODUCT = 0.8  # src/consts.rs:4


@torch.no_grad()
def build_index(base: torch.Tensor, centroids: torch.Tensor, seed: int = 1, chunk: int | None = None) -> dict:
    """base [n, len], centroids [k, len] (original space, same device).  Returns the arrays of `struct RaBitQ`
    (src/rabitq.rs:57-68) as device tensors: dim, base (padded, cluster-sorted, UNROTATED), orthogonal [D, D]
    (row r = P[r,:]), centroids [k, D] ROTATED, offsets [k+1] i32, map_ids [n] i32, codes [n, D/64] i64,
    factors [n, 4] f32 (ip, ppc, err, cds)."""
    dev = base.device
    n, ln = base.shape
    k = centroids.shape[0]
    D = (ln + 63) // 64 * 64  # rabitq.rs:167-179
    if chunk is None:  # keep the chunk x k distance tile around 1 GiB
        chunk = max(1 << 12, min(1 << 17, (1 << 28) // max(k, 1)))
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        # rabitq.rs:182 / utils.rs:16-20: Q factor of a standard-normal D x D matrix
        P = torch.linalg.qr(torch.randn(D, D, device=dev, dtype=torch.float64, generator=g))[0].to(torch.float32).contiguous()
        cent_pad = torch.zeros(k, D, device=dev)
        cent_pad[:, :ln] = centroids
        cent_rot = (cent_pad @ P).contiguous()  # rabitq.rs:189
        c2 = (cent_rot * cent_rot).sum(1)
        label = torch.empty(n, dtype=torch.int64, device=dev)
        cds = torch.empty(n, device=dev)
        fac = torch.empty(n, 4, device=dev)
        W = D // 64
        codes = torch.empty(n, W, dtype=torch.int64, device=dev)
        shifts = torch.arange(64, device=dev, dtype=torch.int64)
        dim_sqrt = math.sqrt(D)
        error_base = 2.0 * EPSILON / math.sqrt(D - 1.0)  # rabitq.rs:220
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            xb = torch.zeros(e - s, D, device=dev)
            xb[:, :ln] = base[s:e]
            xp = xb @ P  # rabitq.rs:188
            d = (xp * xp).sum(1, keepdim=True) - 2.0 * (xp @ cent_rot.T) + c2[None, :]
            lab = d.argmin(1)  # utils.rs:261-277 (first minimum)
            r = xp - cent_rot[lab]  # rabitq.rs:205
            norm = r.norm(dim=1)  # :206
            cd = norm * norm  # :207
            bits = r > 0  # utils.rs:53-67
            sgn_sum = (2.0 * bits.sum(1).to(torch.float32) - D)
            dot = r.abs().sum(1)  # <r, sign(r)>, zeros contribute 0 either way
            nrm = norm * dim_sqrt
            ok = torch.isfinite(nrm) & (nrm >= torch.finfo(torch.float32).tiny)  # is_normal(), :211
            xdp = torch.where(ok, dot / nrm, torch.full_like(dot, DEFAULT_X_DOT_PRODUCT))
            t = norm / xdp  # :223
            fac[s:e, 2] = error_base * torch.sqrt(t * t - cd)  # :225-226
            ip = (-2.0 / dim_sqrt) * t  # :227
            fac[s:e, 0] = ip
            fac[s:e, 1] = ip * sgn_sum  # :228
            fac[s:e, 3] = cd
            label[s:e] = lab
            cds[s:e] = cd
            codes[s:e] = (bits.view(e - s, W, 64).to(torch.int64) << shifts).sum(2)
        # rabitq.rs:231-252: stable sort by distance inside each cluster, clusters in id order
        o1 = torch.sort(cds, stable=True).indices
        o2 = torch.sort(label[o1], stable=True).indices
        perm = o1[o2]
        counts = torch.bincount(label, minlength=k)
        offsets = torch.zeros(k + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(counts, 0)
        base_sorted = torch.zeros(n, D, device=dev)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            base_sorted[s:e, :ln] = base[perm[s:e]]
        out = dict(dim=D, base=base_sorted, orthogonal=P, centroids=cent_rot, offsets=offsets.to(torch.int32).contiguous(),
                   map_ids=perm.to(torch.int32).contiguous(), codes=codes[perm].contiguous(), factors=fac[perm].contiguous())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


def to_numpy(ix: dict) -> dict:
    import numpy as np

    return dict(dim=ix["dim"], base=ix["base"].cpu().numpy(), orthogonal=ix["orthogonal"].cpu().numpy(),
                centroids=ix["centroids"].cpu().numpy(), offsets=ix["offsets"].cpu().numpy().astype(np.uint32),
                map_ids=ix["map_ids"].cpu().numpy().astype(np.uint32), codes=ix["codes"].cpu().numpy().view(np.uint64),
                factors=ix["factors"].cpu().numpy())
