"""Seeded synthetic datasets of the BASELINE.json shapes (harness, not product).

SURVEY.md section 8(d): the data must be *clustered* (an i.i.d. Gaussian has no IVF structure).  Each dataset is
a mixture of `k` Gaussians; the mixture means double as the "precomputed IVF centroids" the configs name.
numpy path for the small CPU/GPU-parity cases, torch path (device-resident, chunked) for the bench sizes.
"""
from __future__ import annotations

import numpy as np

# name -> (n, dim, n_queries, k, flavour)
SHAPES = {
    "c1": (1_000_000, 128, 10_000, 4096, "sift"),
    "c2": (1_000_000, 960, 1_000, 1024, "gist"),
    "c3": (10_000_000, 96, 10_000, 16384, "deep"),
    "c4": (10_000_000, 1536, 10_000, 8192, "embed"),
    "c5": (100_000_000, 128, 65_536, 65536, "sift"),
}
SEEDS = {"c1": 1001, "c2": 2001, "c3": 3001, "c4": 4001, "c5": 5001}


def _flavour_params(flavour: str):
    # (mean_lo, mean_hi, sigma_lo, sigma_hi, clip_nonneg, round_int, normalise)
    return {
        "sift": (0.0, 128.0, 18.0, 36.0, True, True, False),
        "gist": (0.0, 0.3, 0.04, 0.08, True, False, False),
        "deep": (-1.0, 1.0, 0.35, 0.7, False, False, True),
        "embed": (-1.0, 1.0, 0.35, 0.7, False, False, True),
    }[flavour]


def make_numpy(n: int, dim: int, nq: int, k: int, flavour: str = "sift", seed: int = 1):
    """Small datasets on the CPU: returns (base[n,dim], queries[nq,dim], centroids[k,dim]) float32."""
    lo, hi, slo, shi, clip, rnd, norm = _flavour_params(flavour)
    rng = np.random.default_rng(seed)
    means = rng.uniform(lo, hi, size=(k, dim)).astype(np.float32)
    sig = rng.uniform(slo, shi, size=(k, 1)).astype(np.float32)

    def draw(m, r):
        comp = r.integers(0, k, size=m)
        x = means[comp] + sig[comp] * r.standard_normal((m, dim)).astype(np.float32)
        if clip:
            x = np.maximum(x, 0)
        if rnd:
            x = np.round(x)
        if norm:
            x = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
        return x.astype(np.float32)

    base = draw(n, np.random.default_rng(seed + 1))
    queries = draw(nq, np.random.default_rng(seed + 2))
    cent = means.copy()
    if norm:
        cent = cent / np.maximum(np.linalg.norm(cent, axis=1, keepdims=True), 1e-12)
    return base, queries, cent.astype(np.float32)


def brute_force_topk_numpy(base: np.ndarray, queries: np.ndarray, topk: int) -> np.ndarray:
    b2 = (base.astype(np.float64) ** 2).sum(1)
    out = np.empty((queries.shape[0], topk), np.int32)
    for i in range(0, queries.shape[0], 256):
        q = queries[i:i + 256].astype(np.float64)
        d = b2[None, :] - 2.0 * q @ base.astype(np.float64).T
        out[i:i + 256] = np.argsort(d, axis=1, kind="stable")[:, :topk]
    return out


def make_torch(name_or_shape, device, seed: int | None = None, chunk: int = 1 << 18):
    """Bench-size datasets generated on `device` with torch.  Returns (base, queries, centroids) tensors
    (float32, base is [n, dim]).  `name_or_shape` is a key of SHAPES or a tuple (n, dim, nq, k, flavour)."""
    import torch

    if isinstance(name_or_shape, str):
        n, dim, nq, k, flavour = SHAPES[name_or_shape]
        seed = SEEDS[name_or_shape] if seed is None else seed
    else:
        n, dim, nq, k, flavour = name_or_shape
        seed = 1 if seed is None else seed
    lo, hi, slo, shi, clip, rnd, norm = _flavour_params(flavour)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    means = torch.empty(k, dim, device=device).uniform_(lo, hi, generator=g)
    sig = torch.empty(k, 1, device=device).uniform_(slo, shi, generator=g)

    def draw(m, gen):
        out = torch.empty(m, dim, device=device)
        for s in range(0, m, chunk):
            e = min(m, s + chunk)
            comp = torch.randint(0, k, (e - s,), device=device, generator=gen)
            x = means[comp] + sig[comp] * torch.randn(e - s, dim, device=device, generator=gen)
            if clip:
                x.clamp_(min=0)
            if rnd:
                x.round_()
            if norm:
                x = x / x.norm(dim=1, keepdim=True).clamp_(min=1e-12)
            out[s:e] = x
        return out

    g1 = torch.Generator(device=device); g1.manual_seed(seed + 1)
    g2 = torch.Generator(device=device); g2.manual_seed(seed + 2)
    base = draw(n, g1)
    queries = draw(nq, g2)
    cent = means.clone()
    if norm:
        cent = cent / cent.norm(dim=1, keepdim=True).clamp_(min=1e-12)
    return base, queries, cent


def brute_force_topk_torch(base, queries, topk: int, chunk: int = 1 << 20):
    """Exact fp32 top-k ids by squared L2 on the device (ground truth for recall)."""
    import torch

    nq = queries.shape[0]
    best_d = torch.full((nq, topk), float("inf"), device=base.device)
    best_i = torch.zeros((nq, topk), dtype=torch.int64, device=base.device)
    q2 = (queries * queries).sum(1, keepdim=True)
    for s in range(0, base.shape[0], chunk):
        b = base[s:s + chunk]
        d = q2 - 2.0 * (queries @ b.T) + (b * b).sum(1)[None, :]
        dd, ii = torch.topk(d, min(topk, d.shape[1]), dim=1, largest=False)
        cat_d = torch.cat([best_d, dd], 1)
        cat_i = torch.cat([best_i, ii + s], 1)
        best_d, sel = torch.topk(cat_d, topk, dim=1, largest=False)
        best_i = torch.gather(cat_i, 1, sel)
    return best_i.to(torch.int32)
