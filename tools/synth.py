"""Seeded synthetic datasets of the BASELINE.json shapes (harness, not product).

SURVEY.md section 8(d): the data must be *clustered*; an i.i.d. Gaussian has no IVF structure.  Real embedding sets have a
low intrinsic dimension and overlapping clusters, so that recall grows gradually with nprobe.  We emulate that with a
mixture of `k` Gaussians in a LATENT space of `latent` dims, pushed to `dim` dims by a fixed random linear map plus
small isotropic noise, then given the flavour of the named dataset (non-negative integers for SIFT, small non-negative
floats for GIST, unit rows for DEEP / embeddings).  The mixture means (mapped the same way) are the "precomputed IVF
centroids" the configs name.  Everything runs in torch on the given device (CPU for tests, CUDA for the bench).
"""
from __future__ import annotations

import numpy as np
import torch

# name -> (n, dim, n_queries, k, flavour)
SHAPES = {
    "c1": (1_000_000, 128, 10_000, 4096, "sift"),
    "c2": (1_000_000, 960, 1_000, 1024, "gist"),
    "c3": (10_000_000, 96, 10_000, 16384, "deep"),
    "c4": (10_000_000, 1536, 10_000, 8192, "embed"),
    "c5": (100_000_000, 128, 65_536, 65536, "sift"),
}
SEEDS = {"c1": 1001, "c2": 2001, "c3": 3001, "c4": 4001, "c5": 5001}

# flavour -> (latent dims, component sigma range in latent units, ambient noise, post-transform)
_FLAVOUR = {
    "sift": dict(latent=24, sig=(0.55, 0.95), noise=0.10, post="sift"),
    "gist": dict(latent=32, sig=(0.55, 0.95), noise=0.10, post="gist"),
    "deep": dict(latent=24, sig=(0.55, 0.95), noise=0.10, post="unit"),
    "embed": dict(latent=48, sig=(0.55, 0.95), noise=0.10, post="unit"),
}


class Mixture:
    """Frozen mixture parameters (so base and queries come from the same distribution)."""

    def __init__(self, dim: int, k: int, flavour: str, seed: int, device):
        f = _FLAVOUR[flavour]
        self.dim, self.k, self.flavour, self.device = dim, k, flavour, torch.device(device)
        self.post, self.noise = f["post"], f["noise"]
        g = torch.Generator(device=self.device)
        g.manual_seed(seed)
        L = f["latent"]
        self.means_z = torch.randn(k, L, device=self.device, generator=g)
        self.sig = torch.empty(k, 1, device=self.device).uniform_(f["sig"][0], f["sig"][1], generator=g)
        self.A = torch.randn(L, dim, device=self.device, generator=g) / float(L) ** 0.5
        # skewed component weights: real IVF lists are far from equal-sized
        self.weights = torch.exp(0.7 * torch.randn(k, device=self.device, generator=g))
        self.shift = torch.empty(1, dim, device=self.device).uniform_(-0.5, 0.5, generator=g)

    def _post(self, x: torch.Tensor) -> torch.Tensor:
        if self.post == "sift":      # non-negative integers, SIFT-like range
            return (x * 24.0 + 48.0).clamp_(min=0).round_()
        if self.post == "gist":      # small non-negative floats
            return (x * 0.05 + 0.12).clamp_(min=0)
        return x / x.norm(dim=1, keepdim=True).clamp_(min=1e-12)

    def draw(self, m: int, seed: int, chunk: int = 1 << 18) -> torch.Tensor:
        g = torch.Generator(device=self.device)
        g.manual_seed(seed)
        out = torch.empty(m, self.dim, device=self.device)
        for s in range(0, m, chunk):
            e = min(m, s + chunk)
            comp = torch.multinomial(self.weights, e - s, replacement=True, generator=g)
            z = self.means_z[comp] + self.sig[comp] * torch.randn(e - s, self.means_z.shape[1], device=self.device, generator=g)
            x = z @ self.A + self.shift + self.noise * torch.randn(e - s, self.dim, device=self.device, generator=g)
            out[s:e] = self._post(x)
        return out

    def centroids(self) -> torch.Tensor:
        return self._post(self.means_z @ self.A + self.shift)


def make_torch(name_or_shape, device="cpu", seed: int | None = None):
    """(base [n, dim], queries [nq, dim], centroids [k, dim]) float32 tensors on `device`."""
    if isinstance(name_or_shape, str):
        n, dim, nq, k, flavour = SHAPES[name_or_shape]
        seed = SEEDS[name_or_shape] if seed is None else seed
    else:
        n, dim, nq, k, flavour = name_or_shape
        seed = 1 if seed is None else seed
    mix = Mixture(dim, k, flavour, seed + 2, device)
    return mix.draw(n, seed), mix.draw(nq, seed + 1), mix.centroids()


def make_numpy(n: int, dim: int, nq: int, k: int, flavour: str = "sift", seed: int = 1):
    """Small datasets for the tests (generated with torch on the CPU, returned as numpy)."""
    b, q, c = make_torch((n, dim, nq, k, flavour), "cpu", seed)
    return b.numpy(), q.numpy(), c.numpy()


def brute_force_topk_numpy(base: np.ndarray, queries: np.ndarray, topk: int) -> np.ndarray:
    b = base.astype(np.float64)
    b2 = (b ** 2).sum(1)
    out = np.empty((queries.shape[0], topk), np.int32)
    for i in range(0, queries.shape[0], 256):
        q = queries[i:i + 256].astype(np.float64)
        d = b2[None, :] - 2.0 * q @ b.T
        out[i:i + 256] = np.argsort(d, axis=1, kind="stable")[:, :topk]
    return out


def brute_force_topk_torch(base: torch.Tensor, queries: torch.Tensor, topk: int, chunk: int | None = None) -> torch.Tensor:
    """Exact fp32 top-k ids by squared L2 on the device (ground truth for recall)."""
    nq = queries.shape[0]
    if chunk is None:  # keep the nq x chunk distance tile around 1 GiB
        chunk = max(1 << 12, min(1 << 19, (1 << 28) // max(nq, 1)))
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
