"""Multi-GPU host logic: one process per GPU (torch.distributed), index sharded by cluster range, queries replicated,
per-shard top-k exchanged with ONE all-gather of (dist, id) pairs and merged by the K6 CUDA kernel.

The reference is single-process; this module is the only multi-GPU surface.  The collective is
backend-agnostic (NCCL over NVLink on the GPU box, gloo in the CPU tests); the merge itself only exists as a CUDA kernel
(`rabitq_merge_topk_device`) -- there is no CPU merge in the product.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import RaBitQ, _check, lib


def shard_range(offsets, rank: int, world: int) -> tuple[int, int]:
    """Rows [lo, hi) of the cluster-sorted arrays owned by `rank` (contiguous cluster ids, balanced by vector count)."""
    off = np.ascontiguousarray(offsets, dtype=np.uint32)
    lo, hi = C.c_size_t(0), C.c_size_t(0)
    _check(lib().rabitq_shard_range(C.c_void_p(off.ctypes.data), len(off) - 1, rank, world, C.byref(lo), C.byref(hi)))
    return int(lo.value), int(hi.value)


def all_gather_topk(dist_t, ids_t, group=None):
    """dist_t / ids_t: [nq, k] torch tensors (same device on every rank).  Returns ([world, nq, k], [world, nq, k])."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    gd = torch.empty((world,) + tuple(dist_t.shape), dtype=dist_t.dtype, device=dist_t.device)
    gi = torch.empty((world,) + tuple(ids_t.shape), dtype=ids_t.dtype, device=ids_t.device)
    dist.all_gather_into_tensor(gd.view(-1), dist_t.contiguous().view(-1), group=group)
    dist.all_gather_into_tensor(gi.view(-1), ids_t.contiguous().view(-1), group=group)
    return gd, gi


def merge_topk(gd, gi):
    """K6 on the device: [n_lists, nq, k] -> ascending [nq, k] (+ counts)."""
    import torch

    assert gd.is_cuda and gi.is_cuda and gd.is_contiguous() and gi.is_contiguous()
    n_lists, nq, k = gd.shape
    od = torch.empty((nq, k), dtype=torch.float32, device=gd.device)
    oi = torch.empty((nq, k), dtype=torch.int32, device=gd.device)
    oc = torch.empty((nq,), dtype=torch.int32, device=gd.device)
    torch.cuda.current_stream(gd.device).synchronize()
    _check(lib().rabitq_merge_topk_device(gd.device.index, C.c_void_p(gd.data_ptr()), C.c_void_p(gi.data_ptr()), n_lists, nq, k,
                                          C.c_void_p(od.data_ptr()), C.c_void_p(oi.data_ptr()), C.c_void_p(oc.data_ptr())))
    return od, oi, oc


class ShardedRaBitQ:
    """`RaBitQ` whose base set and IVF clusters are sharded over the ranks of a torch.distributed group."""

    def __init__(self, shard: RaBitQ, group=None):
        self.shard, self.group = shard, group

    @classmethod
    def load_from_dir(cls, path, device: int, group=None) -> "ShardedRaBitQ":
        import torch.distributed as dist

        return cls(RaBitQ.load_from_dir(path, device, dist.get_rank(group), dist.get_world_size(group)), group)

    @classmethod
    def from_arrays(cls, *arrays, device: int, group=None) -> "ShardedRaBitQ":
        import torch.distributed as dist

        return cls(RaBitQ.from_arrays(*arrays, device=device, shard_rank=dist.get_rank(group), shard_count=dist.get_world_size(group)), group)

    def query_batch(self, queries, probe: int, topk: int):
        """queries: CUDA tensor [nq, len], identical on every rank.  Returns merged (dist, ids, count) on every rank."""
        d, i, _ = self.shard.query_batch(queries, probe, topk)
        gd, gi = all_gather_topk(d, i, self.group)
        return merge_topk(gd, gi)
