"""Multi-GPU host logic: one process per GPU (torch.distributed), the base set and IVF clusters sharded by cluster range.

Two ways to answer a batch:

* `DistributedRaBitQ` (the product path): every rank is the HOME of its own slice of the query batch.  The phases of the
  C ABI (`rabitq_dist_front_rotate / front_select / round1_split / round2 / finish`) run between the collectives on one CUDA
  stream -- an all-gather of the rotated queries that is STARTED ASYNCHRONOUSLY and overlaps the centroid scan, a small
  all-gather of the probe lists, an all-reduce(min) of the round-1 thresholds, an all-reduce(max) of a status word
  that doubles as the barrier -- while the survivor records themselves travel by peer stores from the kernel that computes
  the exact distances straight into the home rank's inbox (CUDA IPC memory over NVLink).  Results, distances and the
  `precise` counter are IDENTICAL to the single-process reference (src/rerank.rs:81-106 replayed on the union).
* `ShardedRaBitQ` (the simple scheme of SURVEY.md section 8e): queries replicated, per-shard top-k exchanged with one
  all-gather of (dist, id) pairs and merged by the K6 kernel; equal to or better than the reference's list when the
  RaBitQ bound fails, not always identical.

`run_virtual_ranks` drives the same phases for N shards held by ONE process on ONE GPU (the collectives become device
copies); the parity tests use it so that the whole distributed pipeline is checked on a single-GPU box.
The reference is single-process; this module is the only multi-GPU surface.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import RaBitQ, RabitqError, _check, lib


def shard_range(offsets, rank: int, world: int) -> tuple[int, int]:
    """Rows [lo, hi) of the cluster-sorted arrays owned by `rank` (contiguous cluster ids, balanced by vector count)."""
    off = np.ascontiguousarray(offsets, dtype=np.uint32)
    lo, hi = C.c_size_t(0), C.c_size_t(0)
    _check(lib().rabitq_shard_range(C.c_void_p(off.ctypes.data), len(off) - 1, rank, world, C.byref(lo), C.byref(hi)))
    return int(lo.value), int(hi.value)


# ---- collectives (backend-agnostic: NCCL on the GPU box, gloo in the CPU tests) -------------------------------------------
class TorchComm:
    """The three collectives of a step + the one-off exchange of inbox handles, over a torch.distributed group."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_gather(self, out_t, in_t):
        self.dist.all_gather_into_tensor(out_t, in_t, group=self.group)

    def all_gather_start(self, out_t, in_t):
        """Asynchronous all-gather: enqueued behind the work already on the current stream, runs on the backend's own stream
        while the caller keeps launching kernels; `.wait()` on the returned handle orders the current stream after it."""
        return self.dist.all_gather_into_tensor(out_t, in_t, group=self.group, async_op=True)

    def all_reduce_min(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)

    def all_reduce_max(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)

    def exchange_bytes(self, payload: bytes) -> list[bytes]:
        out = [None] * self.world
        self.dist.all_gather_object(out, payload, group=self.group)
        return out


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
    """K6 on the device, on torch's current stream: [n_lists, nq, k] -> ascending [nq, k] (+ counts)."""
    import torch

    assert gd.is_cuda and gi.is_cuda and gd.is_contiguous() and gi.is_contiguous()
    n_lists, nq, k = gd.shape
    od = torch.empty((nq, k), dtype=torch.float32, device=gd.device)
    oi = torch.empty((nq, k), dtype=torch.int32, device=gd.device)
    oc = torch.empty((nq,), dtype=torch.int32, device=gd.device)
    _check(lib().rabitq_merge_topk_device(gd.device.index, C.c_void_p(gd.data_ptr()), C.c_void_p(gi.data_ptr()), n_lists, nq, k,
                                          C.c_void_p(od.data_ptr()), C.c_void_p(oi.data_ptr()), C.c_void_p(oc.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream(gd.device).cuda_stream)))
    return od, oi, oc


class ShardedRaBitQ:
    """Replicated queries, shard-local top-k, all-gather + merge (not strictly identical to the reference; see module doc)."""

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


# ---- the product path -----------------------------------------------------------------------------------------------------
class _RankState:
    """Per-rank buffers of the distributed pipeline (torch CUDA tensors, so that the collectives can take them)."""

    def __init__(self, shard: RaBitQ, rank: int, world: int, nq_local: int, length: int, probe: int, topk: int, records_per_query: int,
                 push: bool = True):
        import torch

        self.shard, self.rank, self.world, self.push = shard, rank, world, push
        self.nq_local, self.len, self.probe, self.topk, self.rpq = nq_local, length, probe, topk, records_per_query
        nbytes = C.c_size_t(0)
        _check(lib().rabitq_dist_init(shard._h, rank, world, nq_local, probe, topk, records_per_query, C.byref(nbytes)))
        self.inbox_bytes = int(nbytes.value)
        dev = torch.device("cuda", shard.device)
        wa = int(lib().rabitq_dist_chunk_words_qy(shard._h, length))     # [q | y]: the big part, gathered while K2 runs
        wb = int(lib().rabitq_dist_chunk_words_meta(shard._h, length))   # [probe ids | probe distances | first non-empty rank]
        # push mode: the [q | y] chunks travel by copy-engine stores into the peers' inboxes, no tensors of ours are involved
        self.send_qy = None if push else torch.empty(wa, dtype=torch.int32, device=dev)
        self.gathered_qy = None if push else torch.empty(wa * world, dtype=torch.int32, device=dev)
        self.send_meta = torch.empty(wb, dtype=torch.int32, device=dev)
        self.gathered_meta = torch.empty(wb * world, dtype=torch.int32, device=dev)
        self.thr = torch.empty(nq_local * world, dtype=torch.float32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.out_d = torch.empty((nq_local, topk), dtype=torch.float32, device=dev)
        self.out_i = torch.empty((nq_local, topk), dtype=torch.int32, device=dev)
        self.out_c = torch.empty((nq_local,), dtype=torch.int32, device=dev)

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib().rabitq_dist_ipc_handle(self.shard._h, buf))
        return buf.raw

    def inbox_ptr(self) -> int:
        p = C.c_void_p()
        _check(lib().rabitq_dist_inbox_ptr(self.shard._h, C.byref(p)))
        return int(p.value)

    # phases (everything asynchronous on the shard's stream, except the one host read that sizes the survivor slots)
    def front_rotate(self, q_dev):
        assert q_dev.is_cuda and q_dev.is_contiguous() and tuple(q_dev.shape) == (self.nq_local, self.len), tuple(q_dev.shape)
        _check(lib().rabitq_dist_front_rotate(self.shard._h, C.c_void_p(q_dev.data_ptr()), self.len,
                                              None if self.push else C.c_void_p(self.send_qy.data_ptr())))

    def front_select(self):
        _check(lib().rabitq_dist_front_select(self.shard._h, C.c_void_p(self.send_meta.data_ptr())))

    def round1(self):
        _check(lib().rabitq_dist_round1_split(self.shard._h, None if self.push else C.c_void_p(self.gathered_qy.data_ptr()),
                                              C.c_void_p(self.gathered_meta.data_ptr()), C.c_void_p(self.thr.data_ptr())))

    def round2(self):
        _check(lib().rabitq_dist_round2(self.shard._h, C.c_void_p(self.status.data_ptr())))

    def finish(self) -> int:
        """Runs the home replay, synchronises the stream and returns the step's status word (0 = results valid)."""
        _check(lib().rabitq_dist_finish(self.shard._h, C.c_void_p(self.out_d.data_ptr()), C.c_void_p(self.out_i.data_ptr()),
                                        C.c_void_p(self.out_c.data_ptr()), C.c_void_p(self.status.data_ptr())))
        st = C.c_uint32(0)
        _check(lib().rabitq_dist_last_status(self.shard._h, C.byref(st)))
        return int(st.value)


class DistributedRaBitQ:
    """`RaBitQ` whose base set and IVF clusters are sharded over the ranks of a torch.distributed group; every rank answers
    its own slice of the batch and gets the reference's exact result for it."""

    def __init__(self, shard: RaBitQ, comm: TorchComm | None = None, records_per_query: int = 256):
        self.shard, self.comm = shard, comm or TorchComm()
        self.rpq = records_per_query
        self.overlap = os.environ.get("RABITQ_DIST_OVERLAP", "1") != "0"
        # the [q | y] exchange: pushed into the peers' inboxes by the copy engines (default), or an NCCL all-gather (A/B switch)
        self.push = os.environ.get("RABITQ_DIST_PUSH", "1") != "0"
        self._st: _RankState | None = None

    @classmethod
    def load_from_dir(cls, path, device: int, group=None, **kw) -> "DistributedRaBitQ":
        comm = TorchComm(group)
        return cls(RaBitQ.load_from_dir(path, device, comm.rank, comm.world), comm, **kw)

    @classmethod
    def from_arrays(cls, *arrays, device: int, group=None, **kw) -> "DistributedRaBitQ":
        comm = TorchComm(group)
        return cls(RaBitQ.from_arrays(*arrays, device=device, shard_rank=comm.rank, shard_count=comm.world), comm, **kw)

    def _state(self, nq_local: int, length: int, probe: int, topk: int) -> _RankState:
        st = self._st
        if st and (st.nq_local, st.len, st.probe, st.topk, st.rpq) == (nq_local, length, probe, topk, self.rpq):
            return st
        import torch

        torch.cuda.synchronize(self.shard.device)
        if st is not None:
            # re-creating the inboxes: every rank unmaps its peers' inboxes FIRST, and only after a barrier does anybody free its
            # own (CUDA IPC: an exported allocation must not be freed while a peer still has it open)
            _check(lib().rabitq_dist_close_peers(self.shard._h))
            self.comm.exchange_bytes(b"")
        st = _RankState(self.shard, self.comm.rank, self.comm.world, nq_local, length, probe, topk, self.rpq, push=self.push)
        handles = self.comm.exchange_bytes(st.ipc_handle())  # also a barrier: nobody writes into a freed inbox
        for r, h in enumerate(handles):
            if r != self.comm.rank:
                _check(lib().rabitq_dist_set_peer(self.shard._h, r, h, None))
        self._st = st
        return st

    def query_batch(self, queries_local, probe: int, topk: int, max_retries: int = 4):
        """queries_local: CUDA tensor [nq_local, len] -- THIS rank's slice (same nq_local on every rank).  Returns
        (dist [nq_local, topk], ids, count) for it, on torch's current stream (which must be the shard's stream)."""
        import torch

        nq_local, length = queries_local.shape
        cur = torch.cuda.current_stream(queries_local.device).cuda_stream
        if cur == 0:
            raise RabitqError(2, "DistributedRaBitQ needs a non-default CUDA stream (torch.cuda.set_stream): the phases and the "
                                 "collectives are ordered by ONE stream, and the legacy default stream cannot be handed to the library")
        self.shard.set_stream(cur)
        for _ in range(max_retries + 1):
            st = self._state(nq_local, length, probe, topk)
            st.front_rotate(queries_local)
            if st.push:    # [q | y] is on its way into every inbox (copy engines); front_select orders this stream behind the pushes
                st.front_select()
                self.comm.all_gather(st.gathered_meta, st.send_meta)
            elif self.overlap:
                work = self.comm.all_gather_start(st.gathered_qy, st.send_qy)   # 7.7 KB per query: in flight during K2 / K2b
                st.front_select()
                self.comm.all_gather(st.gathered_meta, st.send_meta)
                work.wait()
            else:  # (A/B switch: RABITQ_DIST_OVERLAP=0)
                self.comm.all_gather(st.gathered_qy, st.send_qy)
                st.front_select()
                self.comm.all_gather(st.gathered_meta, st.send_meta)
            st.round1()
            self.comm.all_reduce_min(st.thr)
            st.round2()
            self.comm.all_reduce_max(st.status)   # barrier: every shard's records are in the inboxes + global overflow verdict
            status = st.finish()                  # (a home sets bit 2 only if some source set bit 1, which everybody already knows)
            if status == 0:
                return st.out_d, st.out_i, st.out_c
            self.rpq *= 4  # an inbox region overflowed somewhere: every rank grows and repeats the step
        raise RabitqError(5, "survivor-record regions still overflow after growing; raise records_per_query")


def run_virtual_ranks(shards: list[RaBitQ], queries, probe: int, topk: int, records_per_query: int = 256, states=None, grow: bool = True,
                      push: bool = True):
    """The distributed step for `len(shards)` ranks living in THIS process on one GPU: same phases, same kernels, the
    collectives replaced by device copies.  queries: CUDA tensor [world * nq_local, len].  Returns (dist, ids, count, states)."""
    import torch

    world = len(shards)
    nq, length = queries.shape
    assert nq % world == 0
    nq_l = nq // world
    dev = queries.device
    if torch.cuda.current_stream(dev).cuda_stream == 0:  # the legacy default stream cannot be handed to the library: use a side stream
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            out = run_virtual_ranks(shards, queries, probe, topk, records_per_query, states, grow, push)
        torch.cuda.current_stream(dev).wait_stream(side)
        return out
    stream = torch.cuda.current_stream(dev).cuda_stream
    for s in shards:
        s.set_stream(stream)
    while True:
        if states is None:
            torch.cuda.synchronize(dev)
            states = [_RankState(s, r, world, nq_l, length, probe, topk, records_per_query, push=push) for r, s in enumerate(shards)]
            for a in states:
                for b in states:
                    if a is not b:
                        _check(lib().rabitq_dist_set_peer(a.shard._h, b.rank, None, C.c_void_p(b.inbox_ptr())))
        for r, st in enumerate(states):
            st.front_rotate(queries[r * nq_l:(r + 1) * nq_l])
            st.front_select()
        gathered_meta = torch.cat([st.send_meta for st in states])  # the all-gather(s)
        if not states[0].push:
            gathered_qy = torch.cat([st.send_qy for st in states])
        for st in states:
            if not st.push:
                st.gathered_qy.copy_(gathered_qy)
            st.gathered_meta.copy_(gathered_meta)
            st.round1()
        thr = states[0].thr.clone()                                 # all-reduce(min)
        for st in states[1:]:
            _check(lib().rabitq_min_f32_device(dev.index, C.c_void_p(thr.data_ptr()), C.c_void_p(st.thr.data_ptr()), thr.numel(), C.c_void_p(stream)))
        for st in states:
            st.thr.copy_(thr)
            st.round2()
        status = max(st.finish() for st in states)                  # (stream order = the barrier)
        if status == 0:
            return (torch.cat([st.out_d for st in states]), torch.cat([st.out_i for st in states]), torch.cat([st.out_c for st in states]), states)
        if not grow or records_per_query > (1 << 22):
            raise RabitqError(5, f"survivor-record regions overflowed (status {status}); raise records_per_query")
        records_per_query *= 4   # like DistributedRaBitQ.query_batch: every rank grows its inbox and the step is repeated
        states = None
