"""-m gpu: the distributed pipeline (index sharded by cluster range, every rank home of a slice of the batch) against the
UNSHARDED CPU oracle.  The bar is the single-GPU one: top-k distance multisets bit-identical, ids identical up to exact
ties, `rough` and `precise` counters equal -- for any number of shards.

`run_virtual_ranks` runs the ranks inside one process on one GPU (same C-ABI phases and kernels, collectives = device
copies), so these tests need a single GPU.  The NCCL + CUDA-IPC plumbing itself is exercised by
`test_two_process_nccl_ipc`, which needs two GPUs and is skipped otherwise.
"""
import os
import socket

import numpy as np
import pytest

from tests.test_gpu_parity import _same_up_to_ties

pytestmark = pytest.mark.gpu

CASES = ["case_d128", "case_d96", "case_d960"]


def _shards(case, world):
    import rabitq_b200 as rb

    a = case["arrays"]
    return [rb.RaBitQ.from_arrays(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"],
                                  a["factors"], device=0, shard_rank=r, shard_count=world) for r in range(world)]


@pytest.fixture(params=CASES)
def case(request):
    return request.getfixturevalue(request.param)


@pytest.mark.parametrize("r2_seq,sink_mode", [(1, 1), (0, 0)])   # (source-side sequential filter, CTA form of K5) / (flat round 2, warp form)
@pytest.mark.parametrize("world,probe,topk", [(2, 8, 10), (3, 32, 10), (4, 64, 1), (8, 24, 100), (5, 1, 10)])
def test_virtual_ranks_identical_to_unsharded_oracle(case, world, probe, topk, r2_seq, sink_mode):
    import torch

    from rabitq_b200 import distributed as rd

    q = case["queries"]
    nq = q.shape[0] // world * world
    q = np.ascontiguousarray(q[:nq])
    shards = _shards(case, world)
    try:
        for s in shards:
            s.set_option("dist_r2_seq", r2_seq)
            s.set_option("rerank_mode", sink_mode)
        qd = torch.from_numpy(q).cuda()
        _, _, _, st = rd.run_virtual_ranks(shards, qd, probe, topk)   # sizes the inboxes (a too-small region grows and repeats)
        for s in shards:
            s.metrics_reset()
        d, i, c, st = rd.run_virtual_ranks(shards, qd, probe, topk, states=st)
        # a second step on the same state: inbox reuse, counters, stale records
        d2, i2, c2, _ = rd.run_virtual_ranks(shards, qd, probe, topk, states=st)
        torch.cuda.synchronize()
        assert torch.equal(d, d2) and torch.equal(i, i2) and torch.equal(c, c2)
        gd, gi, gc = d.cpu().numpy(), i.cpu().numpy().view(np.uint32), c.cpu().numpy().view(np.uint32)
        o = case["oracle"].query_batch(q, probe, topk)
        assert np.array_equal(gc, o["count"])
        for qi in range(nq):
            n = int(gc[qi])
            assert _same_up_to_ties(case, qi, gd[qi, :n], gi[qi, :n], o["dist"][qi, :n], o["ids"][qi, :n]), f"query {qi}"
            assert np.all(np.diff(gd[qi, :n]) >= 0)
        m = [s.metrics() for s in shards]
        assert sum(x["query"] for x in m) == 2 * nq
        assert sum(x["rough"] for x in m) == 2 * o["rough"]
        assert sum(x["precise"] for x in m) == 2 * o["precise"]      # the reference's counter, not the shards' exact computations
    finally:
        for s in shards:
            s.close()


def test_combined_chunk_entries_equal_split_ones(case_d128):
    """`rabitq_dist_front` + `rabitq_dist_round1` (ONE chunk per rank, one all-gather) give the same step as the split entries
    (`front_rotate` / `front_select` / `round1_split`) that the product path overlaps with its all-gathers."""
    import ctypes as C

    import torch

    from rabitq_b200 import distributed as rd
    from rabitq_b200 import _check, lib

    world, probe, topk = 2, 16, 10
    q = np.ascontiguousarray(case_d128["queries"][: case_d128["queries"].shape[0] // world * world])
    shards = _shards(case_d128, world)
    try:
        qd = torch.from_numpy(q).cuda()
        d0, i0, c0, states = rd.run_virtual_ranks(shards, qd, probe, topk)
        side = torch.cuda.Stream(qd.device)
        with torch.cuda.stream(side):
            for s in shards:
                s.set_stream(side.cuda_stream)
            nq_l, length = q.shape[0] // world, q.shape[1]
            words = int(lib().rabitq_dist_chunk_words(shards[0]._h, length))
            assert words == int(lib().rabitq_dist_chunk_words_qy(shards[0]._h, length)) + int(lib().rabitq_dist_chunk_words_meta(shards[0]._h, length))
            sends = [torch.empty(words, dtype=torch.int32, device=qd.device) for _ in range(world)]
            for r, st in enumerate(states):
                _check(lib().rabitq_dist_front(st.shard._h, C.c_void_p(qd[r * nq_l:(r + 1) * nq_l].data_ptr()), length, C.c_void_p(sends[r].data_ptr())))
            gathered = torch.cat(sends)
            for st in states:
                _check(lib().rabitq_dist_round1(st.shard._h, C.c_void_p(gathered.data_ptr()), C.c_void_p(st.thr.data_ptr())))
            thr = states[0].thr.clone()
            for st in states[1:]:
                _check(lib().rabitq_min_f32_device(qd.device.index, C.c_void_p(thr.data_ptr()), C.c_void_p(st.thr.data_ptr()), thr.numel(),
                                                   C.c_void_p(side.cuda_stream)))
            for st in states:
                st.thr.copy_(thr)
                st.round2()
            assert max(st.finish() for st in states) == 0
            d1 = torch.cat([st.out_d for st in states])
            i1 = torch.cat([st.out_i for st in states])
            c1 = torch.cat([st.out_c for st in states])
            side.synchronize()
        assert torch.equal(d0, d1) and torch.equal(i0, i1) and torch.equal(c0, c1)
    finally:
        for s in shards:
            s.close()


def test_virtual_ranks_region_overflow_is_reported(case_d128):
    """A survivor-record region that is too small must raise, never truncate silently."""
    import torch

    import rabitq_b200 as rb
    from rabitq_b200 import distributed as rd

    q = np.ascontiguousarray(case_d128["queries"][:64])
    shards = _shards(case_d128, 2)
    try:
        # topk = 100 with a 128-vector first round leaves a loose frozen threshold: far more than 8 records per query
        with pytest.raises(rb.RabitqError):
            rd.run_virtual_ranks(shards, torch.from_numpy(q).cuda(), 64, 100, records_per_query=8, grow=False)
    finally:
        for s in shards:
            s.close()


def test_dist_api_argument_errors(case_d128):
    import ctypes as C

    import rabitq_b200 as rb

    s = _shards(case_d128, 2)[0]
    try:
        n = C.c_size_t(0)
        L = rb.lib()
        assert L.rabitq_dist_init(s._h, 1, 2, 8, 8, 10, 64, C.byref(n)) != 0      # rank differs from the handle's shard rank
        assert L.rabitq_dist_init(s._h, 0, 2, 0, 8, 10, 64, C.byref(n)) != 0      # empty slice
        assert L.rabitq_dist_round2(s._h, None) != 0                               # phases out of order
        assert L.rabitq_dist_init(s._h, 0, 2, 8, 8, 10, 64, C.byref(n)) == 0 and n.value > 0
        assert L.rabitq_dist_round1(s._h, None, None) != 0
    finally:
        s.close()


# ---- two processes, NCCL + CUDA IPC (needs two GPUs) ------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, ret):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import oracle as orc
    from rabitq_b200 import distributed as rd
    from tools import synth

    base, queries, cent = synth.make_numpy(20000, 128, 64, 64, "sift", 11)
    o = orc.OracleIndex.from_arrays(base, cent, seed=111, nthreads=4)
    a = o.arrays()
    g = rd.DistributedRaBitQ.from_arrays(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"],
                                         a["factors"], device=rank)
    stream = torch.cuda.Stream(rank)
    torch.cuda.set_stream(stream)
    g.shard.set_stream(stream.cuda_stream)
    nq_l = queries.shape[0] // world
    mine = torch.from_numpy(np.ascontiguousarray(queries[rank * nq_l:(rank + 1) * nq_l])).cuda(rank)
    ok = True
    for _ in range(3):
        d, i, c = g.query_batch(mine, 32, 10)
        torch.cuda.synchronize()
        ref = o.query_batch(queries[rank * nq_l:(rank + 1) * nq_l], 32, 10)
        ok = ok and np.array_equal(np.sort(d.cpu().numpy(), 1).view(np.uint32), np.sort(ref["dist"], 1).view(np.uint32))
    m = g.shard.metrics()
    ret[rank] = (ok, m["precise"], ref["precise"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_nccl_ipc(oracle_lib):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        ok, precise, ref_precise = ret[r]
        assert ok
        assert precise == 3 * ref_precise
