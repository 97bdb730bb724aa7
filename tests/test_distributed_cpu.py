"""CPU, world_size 2, gloo: the host-side multi-GPU logic -- shard geometry from the C ABI, the all-gather layout of
(dist, id) lists, and that per-shard oracle results merged across ranks equal the unsharded oracle result whenever the
RaBitQ bound holds (and are never worse).  The device merge kernel itself is covered by the gpu-marked tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from rabitq_b200 import distributed as rd
    from tools import synth

    base, queries, cent = synth.make_numpy(6000, 64, 16, 24, "sift", 31)
    full = orc.OracleIndex.from_arrays(base, cent, seed=5, nthreads=2)
    a = full.arrays()
    lo, hi = rd.shard_range(a["offsets"], rank, world)
    # every rank derives the same plan
    plan = [rd.shard_range(a["offsets"], r, world) for r in range(world)]
    assert plan[0][0] == 0 and plan[-1][1] == a["base"].shape[0]
    assert all(plan[r][1] == plan[r + 1][0] for r in range(world - 1))
    assert all(x in set(a["offsets"].tolist()) for p in plan for x in p)  # boundaries fall on cluster boundaries
    loc_off = (np.clip(a["offsets"].astype(np.int64), lo, hi) - lo).astype(np.uint32)
    shard = orc.OracleIndex.from_built(a["dim"], a["base"][lo:hi], a["orthogonal"], a["centroids"], loc_off, a["map_ids"][lo:hi],
                                       a["codes"][lo:hi], a["factors"][lo:hi])
    topk, probe = 10, 12
    r = shard.query_batch(queries, probe, topk)
    d = torch.from_numpy(np.where(np.arange(topk)[None, :] < r["count"][:, None], r["dist"], np.inf).astype(np.float32))
    i = torch.from_numpy(np.where(np.arange(topk)[None, :] < r["count"][:, None], r["ids"], 0xFFFFFFFF).astype(np.int64))
    gd, gi = rd.all_gather_topk(d, i)
    assert gd.shape == (world, queries.shape[0], topk)
    assert torch.equal(gd[rank], d) and torch.equal(gi[rank], i)
    # merge (test-side numpy; the product merge is the CUDA kernel)
    cat_d = gd.permute(1, 0, 2).reshape(queries.shape[0], -1).numpy()
    cat_i = gi.permute(1, 0, 2).reshape(queries.shape[0], -1).numpy()
    order = np.argsort(cat_d, axis=1, kind="stable")[:, :topk]
    md = np.take_along_axis(cat_d, order, 1)
    ref = full.query_batch(queries, probe, topk)
    rd_sorted = np.sort(ref["dist"], axis=1)
    assert np.all(md <= rd_sorted)  # looser shard-local thresholds can only find equal or better neighbours
    same = float(np.mean(np.all(md == rd_sorted, axis=1)))
    if rank == 0:
        ret["same"] = same
        ret["rough_total"] = 0
    # the collectives of the distributed step as DistributedRaBitQ issues them: the asynchronous all-gather of the big part
    # overlapping "work", a second synchronous one behind it, min / max all-reduces
    comm = rd.TorchComm()
    assert (comm.rank, comm.world) == (rank, world)
    big_in = torch.full((1024,), float(rank), dtype=torch.float32)
    big_out = torch.empty(1024 * world, dtype=torch.float32)
    small_in = torch.full((8,), float(10 + rank), dtype=torch.float32)
    small_out = torch.empty(8 * world, dtype=torch.float32)
    work = comm.all_gather_start(big_out, big_in)
    small_in += 0  # (the caller keeps computing here)
    comm.all_gather(small_out, small_in)
    work.wait()
    for rr in range(world):
        assert torch.all(big_out[rr * 1024:(rr + 1) * 1024] == float(rr)) and torch.all(small_out[rr * 8:(rr + 1) * 8] == float(10 + rr))
    thr = torch.tensor([3.0 + rank, 7.0 - rank], dtype=torch.float32)
    comm.all_reduce_min(thr)
    assert thr.tolist() == [3.0, 7.0 - (world - 1)]
    status = torch.tensor([rank], dtype=torch.int32)
    comm.all_reduce_max(status)
    assert int(status[0]) == world - 1
    handles = comm.exchange_bytes(bytes([rank]) * 64)
    assert [h[0] for h in handles] == list(range(world))
    t = torch.tensor([r["rough"]], dtype=torch.int64)
    dist.all_reduce(t)
    if rank == 0:
        ret["rough_total"] = int(t[0])
        ret["rough_ref"] = ref["rough"]
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_merge(oracle_lib):
    import rabitq_b200
    from rabitq_b200 import build

    build.build()
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret["same"] >= 0.9
    assert ret["rough_total"] == ret["rough_ref"]  # every probed cluster is scanned by exactly one shard


def test_shard_range_edge_cases():
    from rabitq_b200 import distributed as rd

    off = np.array([0, 0, 5, 5, 9, 20, 20], np.uint32)  # empty clusters at both ends and in the middle
    for world in (1, 2, 3, 7):
        rows = [rd.shard_range(off, r, world) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == 20
        assert all(rows[r][1] == rows[r + 1][0] for r in range(world - 1))
        assert all(lo <= hi for lo, hi in rows)


def _bcast_worker(rank, world, port, ret):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench

    # every rank "generates" slightly different inputs (what happened at 100M vectors); after the chunked broadcast all hold rank 0's
    t = torch.arange(1000, dtype=torch.float32).reshape(100, 10) + float(rank)
    bench._bcast_chunked(t, 0, max_elems=96)  # forces several slices, the last one partial
    ret[rank] = bool(torch.equal(t, torch.arange(1000, dtype=torch.float32).reshape(100, 10)))
    dist.destroy_process_group()


def test_bench_broadcasts_identical_inputs_to_every_rank():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bcast_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
