#!/usr/bin/env python
"""bench.py -- QPS of the IVF-RaBitQ query hot path on B200 (BASELINE.json metric), with the roofline of the code scan
and the reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|...] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic queries (nq of the workload; nq * N when N GPUs share
a cluster-sharded index).  `value` is QPS with the queries already resident in HBM, timed with CUDA events on the stream
the kernels are launched on; `e2e` is the same metric through the host-buffer C-ABI call (pinned host queries in,
results out, copies inside the timed region).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "QPS @ recall@10>=0.95 (IVF-RaBitQ query batch)"
TOPK = 10
PROBE_SWEEP = {"c1": [64], "c2": [16, 32, 64, 128, 256], "c3": [16, 32, 64, 128], "c4": [32], "c5": [64]}
TARGET_RECALL = 0.95


# stdout carries exactly ONE JSON line: anything a library prints to fd 1 (NCCL's version banner, for one) goes to stderr.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit_json(obj) -> None:
    os.write(_JSON_FD, (json.dumps(obj) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """Sample SM clocks and throttle reasons with nvidia-smi DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recall_at_k(ids, truth, k):
    import torch

    hit = (ids[:, :, None].to(torch.int64) == truth[:, None, :k].to(torch.int64)).any(2).sum(1)
    return float(hit.float().mean().item()) / k


def workload_string(wl, probe):
    """`config.workload`: the SAME string in both arms (the driver compares it)."""
    return (f"{wl['name']}: {wl['n']}x{wl['dim']} base ({wl['flavour']}-shaped mixture), {wl['nq']} queries, "
            f"{wl['k']} IVF centroids, nprobe={probe}, top-{TOPK}")


def _bcast_chunked(t, src=0, max_elems=1 << 28):
    """NCCL broadcast of a big contiguous tensor in slices (element counts stay below 2^31)."""
    import torch.distributed as dist

    flat = t.view(-1)
    for s in range(0, flat.numel(), max_elems):
        dist.broadcast(flat[s:s + max_elems], src)


def _index_fingerprint(g):
    """Position-weighted integer checksums of every array of an unsharded handle (device side, chunked): two handles hold the same
    index iff the fingerprints match."""
    import torch

    a = g.export_arrays(device_tensors=True)
    dev = a["base"].device
    w = (torch.arange(1 << 24, device=dev, dtype=torch.int64) % 65521) + 1
    out = []
    for key in ("base", "orthogonal", "centroids", "offsets", "map_ids", "codes", "factors"):
        flat = a[key].contiguous().view(-1).view(torch.int32)
        acc = torch.zeros((), dtype=torch.int64, device=dev)
        for s in range(0, flat.numel(), 1 << 24):
            c = flat[s:s + (1 << 24)].to(torch.int64)
            acc += (c * w[:c.numel()]).sum() + (s >> 24)
        out.append(acc)
    del a
    torch.cuda.empty_cache()
    return torch.stack(out)


def _same_index_everywhere(g0, device, rank, world, log_fn):
    """world > 1: every rank trains the whole index from the same inputs and keeps its cluster range, which is only a valid sharding
    if the trained copies are bit-identical.  They are checked (fingerprints, all-reduced); ranks whose copy differs from rank 0's
    (not observed up to 10M vectors; at 100M the torch data generator is not bit-reproducible from run to run) adopt rank 0's
    arrays.  Returns (handle, identical_before_fix)."""
    import torch
    import torch.distributed as dist

    import rabitq_b200 as rb

    fp = _index_fingerprint(g0)
    ref = fp.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([int(bool((fp == ref).all()))], device=device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if int(same.item()) == 1:
        return g0, True
    if rank == 0:
        log_fn("[bench] trained index copies differ between ranks: every rank adopts rank 0's arrays")
    keys = ("base", "orthogonal", "centroids", "offsets", "map_ids", "codes", "factors")
    if rank == 0:
        a = g0.export_arrays(device_tensors=True)
    else:
        D, n, k = g0.dim, g0.num_vectors, g0.num_clusters
        g0.close()
        torch.cuda.empty_cache()
        a = dict(dim=D, base=torch.empty((n, D), dtype=torch.float32, device=device),
                 orthogonal=torch.empty((D, D), dtype=torch.float32, device=device),
                 centroids=torch.empty((k, D), dtype=torch.float32, device=device),
                 offsets=torch.empty((k + 1,), dtype=torch.int32, device=device),
                 map_ids=torch.empty((n,), dtype=torch.int32, device=device),
                 codes=torch.empty((n, D // 64), dtype=torch.int64, device=device),
                 factors=torch.empty((n, 4), dtype=torch.float32, device=device))
    for key in keys:
        _bcast_chunked(a[key], 0)
    torch.cuda.synchronize(device)
    if rank != 0:
        g0 = rb.RaBitQ.from_arrays(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"], a["factors"],
                                   device=device.index)
    del a
    torch.cuda.empty_cache()
    return g0, False


def build_workload(args, device, rank, world, builder=None, keep_single=True):
    """Synthetic data + index on the device; returns dict with torch tensors.

    `handle` is the handle the product path runs on (the rank's shard when world > 1); `single` is the UNSHARDED handle
    every rank trained (the parity reference of the distributed path and the source of the oracle's arrays -- arrays are
    never exported from a shard); the caller closes it.  builder="torch": the index comes from the torch harness and no
    handle is made at all (`arrays` holds the built arrays; the reference arm, which must not load librabitq_b200.so)."""
    import torch
    from tools import synth, build_index_torch as bi

    if args.shape:
        n, dim, nq, k = [int(x) for x in args.shape.split(",")[:4]]
        flavour = args.shape.split(",")[4] if len(args.shape.split(",")) > 4 else "sift"
        shape, seed, name = (n, dim, nq, k, flavour), 77, f"custom {args.shape}"
    else:
        shape, seed, name = synth.SHAPES[args.workload], synth.SEEDS[args.workload], args.workload
    n, dim, nq, k, flavour = shape
    if args.nq:
        nq = args.nq
    t0 = time.time()
    mix = synth.Mixture(dim, k, flavour, seed + 2, device)
    base = mix.draw(n, seed)
    queries = mix.draw(nq * world, seed + 1)  # weak scaling: the query batch grows with the number of GPUs
    cent = mix.centroids()
    if world > 1:  # one copy of the synthetic inputs: rank 0's (the generator's matmul / rounding is not bit-reproducible at 100M vectors)
        base, queries, cent = base.contiguous(), queries.contiguous(), cent.contiguous()
        for t in (base, queries, cent):
            _bcast_chunked(t, 0)
    torch.cuda.synchronize()
    t1 = time.time()
    builder = builder or getattr(args, "builder", "native")
    arrays, g0 = None, None
    if builder == "torch-arrays":   # reference arm: torch harness only, arrays go straight to the oracle
        arrays = bi.to_numpy(bi.build_index(base, cent, seed=seed + 3))
    elif builder == "torch":
        import rabitq_b200 as rb

        ix = bi.build_index(base, cent, seed=seed + 3)
        g0 = rb.RaBitQ.from_arrays(ix["dim"], ix["base"], ix["orthogonal"], ix["centroids"], ix["offsets"], ix["map_ids"], ix["codes"],
                                   ix["factors"], device=device.index)
        del ix
    else:  # index training by the library itself (rabitq_build = RaBitQ::from_path on the device)
        import rabitq_b200 as rb

        g0 = rb.RaBitQ.build(base.contiguous(), cent.contiguous(), seed=seed + 3, device=device.index)
    torch.cuda.synchronize()
    t2 = time.time()
    # ground truth (exact fp32 brute force) for the queries recall is measured on: all of them, or the first `truth_queries` of
    # every rank's slice when the batch is large
    tq = min(nq, args.truth_queries) if args.truth_queries else nq
    sel = torch.cat([torch.arange(r * nq, r * nq + tq, device=device) for r in range(world)])
    truth = torch.zeros((nq * world, TOPK), dtype=torch.int32, device=device)
    truth[sel] = synth.brute_force_topk_torch(base, queries[sel], TOPK)
    torch.cuda.synchronize()
    if rank == 0:
        log(f"[bench] {name}: n={n} dim={dim} nq={nq * world} k={k} ({flavour}); gen {t1 - t0:.1f}s build {t2 - t1:.1f}s truth {time.time() - t2:.1f}s")
    del base
    torch.cuda.empty_cache()
    g = g0
    identical = None
    if world > 1 and g0 is not None:  # every rank trained the whole index from the same inputs and keeps its cluster range
        g0, identical = _same_index_everywhere(g0, device, rank, world, log)
        g = g0.reshard(rank, world)
        if not keep_single:
            g0.close()
            g0 = None
    return dict(name=name, n=n, dim=dim, nq=nq * world, nq_rank=nq, truth_queries=tq, k=k, flavour=flavour, queries=queries.contiguous(),
                truth=truth, handle=g, single=g0, arrays=arrays, D=(g.dim if g is not None else arrays["dim"]),
                index_identical_across_ranks=identical)


def oracle_from_index(g=None, arrays=None):
    """The CPU oracle over the arrays of an UNSHARDED device handle (or over arrays built by the torch harness)."""
    from oracle import oracle as orc

    orc.build()
    a = arrays if arrays is not None else g.export_arrays()
    return orc.OracleIndex.from_built(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"], a["factors"])


def compare_results(d_a, i_a, c_a, d_b, i_b, c_b):
    """Parity of two result sets (numpy [n, k] f32 / u32 + counts): distance lists bit-identical as multisets, ids identical
    up to exact-distance ties (an id present on one side only must share its distance with an id present only on the other)."""
    import numpy as np

    n = d_a.shape[0]
    dist_ok, ids_ok, ids_strict, bad = 0, 0, 0, []
    for q in range(n):
        ca, cb = int(c_a[q]), int(c_b[q])
        da, db = np.sort(d_a[q, :ca]).view(np.uint32), np.sort(d_b[q, :cb]).view(np.uint32)
        dq = ca == cb and np.array_equal(da, db)
        sa, sb = set(i_a[q, :ca].tolist()), set(i_b[q, :cb].tolist())
        strict = sa == sb
        ties = strict
        if not strict and dq:
            only_a = sorted(float(d_a[q, j]) for j in range(ca) if int(i_a[q, j]) not in sb)
            only_b = sorted(float(d_b[q, j]) for j in range(cb) if int(i_b[q, j]) not in sa)
            ties = only_a == only_b
        dist_ok += dq; ids_ok += bool(ties); ids_strict += strict
        if not (dq and ties) and len(bad) < 4:
            bad.append(q)
    return {"queries": n, "dist_bit_identical": dist_ok == n, "ids_identical_up_to_ties": ids_ok == n,
            "ids_identical_strict": ids_strict, "mismatching_queries": bad}


def measure_ours(args, workload, device, rank, world, stream, main_leg=True):
    """One workload through the product path: nprobe choice, parity legs, device-resident timing, end-to-end timing."""
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist

    import rabitq_b200 as rb

    local = device.index
    args_w = argparse.Namespace(**vars(args))
    args_w.workload = workload
    if not main_leg:
        args_w.shape, args_w.nq, args_w.probe = None, 0, 0
    wl = build_workload(args_w, device, rank, world)
    queries, truth, g, g0 = wl["queries"], wl["truth"], wl["handle"], wl["single"]
    nq, D = wl["nq"], wl["D"]
    torch.cuda.synchronize(device)
    g.set_stream(stream.cuda_stream)
    if args.rounds:
        g.set_rounds([int(x) for x in args.rounds.split(",")])

    nq_l = nq // world
    q_lo = rank * nq_l
    q_local = queries[q_lo:q_lo + nq_l].contiguous()
    truth_local = truth[q_lo:q_lo + nq_l]
    dg = None
    if world > 1:
        from rabitq_b200 import distributed as rd

        # every rank is the home of its own slice of the batch; results identical to the single-process reference
        dg = rd.DistributedRaBitQ(g, rd.TorchComm(), records_per_query=args.records_per_query)

    def one_pass(probe):
        if world == 1:
            d, i, c = g.query_batch(queries, probe, TOPK)
            return d, i, c
        return dg.query_batch(q_local, probe, TOPK)

    tq = wl["truth_queries"]  # recall is measured on the first tq queries of every rank's slice

    def global_recall(ids):
        r = recall_at_k(ids[:tq], truth_local[:tq], TOPK)
        if world > 1:
            t = torch.tensor([r], dtype=torch.float64, device=device)
            dist.all_reduce(t)
            r = float(t[0]) / world
        return r

    # ---- choose nprobe: the smallest of the sweep reaching the target recall (outside the timed region) -------------
    sweep = [args_w.probe] if args_w.probe else PROBE_SWEEP.get(workload, [64])
    probe, recall, sweep_log = None, 0.0, {}
    for p in sweep:
        _, ids, _ = one_pass(p)
        r = global_recall(ids)
        sweep_log[str(p)] = round(r, 4)
        probe, recall = p, r
        if r >= TARGET_RECALL:
            break
    target_met = recall >= TARGET_RECALL
    if rank == 0:
        log(f"[bench] {workload}: recall@{TOPK} by nprobe: {sweep_log} -> nprobe={probe}" + ("" if target_met else "  (TARGET MISSED)"))

    # ---- parity legs (outside the timed region) --------------------------------------------------------------------------
    # (1) world > 1: the distributed result of every rank's slice against the UNSHARDED single-GPU handle on the same queries;
    # (2) rank 0: the CPU oracle on a bounded sample of the same queries (that run is also the cpu_baseline timing), against the
    #     single-GPU handle's ids / distances / rough / precise for exactly those queries.
    d_prod, i_prod, c_prod = one_pass(probe)
    d_prod, i_prod, c_prod = d_prod.clone(), i_prod.clone(), c_prod.clone()
    torch.cuda.synchronize(device)
    parity_check = None
    if world > 1 and g0 is not None:
        g0.set_stream(stream.cuda_stream)
        d1, i1, c1 = g0.query_batch(q_local, probe, TOPK)
        torch.cuda.synchronize(device)
        single_precise = g0.last_timings()["precise"]
        cmp_ = compare_results(d_prod.cpu().numpy(), i_prod.cpu().numpy().view(np.uint32), c_prod.cpu().numpy(),
                               d1.cpu().numpy(), i1.cpu().numpy().view(np.uint32), c1.cpu().numpy())
        dist_precise = g.last_timings()["precise"]  # candidates the replay at home computed exact distances for (this rank's queries)
        flags = torch.tensor([int(cmp_["dist_bit_identical"]), int(cmp_["ids_identical_up_to_ties"]), cmp_["ids_identical_strict"],
                              single_precise, dist_precise], dtype=torch.float64, device=device)
        mn = flags.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        sm = flags.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        parity_check = {"against": "the unsharded single-GPU handle on every rank's own slice of the batch",
                        "queries": nq, "dist_bit_identical": bool(mn[0] > 0), "ids_identical_up_to_ties": bool(mn[1] > 0),
                        "ids_identical_strict": int(sm[2]), "precise_single_gpu": int(sm[3]), "precise_distributed": int(sm[4]),
                        "precise_equal": int(sm[3]) == int(sm[4])}
    cpu = None
    if rank == 0 and not args.no_cpu:
        single = g0 if g0 is not None else g
        cpu = cpu_baseline(wl, single, probe, threads=1, budget_s=args.cpu_seconds if main_leg else min(args.cpu_seconds, 6.0),
                           prod=(d_prod, i_prod, c_prod) if world > 1 else None, stream=stream)
    if world > 1 and g0 is not None:
        g0.close()
        wl["single"] = g0 = None
        torch.cuda.empty_cache()
        dist.barrier()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def timed(fn, steps, after=None):
        evs = []
        for _ in range(steps):
            flush.zero_()  # flush L2 between timed iterations (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            if after:
                after()  # bookkeeping of the harness (stage timers, counters): outside the timed region
            evs.append((e0, e1))
        torch.cuda.synchronize(device)
        return [a.elapsed_time(b) for a, b in evs]

    steps = args.steps if main_leg else max(3, min(args.steps, 10))
    # ---- device-resident leg ------------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        one_pass(probe)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stage_ms = {k: 0.0 for k in rb.TIMING_STAGES}
    counts = {k: 0 for k in rb.COUNT_NAMES}

    dev_d = torch.empty((nq_l if world > 1 else nq, TOPK), dtype=torch.float32, device=device)
    dev_i = torch.empty_like(dev_d, dtype=torch.int32)
    dev_c = torch.empty((dev_d.shape[0],), dtype=torch.int32, device=device)

    def step_dev():
        if world == 1:  # caller-owned outputs, the library's stream = torch's current stream: no allocation, no device-wide sync
            g.query_batch_device_into(queries, probe, TOPK, dev_d, dev_i, dev_c)
        else:
            one_pass(probe)

    def collect_dev():
        t = g.last_timings()
        for k in rb.TIMING_STAGES:
            stage_ms[k] += t["ms_" + k]
        for k in rb.COUNT_NAMES:
            counts[k] += t[k]

    if main_leg:
        torch.cuda.cudart().cudaProfilerStart()  # `ncu --profile-from-start off` sees exactly the timed steps (no data generation / index build)
    ms_dev = timed(step_dev, steps, collect_dev)
    torch.cuda.synchronize(device)
    if main_leg:
        torch.cuda.cudart().cudaProfilerStop()
    if world > 1:
        dist.barrier()

    # ---- end-to-end leg: pinned host queries in, host results out, through the host-pointer C-ABI call ---------------
    q_host = torch.empty((nq_l, queries.shape[1]), dtype=torch.float32).pin_memory()
    q_host.copy_(q_local.cpu())
    o_d = torch.empty((nq_l, TOPK), dtype=torch.float32).pin_memory()
    o_i = torch.empty((nq_l, TOPK), dtype=torch.int32).pin_memory()
    o_c = torch.empty((nq_l,), dtype=torch.int32).pin_memory()
    qh, od, oi, oc = q_host.numpy(), o_d.numpy(), o_i.numpy().view(np.uint32), o_c.numpy().view(np.uint32)
    q_stage = torch.empty_like(q_local)
    # two pinned host batches, alternating: while batch i is answered the library uploads batch i+1 on its copy stream
    # (rabitq_query_batch_pipelined); every timed step still contains one H2D of nq x len floats and one D2H of the results
    q_host2 = torch.empty_like(q_host).pin_memory()
    q_host2.copy_(q_host)
    qh2 = q_host2.numpy()
    flip = [0]

    def step_e2e():
        if world == 1:
            cur, nxt = (qh, qh2) if flip[0] == 0 else (qh2, qh)
            flip[0] ^= 1
            g.query_batch_into(cur, probe, TOPK, od, oi, oc, next_q_host=nxt)  # H2D of queries + all kernels + D2H of results inside
            return
        # every rank: its slice of the batch from pinned host memory, the distributed step, its results back to the host
        q_stage.copy_(q_host, non_blocking=True)
        d, i, c = dg.query_batch(q_stage, probe, TOPK)
        o_d.copy_(d, non_blocking=True)
        o_i.copy_(i, non_blocking=True)
        o_c.copy_(c, non_blocking=True)
        stream.synchronize()

    for _ in range(3):
        step_e2e()
    if world > 1:
        dist.barrier()
    ms_e2e = timed(step_e2e, steps)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ids = torch.from_numpy(oi.view(np.int32).copy()).to(device)
    recall_e2e = global_recall(e2e_ids)

    # ---- optional sweeps the configs name (C2: nprobe 16-256; C5: query batch 1-65536); single GPU, after the main legs ----------
    sweeps = {}
    if world == 1 and main_leg and args.sweep_probe:
        rows = []
        for p in [int(x) for x in args.sweep_probe.split(",")]:
            _, ids_p, _ = one_pass(p)
            rec_p = global_recall(ids_p)
            for _ in range(3):
                g.query_batch_device_into(queries, p, TOPK, dev_d, dev_i, dev_c)
            ms = timed(lambda: g.query_batch_device_into(queries, p, TOPK, dev_d, dev_i, dev_c), max(3, steps // 2))
            tl = g.last_timings()
            ms2 = timed(lambda: g.query_batch_into(qh, p, TOPK, od, oi, oc, next_q_host=qh), max(3, steps // 2))
            rows.append({"nprobe": p, "recall_at_10": round(rec_p, 4), "qps": round(nq / (sum(ms) / len(ms) * 1e-3), 1),
                         "qps_e2e": round(nq / (sum(ms2) / len(ms2) * 1e-3), 1), "ms_per_step": round(sum(ms) / len(ms), 4),
                         "scan_ms": round(tl["ms_scan"], 4), "rerank_ms": round(tl["ms_rerank"], 4), "pairs": tl["pairs"],
                         "scan_gpairs_per_s": round(tl["pairs"] / (tl["ms_scan"] * 1e-3) / 1e9, 1) if tl["ms_scan"] > 0 else None})
            log(f"[bench] nprobe sweep: {rows[-1]}")
        sweeps["probe_sweep"] = rows
    if world == 1 and main_leg and args.sweep_batch:
        rows = []
        for nb in [int(x) for x in args.sweep_batch.split(",")]:
            nb = min(nb, nq)
            qb = queries[:nb].contiguous()
            bd, bi, bc = dev_d[:nb], dev_i[:nb], dev_c[:nb]
            qhb, odb, oib, ocb = qh[:nb], od[:nb], oi[:nb], oc[:nb]
            reps = max(3, min(50, 4096 // nb))
            for _ in range(3):
                g.query_batch_device_into(qb, probe, TOPK, bd, bi, bc)
            ms = timed(lambda: g.query_batch_device_into(qb, probe, TOPK, bd, bi, bc), reps)
            t0 = time.perf_counter()
            for _ in range(reps):
                g.query_batch_into(qhb, probe, TOPK, odb, oib, ocb)
            wall = (time.perf_counter() - t0) / reps * 1e3
            rec_b = recall_at_k(bi[:min(nb, tq)], truth_local[:min(nb, tq)], TOPK)
            rows.append({"batch": nb, "ms_device": round(sum(ms) / len(ms), 4), "qps_device": round(nb / (sum(ms) / len(ms) * 1e-3), 1),
                         "ms_host_call": round(wall, 4), "qps_host_call": round(nb / (wall * 1e-3), 1), "recall_at_10": round(rec_b, 4)})
            log(f"[bench] batch sweep: {rows[-1]}")
        sweeps["batch_sweep"] = rows

    tot_dev, tot_e2e = sum(ms_dev), sum(ms_e2e)
    if world > 1:
        t = torch.tensor([tot_dev, tot_e2e], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_dev, tot_e2e = float(t[0]), float(t[1])
        c = torch.tensor([counts["pairs"], counts["survivors"], counts["exact_computed"], counts["precise"], counts["kernel_launches"]],
                         dtype=torch.float64, device=device)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        pairs_all = float(c[0])
        per_rank = torch.zeros((world, 3), dtype=torch.float64, device=device)
        mine = torch.tensor([counts["pairs"] / steps, stage_ms["scan"] / steps, stage_ms["total"] / steps], dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(per_rank.view(-1), mine)
        per_rank = per_rank.cpu().tolist()
    else:
        pairs_all = float(counts["pairs"])
        per_rank = None

    out = None
    if rank == 0:
        bytes_per_pair = D // 8 + 16  # packed code + Factor
        scan_ms = stage_ms["scan"]
        peak, peak_src = measured_peak_hbm()
        achieved = counts["pairs"] * bytes_per_pair / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
        qps = nq * steps / (tot_dev * 1e-3)
        qps_e2e = nq * steps / (tot_e2e * 1e-3)
        prof = scan_profile(workload) if not args_w.shape else {}
        out = {
            "metric": METRIC, "value": round(qps, 1), "unit": "queries/s", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(tot_dev / steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 x u8 -> s32 tensor-core contraction + f32 estimator", "data": "synthetic",
            "recall_target_met": bool(target_met),
            "config": {"workload": workload_string(wl, probe),
                       "nprobe": probe, "topk": TOPK, "recall_at_10": round(recall, 4), "recall_by_nprobe": sweep_log,
                       "recall_target": TARGET_RECALL, "recall_target_met": bool(target_met),
                       "queries_per_step": nq,
                       "recall_measured_on": f"first {tq} queries of every rank's slice",
                       "timing": "CUDA events on the launch stream, L2 flushed (256 MB write) between timed steps",
                       "parallelism": "single GPU" if world == 1 else f"index sharded by cluster range over {world} GPUs, "
                                      f"every rank home of {nq // world} of the {nq} queries; results identical to the single-process "
                                      f"reference (see parity_check)",
                       "rerank_rounds": args.rounds or "0"},
            "e2e": {"value": round(qps_e2e, 1), "unit": "queries/s", "h2d_bytes_per_step": int(nq * queries.shape[1] * 4),  # all ranks
                    "d2h_bytes_per_step": int(nq * TOPK * 8 + nq * 4), "ms_per_step": round(tot_e2e / steps, 4),
                    "recall_at_10": round(recall_e2e, 4)},
            "gpu_launches": int(counts["kernel_launches"]),
            "roofline": {"bound": prof.get("bound", "hbm"), "kernel": prof.get("kernel", "rq::scan_mma_kernel"),
                         "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": prof.get("dram_bytes_per_launch"),
                         "traffic_source": prof.get("source"), "peak_source": peak_src,
                         "pipes_pct": prof.get("pipes_pct"),
                         "algorithmic_bytes_per_pair": bytes_per_pair, "pairs_per_step": counts["pairs"] // steps,
                         "scan_ms_per_step": round(scan_ms / steps, 4), "scan_launches_per_step": counts["scan_launches"] // steps,
                         "gpairs_per_s": round(counts["pairs"] / (scan_ms * 1e-3) / 1e9, 2) if scan_ms > 0 else 0.0,
                         "note": "rank-0 shard; achieved = ALGORITHMIC bytes (pairs x (D/8+16) B) / scan-kernel time measured in this run "
                                 "with CUDA events; cross-query reuse of a cluster's codes keeps real DRAM traffic (`traffic`, from the "
                                 "committed ncu capture) far below it, so frac can exceed what DRAM alone would allow"},
            "stage_ms_per_step": {k: round(v / steps, 4) for k, v in stage_ms.items()},
            "counters_per_step": {"pairs_all_gpus": int(pairs_all // steps), "survivors": counts["survivors"] // steps,
                                  "exact_computed": counts["exact_computed"] // steps, "precise": counts["precise"] // steps},
            "clocks": clocks,
        }
        if per_rank:
            out["per_rank"] = {"pairs": [int(r[0]) for r in per_rank], "scan_ms": [round(r[1], 4) for r in per_rank],
                               "busy_ms": [round(r[2], 4) for r in per_rank]}
        out.update(sweeps)
        if parity_check is not None:
            parity_check["trained_index_identical_across_ranks"] = wl.get("index_identical_across_ranks")
            out["parity_check"] = parity_check
        if cpu is not None:
            out["parity"] = cpu.pop("parity")
            out["cpu_baseline"] = cpu
    g.close()
    del wl, queries, truth, flush
    torch.cuda.empty_cache()
    return out


def scan_profile(workload):
    """ncu-derived facts about the dominant kernel from the committed capture of this round (profiles/scan_profile.json, written
    by tools/ncu_summary.py from an `ncu --set full` report): DRAM bytes per launch and pipe utilisation.  Never a constant in code."""
    p = os.path.join(ROOT, "profiles", "scan_profile.json")
    try:
        return json.load(open(p)).get(workload, {})
    except Exception:
        return {}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import rabitq_b200 as rb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and rank == 0:
        log(f"[bench] WORLD_SIZE={world} but --gpus {args.gpus}: using WORLD_SIZE")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    rb.lib()  # fail loudly if the CUDA library is missing
    stream = torch.cuda.Stream(device)  # a non-default stream shared by torch (events, NCCL) and the library's kernels
    torch.cuda.synchronize(device)
    torch.cuda.set_stream(stream)
    out = measure_ours(args, args.workload, device, rank, world, stream, main_leg=True)
    # the north-star sentence is quoted on C1 (1M x 128, 10k queries, nprobe=64): a short leg of it rides in the same line
    if world == 1 and args.workload != "c1" and not args.shape and not args.no_c1:
        c1 = measure_ours(args, "c1", device, rank, world, stream, main_leg=False)
        if rank == 0:
            out["north_star_c1"] = {k: c1[k] for k in ("value", "unit", "ms_per_step", "steps", "recall_target_met", "config", "e2e",
                                                       "roofline", "stage_ms_per_step", "counters_per_step", "parity", "cpu_baseline")
                                    if k in c1}
    if rank == 0:
        emit_json(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(wl, single, probe, threads, budget_s, prod=None, stream=None):
    """Time the oracle (port of the reference's AVX2 path) on a bounded sample of the same workload and compare its answers
    with the GPU's for exactly those queries (`parity`).  `single`: an UNSHARDED handle (arrays are never exported from a shard)."""
    import numpy as np
    import torch

    o = oracle_from_index(single)
    nq_l = wl["nq_rank"]
    q = wl["queries"][:nq_l].cpu().numpy()  # rank 0's slice
    # calibrate on a few queries, then size the sample for ~budget_s of CPU work
    n0 = min(8 * threads, q.shape[0])
    r0 = o.query_batch(q[:n0], probe, TOPK, nthreads=threads)
    per_q = max(r0["seconds"] / n0, 1e-6)
    ns = int(min(q.shape[0], max(n0, budget_s / per_q)))
    r = o.query_batch(q[:ns], probe, TOPK, nthreads=threads)
    truth = wl["truth"][:ns].cpu().numpy()
    rec = float(np.mean([len(set(r["ids"][i].tolist()) & set(truth[i].tolist())) / TOPK for i in range(ns)]))
    # the GPU's answers for the same ns queries (single-GPU handle: its counters are for exactly this call)
    qs = wl["queries"][:ns].contiguous()
    gd, gi, gc = single.query_batch(qs, probe, TOPK)
    torch.cuda.synchronize()
    t = single.last_timings()
    par = compare_results(gd.cpu().numpy(), gi.cpu().numpy().view(np.uint32), gc.cpu().numpy(), r["dist"], r["ids"], r["count"])
    par.update({"against": "oracle/ (CPU restatement of the reference) on the cpu_baseline sample, same index, same queries",
                "rough_equal": int(t["pairs"]) == int(r["rough"]), "precise_equal": int(t["precise"]) == int(r["precise"]),
                "rough": int(r["rough"]), "precise": int(r["precise"])})
    if prod is not None:  # the distributed product path against the oracle as well
        d, i, c = prod
        dp = compare_results(d[:ns].cpu().numpy(), i[:ns].cpu().numpy().view(np.uint32), c[:ns].cpu().numpy(), r["dist"], r["ids"], r["count"])
        par["distributed_vs_oracle"] = {k: dp[k] for k in ("queries", "dist_bit_identical", "ids_identical_up_to_ties", "ids_identical_strict")}
    o.close()
    return {"value": round(ns / r["seconds"], 2), "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"first {ns} of {q.shape[0]} queries of the same workload, nprobe={probe}, top-{TOPK}, {threads} thread(s); "
                      f"oracle/ (C++ restatement of the reference's AVX2 path; the Rust reference cannot be built here)",
            "recall_at_10": round(rec, 4), "rough": r["rough"], "precise": r["precise"], "parity": par}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port) on all host threads.  The index is
    trained by the torch harness (tools/build_index_torch.py): librabitq_b200.so is never loaded by this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    if not torch.cuda.is_available():
        emit_json({"impl": "reference", "unavailable": "needs the GPU only to generate the synthetic workload and index"})
        return
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    world_q = max(1, args.gpus)  # same query set as our arm at this N (weak scaling: nq x N queries)
    wl = build_workload(args, device, 0, world_q, builder="torch-arrays")
    threads = os.cpu_count() or 1
    import numpy as np

    o = oracle_from_index(arrays=wl["arrays"])
    wl["arrays"] = None
    q = wl["queries"].cpu().numpy()
    # same rule as our arm: the smallest nprobe of the sweep whose recall@10 reaches the target (measured with this arm)
    probe, sweep_log = args.probe, {}
    if not probe:
        ns0 = min(q.shape[0], 512)
        tr0 = wl["truth"][:ns0].cpu().numpy()
        for p in PROBE_SWEEP.get(args.workload, [64]):
            r0 = o.query_batch(q[:ns0], p, TOPK, nthreads=threads)
            rec0 = float(np.mean([len(set(r0["ids"][i].tolist()) & set(tr0[i].tolist())) / TOPK for i in range(ns0)]))
            sweep_log[str(p)] = round(rec0, 4)
            probe = p
            if rec0 >= TARGET_RECALL:
                break
        log(f"[bench/reference] recall@{TOPK} by nprobe on {ns0} queries: {sweep_log} -> nprobe={probe}")
    n0 = min(4 * threads, q.shape[0])
    r0 = o.query_batch(q[:n0], probe, TOPK, nthreads=threads)
    per_q = max(r0["seconds"] / n0, 1e-6)
    total_steps = args.steps + max(args.warmup, 1)
    ns = int(min(q.shape[0], max(threads, (args.cpu_seconds * 4 / total_steps) / per_q)))
    for _ in range(max(args.warmup, 1)):
        o.query_batch(q[:ns], probe, TOPK, nthreads=threads)
    secs = 0.0
    for _ in range(args.steps):
        r = o.query_batch(q[:ns], probe, TOPK, nthreads=threads)
        secs += r["seconds"]
    nt = min(ns, wl["truth_queries"])
    truth = wl["truth"][:nt].cpu().numpy()
    rec = float(np.mean([len(set(r["ids"][i].tolist()) & set(truth[i].tolist())) / TOPK for i in range(nt)]))
    qps = ns * args.steps / secs
    sample = (f"{ns} of {q.shape[0]} queries per step, nprobe={probe}, top-{TOPK}, {threads} host threads over queries; oracle/ port of the "
              f"reference's AVX2 path (Rust toolchain absent, reference not buildable here); index trained by the torch harness")
    out = {"impl": "reference", "metric": METRIC, "value": round(qps, 2), "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": max(args.warmup, 1), "ms_per_step": round(secs / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64 popcount + f32 (AVX2)", "data": "synthetic",
           "recall_target_met": bool(rec >= TARGET_RECALL),
           "config": {"workload": workload_string(wl, probe),
                      "nprobe": probe, "topk": TOPK, "recall_at_10": round(rec, 4), "recall_by_nprobe": sweep_log,
                      "queries_per_step": ns},
           "cpu_baseline": {"value": round(qps, 2), "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": round(qps, 2), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit_json(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--shape", default=None, help="custom n,dim,nq,k[,flavour] (debug)")
    ap.add_argument("--probe", type=int, default=0, help="fix nprobe instead of sweeping to the target recall")
    ap.add_argument("--nq", type=int, default=0, help="queries per step and GPU (default: the workload's)")
    ap.add_argument("--truth-queries", type=int, default=0, help="measure recall on the first N queries only (large batches)")
    ap.add_argument("--rounds", default=None, help="rerank round boundaries, e.g. 0,1,8")
    ap.add_argument("--builder", default="native", choices=["native", "torch"], help="index training: rabitq_build (CUDA) or the torch harness")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sweep-probe", default=None, help="comma list of nprobe values timed after the main legs (config 2: 16,32,64,128,256)")
    ap.add_argument("--sweep-batch", default=None, help="comma list of query-batch sizes timed after the main legs (config 5: 1,8,...,65536)")
    ap.add_argument("--no-c1", action="store_true", help="skip the short north-star (C1) leg that rides in the N=1 line")
    ap.add_argument("--records-per-query", type=int, default=256, help="multi-GPU: survivor-record capacity per (home query, source shard)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of the cpu_baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
