"""CPU model of the distributed rerank: the source-side sequential filter must ship a superset of what the
reference's sequential reranker computes on every shard, so that the home replay reproduces HeapReRanker::rank_batch
(/root/reference/src/rerank.rs:81-106) exactly -- result set AND `precise` counter -- also when the RaBitQ bound fails
(exact < rough).  Pure Python on random (rough, exact) streams; no GPU, no library.  The CUDA path is checked against the oracle by
tests/test_gpu_distributed.py; this test pins the ARGUMENT the kernel relies on (rerank_cta_kernel<.., SINK = 2>)."""
import heapq
import random

import pytest

INF = float("inf")


class Heap:
    """The reference's state: max-heap of the k best exact distances, threshold = its maximum once it holds k."""

    def __init__(self, k, thr=INF):
        self.k, self.h, self.thr, self.precise = k, [], thr, 0

    def offer(self, rough, exact):
        """rerank.rs:84-101: returns True if the candidate's exact distance is computed."""
        if not rough < self.thr:
            return False
        self.precise += 1
        if exact < self.thr:
            heapq.heappush(self.h, -exact)
            if len(self.h) > self.k:
                heapq.heappop(self.h)
            if len(self.h) == self.k:
                self.thr = -self.h[0]
        return True

    def result(self):
        return sorted(-x for x in self.h)


def make_stream(rng, n_ranks, k, fail_rate):
    """Candidates in visit order: (probe rank, rough, exact); exact >= rough except for bound failures."""
    out = []
    for p in range(n_ranks):
        for _ in range(rng.randint(0, 40)):
            exact = rng.uniform(0.0, 100.0) * (0.3 + p / n_ranks)
            rough = exact - abs(rng.gauss(0.0, 6.0))
            if rng.random() < fail_rate:
                rough = exact + abs(rng.gauss(0.0, 8.0))  # the estimate's lower bound failed
            if rng.random() < 0.05:
                exact = round(exact)  # exact ties
            out.append((p, rough, exact))
    return out


def reference(stream, k):
    h = Heap(k)
    computed = [h.offer(r, e) for _, r, e in stream]
    return h.result(), h.precise, computed


def distributed(stream, k, world, first, key_is_max=True):
    """Round 1 = the first `first` candidates (their owner replays them exactly and freezes the threshold); frozen round = every
    shard filters its own candidates with min(frozen, local k-th smallest key); home = exact replay over everything shipped."""
    owner = lambda p: p % world
    r1 = Heap(k)
    shipped = [False] * len(stream)
    for i in range(min(first, len(stream))):
        shipped[i] = r1.offer(stream[i][1], stream[i][2])
    frozen = r1.thr
    for s in range(world):
        keys, thr = [], frozen  # local state of shard s: nothing of round 1, nothing of the other shards
        for i in range(first, len(stream)):
            p, rough, exact = stream[i]
            if owner(p) != s or not rough < thr:
                continue
            shipped[i] = True
            key = max(exact, rough) if key_is_max else exact
            if key < thr:
                heapq.heappush(keys, -key)
                if len(keys) > k:
                    heapq.heappop(keys)
                if len(keys) == k:
                    thr = min(frozen, -keys[0])
    home = Heap(k)
    for i, (_, rough, exact) in enumerate(stream):
        if shipped[i]:
            home.offer(rough, exact)
    return home.result(), home.precise, shipped


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("k", [1, 10])
def test_local_threshold_ships_a_superset_and_home_replay_is_exact(world, k):
    rng = random.Random(1000 * world + k)
    for trial in range(300):
        stream = make_stream(rng, rng.randint(1, 24), k, fail_rate=rng.choice([0.0, 0.005, 0.05, 0.3]))
        first = rng.randint(0, 12)
        ref_res, ref_precise, ref_computed = reference(stream, k)
        res, precise, shipped = distributed(stream, k, world, first)
        assert all(s or not c for s, c in zip(shipped, ref_computed)), f"trial {trial}: a candidate the reference computes was not shipped"
        assert res == ref_res and precise == ref_precise, f"trial {trial}"


def test_exact_distance_alone_is_not_a_safe_local_key():
    """Why the key is max(exact, rough): with bound failures a shard's own exact distances can push its threshold BELOW the
    reference's, and a candidate the reference reranks is lost.  (Documents the counter-example class; found by search.)"""
    rng = random.Random(7)
    for _ in range(4000):
        stream = make_stream(rng, rng.randint(2, 12), 2, fail_rate=0.4)
        ref_res, ref_precise, ref_computed = reference(stream, 2)
        res, precise, shipped = distributed(stream, 2, 2, 0, key_is_max=False)
        if any(c and not s for s, c in zip(shipped, ref_computed)):
            return
    pytest.fail("no counter-example found: the generator no longer produces bound failures that matter")


def test_stale_thresholds_only_widen_the_queue():
    """The single-GPU replay (rerank_cta_kernel): the scan filters a round with the threshold FROZEN at the round's start, the producer
    filters with a threshold that is some waves old, the compute warps skip with another stale copy -- every one of them is >= the
    live threshold (it only falls), so each stage passes a superset on to the next and the replay warp's exact tests see every
    candidate the reference computes.  Model: any non-increasing sequence of stale thresholds."""
    rng = random.Random(99)
    for trial in range(400):
        k = rng.choice([1, 5, 10])
        stream = make_stream(rng, rng.randint(1, 20), k, fail_rate=rng.choice([0.0, 0.01, 0.2]))
        ref_res, ref_precise, ref_computed = reference(stream, k)
        bounds = sorted(rng.sample(range(len(stream) + 1), min(len(stream) + 1, rng.randint(1, 3))))  # round starts
        live = Heap(k)
        history = [INF]  # the live threshold after every candidate
        frozen = INF
        queued = []
        for i, (_, rough, exact) in enumerate(stream):
            if i in bounds:
                frozen = history[-1]                                 # K4 of the round: the threshold the previous round left
            lag = rng.randint(0, 16)
            stale = history[max(0, len(history) - 1 - lag)]          # what the producer / compute warps happen to read
            passed = rough < frozen and rough < stale
            queued.append(passed)
            if passed:
                live.offer(rough, exact)                             # the replay warp: exact tests on the live threshold
            history.append(live.thr)
        assert all(qd or not c for qd, c in zip(queued, ref_computed)), f"trial {trial}"
        assert live.result() == ref_res and live.precise == ref_precise, f"trial {trial}"
