"""rabitq_b200 -- B200-native IVF-RaBitQ query path behind the reference's `RaBitQ` interface.

Host-side mirror (Python, over the C ABI in include/rabitq_b200.h) of the one path of kemingy/rabitq this
repo replaces: `RaBitQ::load_from_dir` (src/rabitq.rs:84-125) and `RaBitQ::query` (src/rabitq.rs:268-333),
plus `METRICS.to_str()` (src/metrics.rs:30-41).  All compute runs in hand-written sm_100a CUDA kernels inside
librabitq_b200.so; there is no CPU path here, and importing without the built library fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librabitq_b200.so")

c_f32p = C.POINTER(C.c_float)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)

#: every symbol include/rabitq_b200.h declares
ABI_SYMBOLS = [
    "rabitq_load_from_dir", "rabitq_load_from_dir_sharded", "rabitq_from_arrays", "rabitq_from_path", "rabitq_build",
    "rabitq_dump_to_dir", "rabitq_export_arrays", "rabitq_free", "rabitq_dim",
    "rabitq_num_vectors", "rabitq_num_clusters", "rabitq_query", "rabitq_query_batch", "rabitq_query_batch_pipelined", "rabitq_query_batch_device",
    "rabitq_shard_range", "rabitq_merge_topk_device", "rabitq_metrics", "rabitq_metrics_reset", "rabitq_last_error", "rabitq_set_rounds",
    "rabitq_set_option", "rabitq_set_stream", "rabitq_last_timings", "rabitq_stage_rotate", "rabitq_stage_probe", "rabitq_stage_quantize", "rabitq_stage_scan",
    "rabitq_dist_init", "rabitq_dist_ipc_handle", "rabitq_dist_inbox_ptr", "rabitq_dist_set_peer", "rabitq_dist_close_peers", "rabitq_dist_chunk_words", "rabitq_dist_front",
    "rabitq_dist_round1", "rabitq_dist_round2", "rabitq_dist_finish", "rabitq_min_f32_device", "rabitq_set_quantize_bias", "rabitq_reshard", "rabitq_debug_rerank_stats", "rabitq_debug_rec_pos", "rabitq_dist_last_status", "rabitq_dist_chunk_words_qy", "rabitq_dist_chunk_words_meta",
    "rabitq_dist_front_rotate", "rabitq_dist_front_select", "rabitq_dist_round1_split",
]

TIMING_STAGES = ["h2d_pad", "rotate", "centroid_dist", "select", "quantize", "bucket", "scan", "rerank", "d2h", "total"]
COUNT_NAMES = ["pairs", "survivors", "exact_computed", "precise", "scan_launches", "kernel_launches"]

_lib = None


class RabitqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[rabitq_b200 error {code}] {msg}")
        self.code = code


def lib():
    """Load librabitq_b200.so.  Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m rabitq_b200.build` (nvcc, sm_100a). "
            "rabitq_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.rabitq_last_error.restype = C.c_char_p
    L.rabitq_load_from_dir.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.rabitq_load_from_dir_sharded.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.rabitq_from_arrays.argtypes = [C.c_uint32, C.c_size_t, C.c_size_t, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.POINTER(vp)]
    L.rabitq_from_path.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_int, C.POINTER(vp)]
    L.rabitq_build.argtypes = [vp, C.c_size_t, C.c_size_t, vp, C.c_size_t, vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(vp)]
    L.rabitq_dump_to_dir.argtypes = [vp, C.c_char_p]
    L.rabitq_reshard.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.rabitq_export_arrays.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    L.rabitq_free.argtypes = [vp]
    L.rabitq_free.restype = None
    L.rabitq_dim.argtypes = [vp]
    L.rabitq_dim.restype = C.c_uint32
    L.rabitq_num_vectors.argtypes = [vp]
    L.rabitq_num_vectors.restype = C.c_size_t
    L.rabitq_num_clusters.argtypes = [vp]
    L.rabitq_num_clusters.restype = C.c_size_t
    L.rabitq_query.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp, vp, vp]
    L.rabitq_query_batch.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp, vp, vp]
    L.rabitq_query_batch_device.argtypes = L.rabitq_query_batch.argtypes
    L.rabitq_query_batch_pipelined.argtypes = [vp, vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp, vp, vp]
    L.rabitq_merge_topk_device.argtypes = [C.c_int, vp, vp, C.c_int, C.c_size_t, C.c_size_t, vp, vp, vp, vp]
    L.rabitq_dist_init.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]
    L.rabitq_dist_ipc_handle.argtypes = [vp, C.c_char_p]
    L.rabitq_dist_inbox_ptr.argtypes = [vp, C.POINTER(vp)]
    L.rabitq_dist_set_peer.argtypes = [vp, C.c_int, C.c_char_p, vp]
    L.rabitq_dist_close_peers.argtypes = [vp]
    L.rabitq_dist_chunk_words.argtypes = [vp, C.c_size_t]
    L.rabitq_dist_chunk_words.restype = C.c_size_t
    for f in (L.rabitq_dist_chunk_words_qy, L.rabitq_dist_chunk_words_meta):
        f.argtypes = [vp, C.c_size_t]
        f.restype = C.c_size_t
    L.rabitq_dist_front_rotate.argtypes = [vp, vp, C.c_size_t, vp]
    L.rabitq_dist_front_select.argtypes = [vp, vp]
    L.rabitq_dist_round1_split.argtypes = [vp, vp, vp, vp]
    L.rabitq_dist_front.argtypes = [vp, vp, C.c_size_t, vp]
    L.rabitq_dist_round1.argtypes = [vp, vp, vp]
    L.rabitq_dist_round2.argtypes = [vp, vp]
    L.rabitq_dist_finish.argtypes = [vp, vp, vp, vp, vp]
    L.rabitq_min_f32_device.argtypes = [C.c_int, vp, vp, C.c_size_t, vp]
    L.rabitq_shard_range.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.rabitq_metrics.argtypes = [vp, c_u64p]
    L.rabitq_metrics.restype = None
    L.rabitq_metrics_reset.argtypes = [vp]
    L.rabitq_metrics_reset.restype = None
    L.rabitq_set_rounds.argtypes = [vp, c_u32p, C.c_int]
    L.rabitq_last_timings.argtypes = [vp, c_f32p, c_u64p]
    L.rabitq_set_stream.argtypes = [vp, vp]
    L.rabitq_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.rabitq_set_quantize_bias.argtypes = [vp, vp]
    L.rabitq_debug_rerank_stats.argtypes = [vp, vp, C.c_size_t]
    L.rabitq_debug_rec_pos.argtypes = [C.c_int]
    L.rabitq_stage_rotate.argtypes = [vp, vp, C.c_size_t, C.c_size_t, vp]
    L.rabitq_stage_probe.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp, vp, vp]
    L.rabitq_stage_quantize.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp, vp, vp, vp]
    L.rabitq_stage_scan.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp, vp, vp]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise RabitqError(rc, lib().rabitq_last_error().decode(errors="replace"))


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class RaBitQ:
    """`pub struct RaBitQ` (src/rabitq.rs:57-68) resident in the HBM of one B200 (or one shard of it)."""

    def __init__(self, handle: C.c_void_p, device: int):
        self._h = handle
        self.device = device

    # ---- constructors ------------------------------------------------------------------------------------------
    @classmethod
    def load_from_dir(cls, path, device: int = 0, shard_rank: int = 0, shard_count: int = 1) -> "RaBitQ":
        """`RaBitQ::load_from_dir(path)` (src/rabitq.rs:84-125)."""
        h = C.c_void_p()
        _check(lib().rabitq_load_from_dir_sharded(os.fsencode(str(path)), device, shard_rank, shard_count, C.byref(h)))
        return cls(h, device)

    @classmethod
    def from_arrays(cls, dim, base, orthogonal, centroids, offsets, map_ids, codes, factors, device: int = 0,
                    shard_rank: int = 0, shard_count: int = 1) -> "RaBitQ":
        """Adopt built arrays (numpy on the host, or torch CUDA tensors already on `device`)."""
        on_dev = hasattr(base, "data_ptr")
        if on_dev:
            import torch

            def prep(t, dt):
                assert t.is_cuda and t.device.index == device and t.dtype == dt and t.is_contiguous(), (t.dtype, t.device)
                return C.c_void_p(t.data_ptr())

            n, k = base.shape[0], centroids.shape[0]
            u64 = torch.int64 if codes.dtype == torch.int64 else torch.uint64
            ptrs = [prep(base, torch.float32), prep(orthogonal, torch.float32), prep(centroids, torch.float32),
                    prep(offsets, torch.int32), prep(map_ids, torch.int32), prep(codes, u64), prep(factors, torch.float32)]
            torch.cuda.synchronize(device)
        else:
            base = _np(base, np.float32); orthogonal = _np(orthogonal, np.float32); centroids = _np(centroids, np.float32)
            offsets = _np(offsets, np.uint32); map_ids = _np(map_ids, np.uint32); codes = _np(codes, np.uint64)
            factors = _np(factors, np.float32)
            n, k = base.shape[0], centroids.shape[0]
            ptrs = [_ptr(a) for a in (base, orthogonal, centroids, offsets, map_ids, codes, factors)]
        h = C.c_void_p()
        _check(lib().rabitq_from_arrays(dim, n, k, *ptrs, int(on_dev), device, shard_rank, shard_count, C.byref(h)))
        return cls(h, device)

    @classmethod
    def from_path(cls, base_path, centroid_path, seed: int = 1, device: int = 0) -> "RaBitQ":
        """`RaBitQ::from_path(base_path, centroid_path)` (src/rabitq.rs:159-265): index training, on the device."""
        h = C.c_void_p()
        _check(lib().rabitq_from_path(os.fsencode(str(base_path)), os.fsencode(str(centroid_path)), seed, device, C.byref(h)))
        return cls(h, device)

    @classmethod
    def build(cls, base, centroids, orthogonal=None, seed: int = 1, device: int = 0) -> "RaBitQ":
        """from_path on in-memory arrays: base [n, len], centroids [k, len] (numpy, or torch CUDA tensors on `device`)."""
        on_dev = hasattr(base, "data_ptr")
        if on_dev:
            import torch

            for t in (base, centroids) + ((orthogonal,) if orthogonal is not None else ()):
                assert t.is_cuda and t.device.index == device and t.dtype == torch.float32 and t.is_contiguous()
            n, ln = base.shape
            k = centroids.shape[0]
            pb, pc = C.c_void_p(base.data_ptr()), C.c_void_p(centroids.data_ptr())
            po = C.c_void_p(orthogonal.data_ptr()) if orthogonal is not None else None
            torch.cuda.synchronize(device)
        else:
            base = _np(base, np.float32); centroids = _np(centroids, np.float32)
            orthogonal = None if orthogonal is None else _np(orthogonal, np.float32)
            n, ln = base.shape
            k = centroids.shape[0]
            pb, pc, po = _ptr(base), _ptr(centroids), _ptr(orthogonal)
        assert centroids.shape[1] == ln
        h = C.c_void_p()
        _check(lib().rabitq_build(pb, n, ln, pc, k, po, seed, int(on_dev), device, C.byref(h)))
        return cls(h, device)

    def dump_to_dir(self, path) -> None:
        """`RaBitQ::dump_to_dir(path)` (src/rabitq.rs:128-156)."""
        _check(lib().rabitq_dump_to_dir(self._h, os.fsencode(str(path))))

    def reshard(self, shard_rank: int, shard_count: int) -> "RaBitQ":
        """Shard `shard_rank` of `shard_count` of this (unsharded) index as a new handle on the same device."""
        h = C.c_void_p()
        _check(lib().rabitq_reshard(self._h, shard_rank, shard_count, C.byref(h)))
        return RaBitQ(h, self.device)

    def export_arrays(self, device_tensors: bool = False) -> dict:
        """The arrays of `struct RaBitQ` as numpy arrays (or torch CUDA tensors on this handle's device)."""
        D, n, k = self.dim, self.num_vectors, self.num_clusters
        if device_tensors:
            import torch

            dev = torch.device("cuda", self.device)
            out = dict(dim=D, base=torch.empty((n, D), dtype=torch.float32, device=dev),
                       orthogonal=torch.empty((D, D), dtype=torch.float32, device=dev),
                       centroids=torch.empty((k, D), dtype=torch.float32, device=dev),
                       offsets=torch.empty((k + 1,), dtype=torch.int32, device=dev),
                       map_ids=torch.empty((n,), dtype=torch.int32, device=dev),
                       codes=torch.empty((n, D // 64), dtype=torch.int64, device=dev),
                       factors=torch.empty((n, 4), dtype=torch.float32, device=dev))
            ptrs = [C.c_void_p(out[key].data_ptr()) for key in ("base", "orthogonal", "centroids", "offsets", "map_ids", "codes", "factors")]
            torch.cuda.synchronize(dev)
        else:
            out = dict(dim=D, base=np.empty((n, D), np.float32), orthogonal=np.empty((D, D), np.float32),
                       centroids=np.empty((k, D), np.float32), offsets=np.empty(k + 1, np.uint32), map_ids=np.empty(n, np.uint32),
                       codes=np.empty((n, D // 64), np.uint64), factors=np.empty((n, 4), np.float32))
            ptrs = [_ptr(out[key]) for key in ("base", "orthogonal", "centroids", "offsets", "map_ids", "codes", "factors")]
        _check(lib().rabitq_export_arrays(self._h, *ptrs, int(device_tensors)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().rabitq_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- shape ---------------------------------------------------------------------------------------------------
    @property
    def dim(self) -> int:
        return int(lib().rabitq_dim(self._h))

    @property
    def num_vectors(self) -> int:
        return int(lib().rabitq_num_vectors(self._h))

    @property
    def num_clusters(self) -> int:
        return int(lib().rabitq_num_clusters(self._h))

    # ---- query ---------------------------------------------------------------------------------------------------
    def query(self, query, probe: int, topk: int, heuristic_rank: bool = False):
        """`RaBitQ::query(&self, query, probe, topk, heuristic_rank) -> Vec<(f32, u32)>` (src/rabitq.rs:268-333).
        Ascending by distance (the reference returns heap order)."""
        q = _np(query, np.float32).reshape(-1)
        d = np.empty(topk, np.float32)
        ids = np.empty(topk, np.uint32)
        cnt = np.zeros(1, np.uint32)
        _check(lib().rabitq_query(self._h, _ptr(q), q.shape[0], probe, topk, int(heuristic_rank), _ptr(d), _ptr(ids), _ptr(cnt)))
        return [(float(d[i]), int(ids[i])) for i in range(int(cnt[0]))]

    def query_batch(self, queries, probe: int, topk: int, heuristic_rank: bool = False):
        """The CLI loop (crates/cli/src/main.rs:69-75) as one call.  numpy in -> numpy out (host buffers, copies
        inside); torch CUDA tensor in -> torch CUDA tensors out (nothing leaves HBM)."""
        if hasattr(queries, "data_ptr"):
            import torch

            assert queries.is_cuda and queries.dtype == torch.float32 and queries.is_contiguous()
            nq, ln = queries.shape
            d = torch.empty((nq, topk), dtype=torch.float32, device=queries.device)
            ids = torch.empty((nq, topk), dtype=torch.int32, device=queries.device)
            cnt = torch.empty((nq,), dtype=torch.int32, device=queries.device)
            torch.cuda.synchronize(queries.device)
            _check(lib().rabitq_query_batch_device(self._h, C.c_void_p(queries.data_ptr()), nq, ln, probe, topk, int(heuristic_rank),
                                                   C.c_void_p(d.data_ptr()), C.c_void_p(ids.data_ptr()), C.c_void_p(cnt.data_ptr())))
            return d, ids, cnt
        q = _np(queries, np.float32)
        nq, ln = q.shape
        d = np.empty((nq, topk), np.float32)
        ids = np.empty((nq, topk), np.uint32)
        cnt = np.zeros(nq, np.uint32)
        _check(lib().rabitq_query_batch(self._h, _ptr(q), nq, ln, probe, topk, int(heuristic_rank), _ptr(d), _ptr(ids), _ptr(cnt)))
        return d, ids, cnt

    def query_batch_device_into(self, queries, probe: int, topk: int, d, ids, cnt, heuristic_rank: bool = False):
        """Device-resident call with caller-owned CUDA output tensors: nothing is allocated and no device-wide synchronisation is
        issued here, so the caller's stream (see `set_stream`) orders the inputs.  The serving loop / bench.py's device leg."""
        nq, ln = queries.shape
        _check(lib().rabitq_query_batch_device(self._h, C.c_void_p(queries.data_ptr()), nq, ln, probe, topk, int(heuristic_rank),
                                               C.c_void_p(d.data_ptr()), C.c_void_p(ids.data_ptr()), C.c_void_p(cnt.data_ptr())))

    def query_batch_into(self, q_host: np.ndarray, probe: int, topk: int, d: np.ndarray, ids: np.ndarray, cnt: np.ndarray,
                         next_q_host: np.ndarray | None = None):
        """Host-buffer call with caller-owned outputs (used by bench.py's end-to-end leg with pinned memory).  `next_q_host`: the
        NEXT batch (same shape), uploaded on a copy stream while this one is answered (`rabitq_query_batch_pipelined`)."""
        nq, ln = q_host.shape
        if next_q_host is None:
            _check(lib().rabitq_query_batch(self._h, _ptr(q_host), nq, ln, probe, topk, 0, _ptr(d), _ptr(ids), _ptr(cnt)))
        else:
            assert next_q_host.shape == q_host.shape and next_q_host.dtype == np.float32
            _check(lib().rabitq_query_batch_pipelined(self._h, _ptr(q_host), _ptr(next_q_host), nq, ln, probe, topk, 0, _ptr(d), _ptr(ids),
                                                      _ptr(cnt)))

    # ---- metrics ---------------------------------------------------------------------------------------------------
    def metrics(self) -> dict:
        m = np.zeros(4, np.uint64)
        lib().rabitq_metrics(self._h, m.ctypes.data_as(c_u64p))
        return dict(query=int(m[0]), rough=int(m[1]), precise=int(m[2]), miss=int(m[3]))

    def metrics_reset(self) -> None:
        lib().rabitq_metrics_reset(self._h)

    def metrics_str(self) -> str:
        """`Metrics::to_str` (src/metrics.rs:30-41)."""
        m = self.metrics()
        ratio = (m["rough"] / m["precise"]) if m["precise"] else float("nan")
        return f"query: {m['query']}, rough: {m['rough']}, precise: {m['precise']}, ratio: {ratio:.2f}, cache miss: {m['miss']}"

    # ---- tuning / measurement --------------------------------------------------------------------------------------
    def set_rounds(self, rounds) -> None:
        r = _np(rounds, np.uint32)
        _check(lib().rabitq_set_rounds(self._h, r.ctypes.data_as(c_u32p), len(r)))

    def set_option(self, name: str, value: int) -> None:
        _check(lib().rabitq_set_option(self._h, name.encode(), int(value)))

    def debug_rerank_stats(self, nq: int) -> np.ndarray:
        """[nq, 2 rounds, (cycles, waves, exact computed, enqueue, wait, l2, replay, stage cycles)] of the last batch (after set_option("debug_rerank", 1))."""
        out = np.zeros((nq, 2, 8), np.uint32)
        _check(lib().rabitq_debug_rerank_stats(self._h, C.c_void_p(out.ctypes.data), nq))
        return out

    def set_quantize_bias(self, bias) -> None:
        """Switch the query quantiser to `scalar_quantize_raw` (src/utils.rs:194-209, the reference on a host without AVX2)
        with this `rand_bias` (dim floats); None restores the AVX2 semantics (round-half-even, no bias)."""
        if bias is None:
            _check(lib().rabitq_set_quantize_bias(self._h, None))
            return
        b = _np(bias, np.float32).reshape(-1)
        assert b.shape[0] == self.dim
        _check(lib().rabitq_set_quantize_bias(self._h, _ptr(b)))

    def set_stream(self, cuda_stream: int | None) -> None:
        """Run on the caller's stream (e.g. `torch.cuda.current_stream().cuda_stream`); None = private stream."""
        _check(lib().rabitq_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def last_timings(self) -> dict:
        ms = np.zeros(10, np.float32)
        cn = np.zeros(6, np.uint64)
        _check(lib().rabitq_last_timings(self._h, ms.ctypes.data_as(c_f32p), cn.ctypes.data_as(c_u64p)))
        out = {f"ms_{n}": float(v) for n, v in zip(TIMING_STAGES, ms)}
        out.update({n: int(v) for n, v in zip(COUNT_NAMES, cn)})
        return out

    # ---- stage-level entries (parity tests) ------------------------------------------------------------------------
    def stage_rotate(self, queries) -> np.ndarray:
        q = _np(queries, np.float32)
        y = np.empty((q.shape[0], self.dim), np.float32)
        _check(lib().rabitq_stage_rotate(self._h, _ptr(q), q.shape[0], q.shape[1], _ptr(y)))
        return y

    def stage_probe(self, queries, probe: int, want_all: bool = True):
        q = _np(queries, np.float32)
        nq, K = q.shape[0], self.num_clusters
        P = min(probe, K)
        cd = np.empty((nq, K), np.float32) if want_all else None
        pid = np.empty((nq, P), np.uint32)
        pd = np.empty((nq, P), np.float32)
        _check(lib().rabitq_stage_probe(self._h, _ptr(q), nq, q.shape[1], probe, _ptr(cd), _ptr(pid), _ptr(pd)))
        return cd, pid, pd

    def stage_quantize(self, queries, probe: int):
        q = _np(queries, np.float32)
        nq, K, W = q.shape[0], self.num_clusters, self.dim // 64
        P = min(probe, K)
        lo = np.empty((nq, P), np.float32)
        delta = np.empty((nq, P), np.float32)
        s = np.empty((nq, P), np.uint32)
        planes = np.empty((nq, P, 4 * W), np.uint64)
        _check(lib().rabitq_stage_quantize(self._h, _ptr(q), nq, q.shape[1], probe, _ptr(lo), _ptr(delta), _ptr(s), _ptr(planes)))
        return lo, delta, s, planes

    def stage_scan(self, queries, probe: int, pair_capacity: int):
        q = _np(queries, np.float32)
        nq = q.shape[0]
        rough = np.empty(pair_capacity, np.float32)
        abdp = np.empty(pair_capacity, np.uint32)
        start = np.zeros(nq + 1, np.uint64)
        _check(lib().rabitq_stage_scan(self._h, _ptr(q), nq, q.shape[1], probe, pair_capacity, _ptr(rough), _ptr(abdp), _ptr(start)))
        total = int(start[nq])
        return rough[:total], abdp[:total], start


def calculate_recall(truth, res, topk: int) -> float:
    """`calculate_recall` (src/utils.rs:367-379): set membership over the first `topk` truth ids."""
    assert len(res) == topk
    t = list(truth)[:topk]
    return sum(1 for i in res if i in t) / float(topk)
