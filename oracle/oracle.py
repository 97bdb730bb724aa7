"""ctypes wrapper over oracle/liboracle.so -- the CPU oracle of the IVF-RaBitQ query path.

TEST INFRASTRUCTURE ONLY (see the header of rabitq_oracle.cpp).  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by rabitq_b200.
PARITY UNPINNED: the reference (Rust) cannot be built in this image and ships no golden vectors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

c_f32p = C.POINTER(C.c_float)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)
c_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rabitq_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    L.orc_last_error.restype = C.c_char_p
    L.orc_load_from_dir.restype = C.c_void_p
    L.orc_load_from_dir.argtypes = [C.c_char_p]
    L.orc_dump_to_dir.argtypes = [C.c_void_p, C.c_char_p]
    L.orc_build.restype = C.c_void_p
    L.orc_build.argtypes = [c_f32p, C.c_size_t, C.c_size_t, c_f32p, C.c_size_t, c_f32p, C.c_uint64, C.c_int]
    L.orc_from_arrays.restype = C.c_void_p
    L.orc_from_arrays.argtypes = [C.c_uint32, C.c_size_t, C.c_size_t, c_f32p, c_f32p, c_f32p, c_u32p, c_u32p, c_u64p, c_f32p]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_set_raw_bias.argtypes = [C.c_void_p, c_f32p]
    L.orc_set_raw_bias.restype = None
    L.orc_dim.restype = C.c_uint32
    L.orc_dim.argtypes = [C.c_void_p]
    for name in ("orc_n", "orc_k"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [C.c_void_p]
    for name, t in (("orc_base", c_f32p), ("orc_orthogonal", c_f32p), ("orc_centroids", c_f32p), ("orc_offsets", c_u32p),
                    ("orc_map_ids", c_u32p), ("orc_codes", c_u64p), ("orc_factors", c_f32p)):
        getattr(L, name).restype = t
        getattr(L, name).argtypes = [C.c_void_p]
    L.orc_query.restype = C.c_int
    L.orc_query.argtypes = [C.c_void_p, c_f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, c_f32p, c_u32p]
    L.orc_query_batch.restype = C.c_double
    L.orc_query_batch.argtypes = [C.c_void_p, c_f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                  c_f32p, c_u32p, c_u32p, c_u64p]
    L.orc_query_trace.restype = C.c_int
    L.orc_query_trace.argtypes = [C.c_void_p, c_f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, c_f32p, c_u32p,
                                  c_f32p, c_f32p, c_u32p, c_f32p, c_f32p, c_f32p, c_u32p, c_u64p, c_u8p,
                                  C.c_size_t, c_f32p, c_u32p, c_u32p, c_f32p, c_u8p, c_u64p, c_u64p]
    L.orc_metrics.argtypes = [c_u64p]
    L.orc_l2_squared_distance.restype = C.c_float
    L.orc_l2_squared_distance.argtypes = [c_f32p, c_f32p, C.c_size_t]
    L.orc_vector_dot_product.restype = C.c_float
    L.orc_vector_dot_product.argtypes = [c_f32p, c_f32p, C.c_size_t]
    L.orc_min_max_residual.argtypes = [c_f32p, c_f32p, c_f32p, C.c_size_t, c_f32p, c_f32p]
    L.orc_min_max_raw.argtypes = [c_f32p, c_f32p, c_f32p, C.c_size_t, c_f32p, c_f32p]
    L.orc_scalar_quantize.restype = C.c_uint32
    L.orc_scalar_quantize.argtypes = [c_u8p, c_f32p, C.c_size_t, C.c_float, C.c_float]
    L.orc_scalar_quantize_raw.restype = C.c_uint32
    L.orc_scalar_quantize_raw.argtypes = [c_u8p, c_f32p, c_f32p, C.c_size_t, C.c_float, C.c_float]
    L.orc_vector_binarize_query.argtypes = [c_u8p, C.c_size_t, c_u64p]
    L.orc_vector_binarize_query_raw.argtypes = [c_u8p, C.c_size_t, c_u64p]
    for name in ("orc_binary_dot_product", "orc_binary_dot_product_raw", "orc_asymmetric_binary_dot_product"):
        getattr(L, name).restype = C.c_uint32
        getattr(L, name).argtypes = [c_u64p, c_u64p, C.c_size_t]
    L.orc_ord32_from_f32.restype = C.c_int32
    L.orc_ord32_from_f32.argtypes = [C.c_float]
    L.orc_ord32_to_f32.restype = C.c_float
    L.orc_ord32_to_f32.argtypes = [C.c_int32]
    L.orc_scalar_const.restype = C.c_float
    L.orc_gen_orthogonal.argtypes = [C.c_size_t, C.c_uint64, c_f32p]
    L.orc_heap_replay.restype = C.c_int
    L.orc_heap_replay.argtypes = [c_f32p, c_u32p, C.c_size_t, C.c_size_t, c_f32p, c_u32p]
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleIndex:
    """Mirror of `RaBitQ` (src/rabitq.rs:57-68) on the CPU oracle."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError(lib().orc_last_error().decode())
        self._h = C.c_void_p(handle)

    # -- constructors ------------------------------------------------------------------------------
    @classmethod
    def load_from_dir(cls, path: str) -> "OracleIndex":
        return cls(lib().orc_load_from_dir(os.fsencode(path)))

    @classmethod
    def from_arrays(cls, base, centroids, P=None, seed: int = 1, nthreads: int = 8) -> "OracleIndex":
        """`RaBitQ::from_path` (src/rabitq.rs:159-265) on in-memory arrays (original space)."""
        base = _f32(base)
        centroids = _f32(centroids)
        n, ln = base.shape
        k = centroids.shape[0]
        assert centroids.shape[1] == ln
        Pp = None
        if P is not None:
            P = _f32(P)
            Pp = _p(P, c_f32p)
        return cls(lib().orc_build(_p(base, c_f32p), n, ln, _p(centroids, c_f32p), k, Pp, seed, nthreads))

    @classmethod
    def from_built(cls, dim, base, orthogonal, centroids, offsets, map_ids, codes, factors) -> "OracleIndex":
        base = _f32(base); orthogonal = _f32(orthogonal); centroids = _f32(centroids)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        map_ids = np.ascontiguousarray(map_ids, dtype=np.uint32)
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        factors = _f32(factors)
        return cls(lib().orc_from_arrays(dim, base.shape[0], centroids.shape[0], _p(base, c_f32p), _p(orthogonal, c_f32p),
                                         _p(centroids, c_f32p), _p(offsets, c_u32p), _p(map_ids, c_u32p),
                                         _p(codes, c_u64p), _p(factors, c_f32p)))

    def set_raw_bias(self, bias) -> None:
        """bias (dim floats): quantise queries with scalar_quantize_raw (src/utils.rs:194-209); None: the AVX2 branch."""
        if bias is None:
            lib().orc_set_raw_bias(self._h, None)
        else:
            b = _f32(bias).reshape(-1)
            assert b.shape[0] == self.dim
            lib().orc_set_raw_bias(self._h, _p(b, c_f32p))

    def dump_to_dir(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        if lib().orc_dump_to_dir(self._h, os.fsencode(path)) != 0:
            raise RuntimeError("dump_to_dir failed")

    def close(self):
        if self._h:
            lib().orc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- views (copies) ----------------------------------------------------------------------------
    @property
    def dim(self): return int(lib().orc_dim(self._h))
    @property
    def n(self): return int(lib().orc_n(self._h))
    @property
    def k(self): return int(lib().orc_k(self._h))

    def _arr(self, fn, shape, dtype):
        ptr = fn(self._h)
        return np.ctypeslib.as_array(ptr, shape=shape).astype(dtype, copy=True)

    def arrays(self) -> dict:
        D, n, k = self.dim, self.n, self.k
        L = lib()
        return dict(
            dim=D,
            base=self._arr(L.orc_base, (n, D), np.float32),
            orthogonal=self._arr(L.orc_orthogonal, (D, D), np.float32),
            centroids=self._arr(L.orc_centroids, (k, D), np.float32),
            offsets=self._arr(L.orc_offsets, (k + 1,), np.uint32),
            map_ids=self._arr(L.orc_map_ids, (n,), np.uint32),
            codes=self._arr(L.orc_codes, (n, D // 64), np.uint64),
            factors=self._arr(L.orc_factors, (n, 4), np.float32),
        )

    # -- queries -----------------------------------------------------------------------------------
    def query(self, q, probe: int, topk: int, heuristic_rank: bool = False):
        """`RaBitQ::query`: list of (dist, id) in the reference's (heap) order."""
        q = _f32(q)
        d = np.empty(topk + 1, np.float32)
        ids = np.empty(topk + 1, np.uint32)
        c = lib().orc_query(self._h, _p(q, c_f32p), q.shape[0], probe, topk, int(heuristic_rank), _p(d, c_f32p), _p(ids, c_u32p))
        if c < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return [(float(d[i]), int(ids[i])) for i in range(c)]

    def query_batch(self, queries, probe: int, topk: int, heuristic_rank: bool = False, nthreads: int = 1):
        queries = _f32(queries)
        nq, ln = queries.shape
        d = np.full((nq, topk), np.inf, np.float32)
        ids = np.full((nq, topk), 0xFFFFFFFF, np.uint32)
        cnt = np.zeros(nq, np.uint32)
        counters = np.zeros(2, np.uint64)
        secs = lib().orc_query_batch(self._h, _p(queries, c_f32p), nq, ln, probe, topk, int(heuristic_rank), nthreads,
                                     _p(d, c_f32p), _p(ids, c_u32p), _p(cnt, c_u32p), _p(counters, c_u64p))
        if secs < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return dict(dist=d, ids=ids, count=cnt, rough=int(counters[0]), precise=int(counters[1]), seconds=float(secs))

    def trace(self, q, probe: int, topk: int, heuristic_rank: bool = False, pair_capacity: int | None = None) -> dict:
        """One query with every intermediate of src/rabitq.rs:282-329 exposed."""
        q = _f32(q)
        D, K = self.dim, self.k
        W = D // 64
        P = min(probe, K)
        if pair_capacity is None:
            pair_capacity = self.n
        out = dict(
            y=np.empty(D, np.float32), centroid_dist=np.empty(K, np.float32), probe_ids=np.empty(P, np.uint32),
            probe_dist=np.empty(P, np.float32), lo=np.empty(P, np.float32), delta=np.empty(P, np.float32),
            sum=np.empty(P, np.uint32), planes=np.zeros((P, 4 * W), np.uint64), quantized=np.empty((P, D), np.uint8),
            rough=np.empty(pair_capacity, np.float32), abdp=np.empty(pair_capacity, np.uint32),
            pair_pos=np.empty(pair_capacity, np.uint32), exact=np.empty(pair_capacity, np.float32),
            action=np.empty(pair_capacity, np.uint8))
        d = np.empty(topk + 1, np.float32)
        ids = np.empty(topk + 1, np.uint32)
        pairs = C.c_uint64(0)
        precise = C.c_uint64(0)
        c = lib().orc_query_trace(self._h, _p(q, c_f32p), q.shape[0], probe, topk, int(heuristic_rank), _p(d, c_f32p), _p(ids, c_u32p),
                                  _p(out["y"], c_f32p), _p(out["centroid_dist"], c_f32p), _p(out["probe_ids"], c_u32p),
                                  _p(out["probe_dist"], c_f32p), _p(out["lo"], c_f32p), _p(out["delta"], c_f32p),
                                  _p(out["sum"], c_u32p), _p(out["planes"], c_u64p), _p(out["quantized"], c_u8p),
                                  pair_capacity, _p(out["rough"], c_f32p), _p(out["abdp"], c_u32p), _p(out["pair_pos"], c_u32p),
                                  _p(out["exact"], c_f32p), _p(out["action"], c_u8p), C.byref(pairs), C.byref(precise))
        if c < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        npairs = int(pairs.value)
        assert npairs <= pair_capacity
        for key in ("rough", "abdp", "pair_pos", "exact", "action"):
            out[key] = out[key][:npairs]
        out["pairs"] = npairs
        out["precise"] = int(precise.value)
        out["result"] = [(float(d[i]), int(ids[i])) for i in range(c)]
        return out


def metrics() -> dict:
    m = np.zeros(4, np.uint64)
    lib().orc_metrics(_p(m, c_u64p))
    return dict(query=int(m[0]), rough=int(m[1]), precise=int(m[2]), miss=int(m[3]))


def metrics_reset() -> None:
    lib().orc_metrics_reset()


def gen_orthogonal(D: int, seed: int) -> np.ndarray:
    out = np.empty((D, D), np.float32)
    lib().orc_gen_orthogonal(D, seed, _p(out, c_f32p))
    return out
