"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes loaders for the two CPU checkers:
  * ``port``  -- oracle/liboracle_port.so, our plain-C restatement (oracle/knn_port.c, oracle/ft_port.c)
  * ``ref``   -- oracle/_ref/liboracle_ref_*.so, the reference's own translation units compiled in place from
                 /root/reference by oracle/Makefile (prebuilt files travel to the GPU box; the sources do not)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package (reindexer_b200) must never import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
L2, IP, COS = 0, 1, 2

_f32p = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def _p(a, t):
    return a.ctypes.data_as(t)


def build(ref: bool = True) -> None:
    """Build the checkers (building the checker is not using it)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref and os.path.isdir("/root/reference/cpp_src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_port = None
_ref_knn = None


def port_lib():
    global _port
    if _port is None:
        path = os.path.join(HERE, "liboracle_port.so")
        if not os.path.exists(path):
            build(ref=False)
        lib = C.CDLL(path)
        lib.port_synth_value.restype = C.c_float
        lib.port_synth_value.argtypes = [C.c_uint64, C.c_uint64]
        lib.port_synth_fill.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _f32p]
        for fn in (lib.port_l2sqr, lib.port_ip):
            fn.restype = C.c_float
            fn.argtypes = [_f32p, _f32p, C.c_size_t]
        lib.port_calc_l2_module.restype = C.c_float
        lib.port_calc_l2_module.argtypes = [_f32p, C.c_int32]
        lib.port_normalize_copy.restype = C.c_float
        lib.port_normalize_copy.argtypes = [_f32p, C.c_int32, _f32p]
        lib.port_bf_create.restype = C.c_void_p
        lib.port_bf_create.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
        lib.port_bf_clone.restype = C.c_void_p
        lib.port_bf_clone.argtypes = [C.c_void_p, C.c_size_t]
        lib.port_bf_destroy.argtypes = [C.c_void_p]
        for fn in (lib.port_bf_size, lib.port_bf_capacity, lib.port_bf_element_size):
            fn.restype = C.c_size_t
            fn.argtypes = [C.c_void_p]
        lib.port_bf_add.argtypes = [C.c_void_p, _f32p, C.c_uint64]
        lib.port_bf_remove.argtypes = [C.c_void_p, C.c_uint64]
        lib.port_bf_resize.argtypes = [C.c_void_p, C.c_size_t]
        lib.port_bf_get.restype = _f32p
        lib.port_bf_get.argtypes = [C.c_void_p, C.c_uint64]
        lib.port_bf_search_knn.restype = C.c_int64
        lib.port_bf_search_knn.argtypes = [C.c_void_p, _f32p, C.c_size_t, _f32p, _u64p]
        lib.port_bf_search_range.restype = C.c_int64
        lib.port_bf_search_range.argtypes = [C.c_void_p, _f32p, C.c_float, C.c_size_t, _f32p, _u64p]
        lib.port_select_postprocess.restype = C.c_size_t
        lib.port_select_postprocess.argtypes = [C.c_void_p, C.c_size_t, _f32p, _u64p, _i32p, _f32p]
        _port = lib
    return _port


def ref_knn_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "liboracle_ref_knn.so"))


def ref_knn_lib():
    global _ref_knn
    if _ref_knn is None:
        lib = C.CDLL(os.path.join(HERE, "_ref", "liboracle_ref_knn.so"))
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_isa_level.restype = C.c_int
        for fn in (lib.ref_l2sqr, lib.ref_ip):
            fn.restype = C.c_float
            fn.argtypes = [_f32p, _f32p, C.c_size_t]
        lib.ref_calc_l2_module.restype = C.c_float
        lib.ref_calc_l2_module.argtypes = [_f32p, C.c_int32]
        lib.ref_normalize_copy.restype = C.c_float
        lib.ref_normalize_copy.argtypes = [_f32p, C.c_int32, _f32p]
        lib.ref_bf_create.restype = C.c_void_p
        lib.ref_bf_create.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
        lib.ref_bf_clone.restype = C.c_void_p
        lib.ref_bf_clone.argtypes = [C.c_void_p, C.c_size_t]
        lib.ref_bf_destroy.argtypes = [C.c_void_p]
        for fn in (lib.ref_bf_size, lib.ref_bf_capacity, lib.ref_bf_element_size):
            fn.restype = C.c_size_t
            fn.argtypes = [C.c_void_p]
        lib.ref_bf_add.argtypes = [C.c_void_p, C.c_size_t, _f32p, C.c_uint64]
        lib.ref_bf_add_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, _u64p, _f32p]
        lib.ref_bf_remove.argtypes = [C.c_void_p, C.c_uint64]
        lib.ref_bf_resize.argtypes = [C.c_void_p, C.c_size_t]
        lib.ref_bf_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(_f32p)]
        lib.ref_bf_search_knn.restype = C.c_int64
        lib.ref_bf_search_knn.argtypes = [C.c_void_p, _f32p, C.c_size_t, _f32p, _u64p]
        lib.ref_bf_search_range.restype = C.c_int64
        lib.ref_bf_search_range.argtypes = [C.c_void_p, _f32p, C.c_float, C.c_size_t, _f32p, _u64p]
        lib.ref_bf_search_knn_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, _f32p, C.c_size_t, C.c_int, _f32p, _u64p, _u32p]
        lib.ref_hnsw_create.restype = C.c_void_p
        lib.ref_hnsw_create.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        lib.ref_hnsw_destroy.argtypes = [C.c_void_p]
        lib.ref_hnsw_size.restype = C.c_size_t
        lib.ref_hnsw_size.argtypes = [C.c_void_p]
        lib.ref_hnsw_add_batch.argtypes = [C.c_void_p, C.c_size_t, _u64p, _f32p, C.c_int]
        lib.ref_hnsw_search_knn.restype = C.c_int64
        lib.ref_hnsw_search_knn.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_size_t, C.c_size_t, _f32p, _u64p]
        lib.ref_hnsw_mark_delete.argtypes = [C.c_void_p, C.c_uint64]
        lib.ref_hnsw_quantize.restype = C.c_void_p
        lib.ref_hnsw_quantize.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_size_t]
        lib.ref_hnsw_q_destroy.argtypes = [C.c_void_p]
        lib.ref_hnsw_q_params.argtypes = [C.c_void_p, _f32p]
        lib.ref_hnsw_q_export.argtypes = [C.c_void_p, C.c_void_p, _f32p]
        lib.ref_hnsw_q_prepare_query.argtypes = [C.c_void_p, _f32p, C.c_void_p, _f32p]
        lib.ref_hnsw_q_search_knn.restype = C.c_int64
        lib.ref_hnsw_q_search_knn.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_size_t, C.c_size_t, _f32p, _u64p]
        lib.ref_hnsw_stream_begin.restype = C.c_void_p
        lib.ref_hnsw_stream_begin.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_size_t]
        lib.ref_hnsw_stream_next.restype = C.c_int64
        lib.ref_hnsw_stream_next.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, _f32p, _u64p, C.POINTER(C.c_int)]
        lib.ref_hnsw_stream_end.argtypes = [C.c_void_p]
        lib.ref_hnsw_search_range.restype = C.c_int64
        lib.ref_hnsw_search_range.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_float, C.c_size_t, C.c_size_t, _f32p, _u64p]
        lib.ref_hnsw_search_knn_batch.argtypes = [C.c_void_p, C.c_uint32, _f32p, _f32p, C.c_size_t, C.c_size_t, C.c_int, _f32p,
                                                  _u64p, _u32p]
        lib.ref_hnsw_search_metrics.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_size_t, _i64p, _i64p]
        lib.ref_hnsw_export_header.argtypes = [C.c_void_p, _i64p]
        lib.ref_hnsw_export.argtypes = [C.c_void_p, _u32p, _i32p, _i64p, _u32p, _u64p, _f32p]
        _ref_knn = lib
    return _ref_knn


# ----------------------------------------------------------------------------- synthetic data
_MASK = (1 << 64) - 1


def synth(seed: int, first_index: int, count: int) -> np.ndarray:
    """numpy twin of port_synth_fill / csrc/synth.cuh (bit-identical)."""
    idx = np.arange(first_index, first_index + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) ^ (idx * np.uint64(0xD1342543DE82EF95))
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        h = z ^ (z >> np.uint64(31))
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)) + ((h >> np.uint64(32)) & np.uint64(0xFFFF))
         + (h >> np.uint64(48))).astype(np.int64)
    return ((s - 131070).astype(np.float32) * np.float32(6.6072488e-06)).astype(np.float32)


def synth_matrix(seed: int, n: int, dim: int, first_row: int = 0) -> np.ndarray:
    return synth(seed, first_row * dim, n * dim).reshape(n, dim)


def row_labels(n: int, first_row: int = 0, array_idx: int = 0) -> np.ndarray:
    """FloatVectorId{rowId, arrayIdx}.AsNumber() (cpp_src/core/index/float_vector/float_vector_id.h:11)."""
    return ((np.arange(first_row, first_row + n, dtype=np.uint64) << np.uint64(32)) | np.uint64(array_idx)).astype(np.uint64)


# ----------------------------------------------------------------------------- brute-force wrappers (same surface for port / ref)
class _BFBase:
    def search_knn(self, q, k):
        q = np.ascontiguousarray(q, dtype=np.float32)
        d = np.empty(max(k, 1), np.float32)
        l = np.empty(max(k, 1), np.uint64)
        n = self._knn(q, k, d, l)
        assert n >= 0, self._err()
        return d[:n].copy(), l[:n].copy()

    def search_range(self, q, radius, max_out=None):
        q = np.ascontiguousarray(q, dtype=np.float32)
        max_out = self.size() if max_out is None else max_out
        d = np.empty(max(max_out, 1), np.float32)
        l = np.empty(max(max_out, 1), np.uint64)
        n = self._range(q, float(radius), max_out, d, l)
        assert n >= 0, self._err()
        n = min(n, max_out)
        return d[:n].copy(), l[:n].copy()

    def add_batch(self, labels, vecs):
        vecs = np.ascontiguousarray(vecs, dtype=np.float32)
        for i in range(len(labels)):
            rc = self.add(vecs[i], int(labels[i]))
            if rc:
                return rc
        return 0


class PortBF(_BFBase):
    def __init__(self, metric, dim, capacity, _h=None):
        self.lib = port_lib()
        self.dim = dim
        self.h = _h if _h is not None else self.lib.port_bf_create(metric, dim, capacity)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.port_bf_destroy(self.h)
            self.h = None

    def _err(self):
        return "port error"

    def clone(self, cap):
        return PortBF(None, self.dim, None, _h=self.lib.port_bf_clone(self.h, cap))

    def size(self):
        return self.lib.port_bf_size(self.h)

    def capacity(self):
        return self.lib.port_bf_capacity(self.h)

    def element_size(self):
        return self.lib.port_bf_element_size(self.h)

    def add(self, vec, label):
        vec = np.ascontiguousarray(vec, dtype=np.float32)
        return self.lib.port_bf_add(self.h, _p(vec, _f32p), label)

    def remove(self, label):
        return self.lib.port_bf_remove(self.h, label)

    def resize(self, cap):
        return self.lib.port_bf_resize(self.h, cap)

    def get(self, label):
        p = self.lib.port_bf_get(self.h, label)
        return None if not p else np.ctypeslib.as_array(p, (self.dim,)).copy()

    def _knn(self, q, k, d, l):
        return self.lib.port_bf_search_knn(self.h, _p(q, _f32p), k, _p(d, _f32p), _p(l, _u64p))

    def _range(self, q, radius, max_out, d, l):
        return self.lib.port_bf_search_range(self.h, _p(q, _f32p), radius, max_out, _p(d, _f32p), _p(l, _u64p))


class RefBF(_BFBase):
    def __init__(self, metric, dim, capacity, _h=None):
        self.lib = ref_knn_lib()
        self.dim = dim
        self.h = _h if _h is not None else self.lib.ref_bf_create(metric, dim, capacity)
        assert self.h, self._err()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_bf_destroy(self.h)
            self.h = None

    def _err(self):
        return self.lib.ref_last_error().decode()

    def clone(self, cap):
        return RefBF(None, self.dim, None, _h=self.lib.ref_bf_clone(self.h, cap))

    def size(self):
        return self.lib.ref_bf_size(self.h)

    def capacity(self):
        return self.lib.ref_bf_capacity(self.h)

    def element_size(self):
        return self.lib.ref_bf_element_size(self.h)

    def add(self, vec, label):
        vec = np.ascontiguousarray(vec, dtype=np.float32)
        return self.lib.ref_bf_add(self.h, self.dim, _p(vec, _f32p), label)

    def add_batch(self, labels, vecs):
        vecs = np.ascontiguousarray(vecs, dtype=np.float32)
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        return self.lib.ref_bf_add_batch(self.h, self.dim, len(labels), _p(labels, _u64p), _p(vecs, _f32p))

    def remove(self, label):
        return self.lib.ref_bf_remove(self.h, label)

    def resize(self, cap):
        return self.lib.ref_bf_resize(self.h, cap)

    def get(self, label):
        p = _f32p()
        rc = self.lib.ref_bf_get(self.h, label, C.byref(p))
        return None if rc else np.ctypeslib.as_array(p, (self.dim,)).copy()

    def _knn(self, q, k, d, l):
        return self.lib.ref_bf_search_knn(self.h, _p(q, _f32p), k, _p(d, _f32p), _p(l, _u64p))

    def _range(self, q, radius, max_out, d, l):
        return self.lib.ref_bf_search_range(self.h, _p(q, _f32p), radius, max_out, _p(d, _f32p), _p(l, _u64p))

    def search_knn_batch(self, queries, k, threads=1):
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = queries.shape[0]
        d = np.zeros((nq, k), np.float32)
        l = np.zeros((nq, k), np.uint64)
        c = np.zeros(nq, np.uint32)
        rc = self.lib.ref_bf_search_knn_batch(self.h, self.dim, nq, _p(queries, _f32p), k, threads, _p(d, _f32p), _p(l, _u64p),
                                              _p(c, _u32p))
        assert rc == 0, self._err()
        return d, l, c


def best_bf(metric, dim, capacity):
    """The strongest available checker: the reference's own code when its .so is present, else the port."""
    return RefBF(metric, dim, capacity) if ref_knn_available() else PortBF(metric, dim, capacity)


def normalize_copy(x: np.ndarray, use_ref: bool | None = None):
    """NormalizeCopyVector: returns (normalised copy, coefficient 1/||x||)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    if use_ref is None:
        use_ref = ref_knn_available()
    if use_ref:
        k = ref_knn_lib().ref_normalize_copy(_p(x, _f32p), x.size, _p(out, _f32p))
    else:
        k = port_lib().port_normalize_copy(_p(x, _f32p), x.size, _p(out, _f32p))
    return out, np.float32(k)


class SelectOpts(C.Structure):
    _fields_ = [("metric", C.c_int), ("need_sort", C.c_int), ("is_array", C.c_int), ("raw", C.c_int), ("has_k", C.c_int),
                ("k", C.c_size_t), ("has_radius", C.c_int)]


def select_postprocess(metric, dists, labels, need_sort=True, is_array=False, raw=False, k=None, has_radius=False):
    lib = port_lib()
    dists = np.ascontiguousarray(dists, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.uint64)
    n = len(dists)
    ids = np.empty(max(n, 1), np.int32)
    ranks = np.empty(max(n, 1), np.float32)
    o = SelectOpts(metric, int(need_sort), int(is_array), int(raw), int(k is not None), k or 0, int(has_radius))
    cnt = lib.port_select_postprocess(C.byref(o), n, _p(dists, _f32p), _p(labels, _u64p), _p(ids, _i32p), _p(ranks, _f32p))
    return ids[:cnt].copy(), ranks[:cnt].copy()


# ----------------------------------------------------------------------------- HNSW (reference only)
class RefHnsw:
    def __init__(self, metric, dim, capacity, M=16, ef_construction=200, seed=100, multithread=False):
        self.lib = ref_knn_lib()
        self.metric, self.dim, self.M = metric, dim, M
        self.h = self.lib.ref_hnsw_create(metric, dim, capacity, M, ef_construction, seed, int(multithread))
        assert self.h, self.lib.ref_last_error().decode()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_hnsw_destroy(self.h)
            self.h = None

    def size(self):
        return self.lib.ref_hnsw_size(self.h)

    def add_batch(self, labels, vecs, threads=1):
        vecs = np.ascontiguousarray(vecs, dtype=np.float32)
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        rc = self.lib.ref_hnsw_add_batch(self.h, len(labels), _p(labels, _u64p), _p(vecs, _f32p), threads)
        assert rc == 0, self.lib.ref_last_error().decode()

    def search_knn(self, q, k, ef=0, qnorm=None):
        q = np.ascontiguousarray(q, dtype=np.float32)
        d = np.empty(max(k, 1), np.float32)
        l = np.empty(max(k, 1), np.uint64)
        n = self.lib.ref_hnsw_search_knn(self.h, _p(q, _f32p), int(qnorm is not None), float(qnorm or 0.0), k, ef, _p(d, _f32p),
                                         _p(l, _u64p))
        assert n >= 0, self.lib.ref_last_error().decode()
        return d[:n].copy(), l[:n].copy()

    def quantize(self, quantile=None, sample_size=0):
        """the reference's SQ8 copy of this graph (HierarchicalNSW::Impl::Quantize) -> RefHnswSq8"""
        return RefHnswSq8(self, quantile, sample_size)

    def stream(self, q, batch_size, ef=0, max_batches=10**9):
        """BeginStreamingSearch / ContinueStreamingSearch (hnswalg.h:1864-1975): yields (dist, label) batches, best first inside a batch"""
        q = np.ascontiguousarray(q, dtype=np.float32)
        s = self.lib.ref_hnsw_stream_begin(self.h, _p(q, _f32p), 0, 0.0, ef)
        assert s, self.lib.ref_last_error().decode()
        try:
            for _ in range(max_batches):
                d = np.empty(batch_size, np.float32)
                l = np.empty(batch_size, np.uint64)
                ex = C.c_int(0)
                n = self.lib.ref_hnsw_stream_next(self.h, s, batch_size, _p(d, _f32p), _p(l, _u64p), C.byref(ex))
                assert n >= 0, self.lib.ref_last_error().decode()
                yield d[:n].copy(), l[:n].copy()
                if ex.value:
                    break
        finally:
            self.lib.ref_hnsw_stream_end(s)

    def mark_delete(self, label):
        rc = self.lib.ref_hnsw_mark_delete(self.h, int(label))
        assert rc == 0, self.lib.ref_last_error().decode()

    def search_range(self, q, radius, ef, qnorm=None, max_out=None):
        q = np.ascontiguousarray(q, dtype=np.float32)
        max_out = self.size() if max_out is None else max_out
        d = np.empty(max(max_out, 1), np.float32)
        l = np.empty(max(max_out, 1), np.uint64)
        n = self.lib.ref_hnsw_search_range(self.h, _p(q, _f32p), int(qnorm is not None), float(qnorm or 0.0), float(radius), ef, max_out,
                                           _p(d, _f32p), _p(l, _u64p))
        assert n >= 0, self.lib.ref_last_error().decode()
        m = min(n, max_out)
        return d[:m].copy(), l[:m].copy(), n

    def search_knn_batch(self, queries, k, ef=0, threads=1):
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = queries.shape[0]
        d = np.zeros((nq, k), np.float32)
        l = np.zeros((nq, k), np.uint64)
        c = np.zeros(nq, np.uint32)
        rc = self.lib.ref_hnsw_search_knn_batch(self.h, nq, _p(queries, _f32p), None, k, ef, threads, _p(d, _f32p), _p(l, _u64p),
                                                _p(c, _u32p))
        assert rc == 0, self.lib.ref_last_error().decode()
        return d, l, c

    def search_metrics(self, q, ef):
        q = np.ascontiguousarray(q, dtype=np.float32)
        dc, hops = C.c_int64(0), C.c_int64(0)
        rc = self.lib.ref_hnsw_search_metrics(self.h, _p(q, _f32p), 0, 0.0, ef, C.byref(dc), C.byref(hops))
        assert rc == 0, self.lib.ref_last_error().decode()
        return dc.value, hops.value

    def export(self, with_vectors=True):
        """Graph arrays for the device upload (see ref_hnsw_export in oracle/ref_knn_facade.cc)."""
        hdr = np.zeros(6, np.int64)
        assert self.lib.ref_hnsw_export_header(self.h, _p(hdr, _i64p)) == 0
        n, maxlevel, ep, M, m0, upper_slots = (int(x) for x in hdr)
        level0 = np.zeros((n, 1 + m0), np.uint32)
        levels = np.zeros(n, np.int32)
        offs = np.zeros(n + 1, np.int64)
        upper = np.zeros((max(upper_slots, 1), 1 + M), np.uint32)
        labels = np.zeros(n, np.uint64)
        vecs = np.zeros((n, self.dim), np.float32) if with_vectors else None
        rc = self.lib.ref_hnsw_export(self.h, _p(level0, _u32p), _p(levels, _i32p), _p(offs, _i64p), _p(upper, _u32p),
                                      _p(labels, _u64p), _p(vecs, _f32p) if with_vectors else None)
        assert rc == 0, self.lib.ref_last_error().decode()
        return dict(n=n, maxlevel=maxlevel, enterpoint=ep, M=M, maxM0=m0, level0=level0, levels=levels, upper_offsets=offs,
                    upper=upper[:upper_slots], labels=labels, vectors=vecs)


class RefHnswSq8:
    """HierarchicalNSWImpl<uint8_t> built by the reference from a float graph (scalar_quantization/quantizer.h, hnswlib.h:192-197)"""

    def __init__(self, graph, quantile, sample_size):
        self.lib, self.dim, self.graph = graph.lib, graph.dim, graph
        self.h = self.lib.ref_hnsw_quantize(graph.h, int(quantile is not None), float(quantile or 0.0), sample_size)
        assert self.h, self.lib.ref_last_error().decode()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_hnsw_q_destroy(self.h)
            self.h = None

    def params(self):
        p = np.zeros(5, np.float32)
        assert self.lib.ref_hnsw_q_params(self.h, _p(p, _f32p)) == 0
        return dict(minQ=p[0], maxQ=p[1], alpha=p[2], alpha_2=p[3], delta=p[4])

    def export(self):
        n = self.graph.size()
        codes = np.zeros((n, self.dim), np.uint8)
        offs = np.zeros(n, np.float32)
        assert self.lib.ref_hnsw_q_export(self.h, codes.ctypes.data, _p(offs, _f32p)) == 0
        return codes, offs

    def prepare_query(self, q):
        q = np.ascontiguousarray(q, np.float32)
        codes = np.zeros(self.dim, np.uint8)
        off = np.zeros(1, np.float32)
        assert self.lib.ref_hnsw_q_prepare_query(self.h, _p(q, _f32p), codes.ctypes.data, _p(off, _f32p)) == 0
        return codes, float(off[0])

    def search_knn(self, q, k, ef=0, qnorm=None):
        q = np.ascontiguousarray(q, np.float32)
        d = np.empty(max(k, 1), np.float32)
        l = np.empty(max(k, 1), np.uint64)
        n = self.lib.ref_hnsw_q_search_knn(self.h, _p(q, _f32p), int(qnorm is not None), float(qnorm or 0.0), k, ef, _p(d, _f32p), _p(l, _u64p))
        assert n >= 0, self.lib.ref_last_error().decode()
        return d[:n].copy(), l[:n].copy()


# ---- IVF: the reference's vendored FAISS (oracle/_ref/liboracle_ref_ivf.so) ----------------------------------------------------
_ref_ivf = None


def ref_ivf_available():
    return os.path.exists(os.path.join(HERE, "_ref", "liboracle_ref_ivf.so"))


def ref_ivf_lib():
    global _ref_ivf
    if _ref_ivf is None:
        lib = C.CDLL(os.path.join(HERE, "_ref", "liboracle_ref_ivf.so"))
        lib.ref_ivf_last_error.restype = C.c_char_p
        lib.ref_ivf_create.restype = C.c_void_p
        lib.ref_ivf_create.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
        lib.ref_ivf_destroy.argtypes = [C.c_void_p]
        lib.ref_ivf_train_add.argtypes = [C.c_void_p, C.c_size_t, _f32p, _i64p]
        lib.ref_ivf_search.restype = C.c_int64
        lib.ref_ivf_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, _f32p, _i64p]
        lib.ref_ivf_search_batch.argtypes = [C.c_void_p, C.c_size_t, _f32p, C.c_size_t, C.c_size_t, _f32p, _i64p]
        lib.ref_ivf_range_search.restype = C.c_int64
        lib.ref_ivf_range_search.argtypes = [C.c_void_p, _f32p, C.c_float, C.c_size_t, C.c_size_t, _f32p, _i64p]
        lib.ref_ivf_export_header.argtypes = [C.c_void_p, _i64p]
        lib.ref_ivf_export.argtypes = [C.c_void_p, _f32p, _i64p, _i64p, _f32p]
        _ref_ivf = lib
    return _ref_ivf


class RefIvf:
    """faiss::IndexIVFFlat of the reference, driven like reindexer::IvfIndex (L2 / IP)."""

    def __init__(self, metric, dim, nlist):
        self.lib = ref_ivf_lib()
        self.metric, self.dim, self.nlist = metric, dim, nlist
        self.h = self.lib.ref_ivf_create(metric, dim, nlist)
        assert self.h, self.lib.ref_ivf_last_error().decode()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_ivf_destroy(self.h)
            self.h = None

    def train_add(self, labels, vecs):
        vecs = np.ascontiguousarray(vecs, np.float32)
        ids = np.ascontiguousarray(labels).astype(np.int64)
        rc = self.lib.ref_ivf_train_add(self.h, len(ids), _p(vecs, _f32p), _p(ids, _i64p))
        assert rc == 0, self.lib.ref_ivf_last_error().decode()

    def add(self, labels, vecs):
        """IvfIndex::upsert on the trained index, one add_with_ids per row"""
        vecs = np.ascontiguousarray(vecs, np.float32)
        ids = np.ascontiguousarray(labels).astype(np.int64)
        self.lib.ref_ivf_add.restype = C.c_int
        self.lib.ref_ivf_add.argtypes = [C.c_void_p, C.c_size_t, _f32p, _i64p]
        assert self.lib.ref_ivf_add(self.h, len(ids), _p(vecs, _f32p), _p(ids, _i64p)) == 0, self.lib.ref_ivf_last_error().decode()

    def remove(self, label):
        self.lib.ref_ivf_remove.restype = C.c_int
        self.lib.ref_ivf_remove.argtypes = [C.c_void_p, C.c_int64]
        assert self.lib.ref_ivf_remove(self.h, int(label)) == 0, self.lib.ref_ivf_last_error().decode()

    def list_of(self, labels):
        ids = np.ascontiguousarray(labels).astype(np.int64)
        out = np.zeros(len(ids), np.uint32)
        self.lib.ref_ivf_list_of.restype = C.c_int
        self.lib.ref_ivf_list_of.argtypes = [C.c_void_p, C.c_size_t, _i64p, _u32p]
        assert self.lib.ref_ivf_list_of(self.h, len(ids), _p(ids, _i64p), _p(out, _u32p)) == 0, self.lib.ref_ivf_last_error().decode()
        return out

    def search(self, q, k, nprobe):
        """returns (dist, label) best first; dist in FAISS' convention (L2: squared distance, IP: +inner product)"""
        q = np.ascontiguousarray(q, np.float32)
        d = np.zeros(k, np.float32)
        i = np.zeros(k, np.int64)
        n = self.lib.ref_ivf_search(self.h, _p(q, _f32p), k, nprobe, _p(d, _f32p), _p(i, _i64p))
        assert n >= 0, self.lib.ref_ivf_last_error().decode()
        return d[:n].copy(), i[:n].astype(np.uint64)

    def range_search(self, q, radius, nprobe, max_out=100000):
        """(dist, label) unsorted; radius and dist in FAISS' convention"""
        q = np.ascontiguousarray(q, np.float32)
        d = np.zeros(max_out, np.float32)
        i = np.zeros(max_out, np.int64)
        n = self.lib.ref_ivf_range_search(self.h, _p(q, _f32p), float(radius), nprobe, max_out, _p(d, _f32p), _p(i, _i64p))
        assert 0 <= n <= max_out, (n, self.lib.ref_ivf_last_error().decode())
        return d[:n].copy(), i[:n].astype(np.uint64)

    def search_batch(self, queries, k, nprobe):
        q = np.ascontiguousarray(queries, np.float32)
        d = np.zeros((len(q), k), np.float32)
        i = np.zeros((len(q), k), np.int64)
        rc = self.lib.ref_ivf_search_batch(self.h, len(q), _p(q, _f32p), k, nprobe, _p(d, _f32p), _p(i, _i64p))
        assert rc == 0, self.lib.ref_ivf_last_error().decode()
        return d, i.astype(np.uint64)

    def export(self):
        hdr = np.zeros(2, np.int64)
        assert self.lib.ref_ivf_export_header(self.h, _p(hdr, _i64p)) == 0
        nlist, ntotal = int(hdr[0]), int(hdr[1])
        cent = np.zeros((nlist, self.dim), np.float32)
        sizes = np.zeros(nlist, np.int64)
        ids = np.zeros(ntotal, np.int64)
        vecs = np.zeros((ntotal, self.dim), np.float32)
        rc = self.lib.ref_ivf_export(self.h, _p(cent, _f32p), _p(sizes, _i64p), _p(ids, _i64p), _p(vecs, _f32p))
        assert rc == 0, self.lib.ref_ivf_last_error().decode()
        return dict(centroids=cent, list_sizes=sizes.astype(np.uint64), labels=ids.astype(np.uint64), vecs=vecs)
