"""ctypes mirror of include/b200hnsw.h with the reference's class and method names.

Reference interface mirrored (file:line under /root/reference/hnswlib):
  L2Space / InnerProductSpace         space_l2.h:207-253, space_ip.h:343-398
  HierarchicalNSW<float>              hnswalg.h:17 (ctor overloads :78-144, setEf :173, addPoint :954,
                                      searchKnn :1270, saveIndex :685, loadIndex :716, markDelete :853, ...)
  BruteforceSearch<float>             bruteforce.h:10-172
Error behaviour: every reference ``std::runtime_error`` message is raised as ``B200Error`` with the same text.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

L2, IP = 0, 1
F32, BF16 = 0, 1


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class _Params(C.Structure):
    _fields_ = [("metric", C.c_int32), ("storage", C.c_int32), ("device", C.c_int32),
                ("allow_replace_deleted", C.c_int32), ("dim", C.c_uint64), ("max_elements", C.c_uint64),
                ("M", C.c_uint64), ("ef_construction", C.c_uint64), ("random_seed", C.c_uint64)]


class _Info(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("cur_element_count", "max_elements", "num_deleted", "dim", "M", "maxM", "maxM0", "ef_construction",
                 "ef", "size_data_per_element", "size_links_per_element", "size_links_level0", "offset_data",
                 "label_offset")] + [("mult", C.c_double), ("maxlevel", C.c_int32), ("enterpoint_node", C.c_uint32),
                                     ("metric", C.c_int32), ("storage", C.c_int32), ("device", C.c_int32),
                                     ("reserved", C.c_int32)]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("queries", "dist_evals", "hops_base", "hops_upper", "visited_resets",
                                          "kernel_launches")] + [("last_kernel_ms", C.c_double),
                                                                 ("dropped_reverse_edges", C.c_uint64)]


EXPORTS = [
    "b200hnsw_last_error", "b200hnsw_abi_version", "b200hnsw_device_count", "b200hnsw_create", "b200hnsw_load",
    "b200hnsw_save", "b200hnsw_destroy", "b200hnsw_set_ef", "b200hnsw_add_batch", "b200hnsw_add_batch_replace_deleted", "b200hnsw_flush",
    "b200hnsw_search_batch", "b200hnsw_search_batch_submit", "b200hnsw_search_batch_wait", "b200hnsw_search_batch_filtered", "b200hnsw_get_labels", "b200hnsw_search_batch_device", "b200hnsw_get_info", "b200hnsw_get_levels",
    "b200hnsw_get_linklist", "b200hnsw_get_label", "b200hnsw_get_data", "b200hnsw_get_data_by_label",
    "b200hnsw_mark_delete", "b200hnsw_unmark_delete", "b200hnsw_resize", "b200hnsw_index_file_size",
    "b200hnsw_get_stats", "b200hnsw_sharded_create", "b200hnsw_sharded_load", "b200hnsw_sharded_save",
    "b200hnsw_sharded_destroy", "b200hnsw_sharded_num_shards", "b200hnsw_sharded_get_shard", "b200hnsw_sharded_count",
    "b200hnsw_sharded_add_batch", "b200hnsw_sharded_flush", "b200hnsw_sharded_search_batch", "b200hnsw_sharded_last_ms",
    "b200hnsw_exchange_create", "b200hnsw_exchange_connect", "b200hnsw_exchange_slot", "b200hnsw_exchange_step",
    "b200hnsw_exchange_destroy", "b200hnsw_merge_topk_device", "b200hnsw_merge_topk_packed_device", "b200bf_create", "b200bf_load", "b200bf_save",
    "b200bf_destroy", "b200bf_add_batch", "b200bf_remove", "b200bf_search_batch", "b200bf_search_batch_device",
    "b200bf_search_batch_filtered", "b200bf_get_labels", "b200bf_count", "b200bf_get_stats",
]


def lib_path():
    # B200HNSW_LIB: alternative build of the same library (kernel-tuning experiments); default is the in-tree build
    return os.environ.get("B200HNSW_LIB") or os.path.join(_HERE, "libb200hnsw.so")


def build_library(verbose=False):
    """Compile libb200hnsw.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("libb200hnsw.so build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return lib_path()


def load_library():
    """Load the CUDA library; fail loudly if it is missing (there is no CPU fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError("libb200hnsw.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`"
                          % path)
    L = C.CDLL(path)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
    L.b200hnsw_last_error.restype = C.c_char_p
    L.b200hnsw_create.argtypes = [C.POINTER(_Params), C.POINTER(vp)]
    L.b200hnsw_load.argtypes = [C.c_char_p, C.POINTER(_Params), C.POINTER(vp)]
    L.b200hnsw_save.argtypes = [vp, C.c_char_p]
    L.b200hnsw_destroy.argtypes = [vp]
    L.b200hnsw_destroy.restype = None
    L.b200hnsw_set_ef.argtypes = [vp, sz]
    L.b200hnsw_add_batch.argtypes = [vp, vp, vp, sz]
    L.b200hnsw_add_batch_replace_deleted.argtypes = [vp, vp, vp, sz]
    L.b200hnsw_flush.argtypes = [vp]
    L.b200hnsw_search_batch.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp, vp]
    L.b200hnsw_search_batch_submit.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp, C.POINTER(C.c_uint64)]
    L.b200hnsw_search_batch_wait.argtypes = [vp, C.c_uint64]
    L.b200hnsw_search_batch_device.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp, vp, vp]
    L.b200hnsw_search_batch_filtered.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp, vp]
    L.b200hnsw_get_labels.argtypes = [vp, vp, sz]
    L.b200hnsw_get_info.argtypes = [vp, C.POINTER(_Info)]
    L.b200hnsw_get_levels.argtypes = [vp, C.POINTER(C.POINTER(C.c_int32))]
    L.b200hnsw_get_linklist.argtypes = [vp, u32, i32, C.POINTER(C.POINTER(C.c_uint32))]
    L.b200hnsw_get_label.argtypes = [vp, u32, C.POINTER(C.c_uint64)]
    L.b200hnsw_get_data.argtypes = [vp, u32, C.POINTER(C.POINTER(C.c_float))]
    L.b200hnsw_get_data_by_label.argtypes = [vp, C.c_uint64, vp]
    L.b200hnsw_mark_delete.argtypes = [vp, C.c_uint64]
    L.b200hnsw_unmark_delete.argtypes = [vp, C.c_uint64]
    L.b200hnsw_resize.argtypes = [vp, sz]
    L.b200hnsw_index_file_size.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.b200hnsw_get_stats.argtypes = [vp, C.POINTER(_Stats)]
    L.b200hnsw_sharded_create.argtypes = [C.POINTER(_Params), C.POINTER(C.c_int), sz, C.POINTER(vp)]
    L.b200hnsw_sharded_load.argtypes = [C.POINTER(C.c_char_p), C.POINTER(_Params), C.POINTER(C.c_int), sz, C.POINTER(vp)]
    L.b200hnsw_sharded_save.argtypes = [vp, C.POINTER(C.c_char_p)]
    L.b200hnsw_sharded_destroy.argtypes = [vp]
    L.b200hnsw_sharded_destroy.restype = None
    L.b200hnsw_sharded_num_shards.argtypes = [vp, C.POINTER(sz)]
    L.b200hnsw_sharded_get_shard.argtypes = [vp, sz, C.POINTER(vp)]
    L.b200hnsw_sharded_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.b200hnsw_sharded_add_batch.argtypes = [vp, vp, vp, sz]
    L.b200hnsw_sharded_flush.argtypes = [vp]
    L.b200hnsw_sharded_search_batch.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
    L.b200hnsw_sharded_last_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.b200hnsw_exchange_create.argtypes = [C.c_int, sz, sz, sz, C.POINTER(vp), vp]
    L.b200hnsw_exchange_connect.argtypes = [vp, vp]
    L.b200hnsw_exchange_slot.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.POINTER(vp)]
    L.b200hnsw_exchange_step.argtypes = [vp, C.c_uint32, vp]
    L.b200hnsw_exchange_destroy.argtypes = [vp]
    L.b200hnsw_exchange_destroy.restype = None
    L.b200hnsw_merge_topk_device.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
    L.b200hnsw_merge_topk_packed_device.argtypes = [vp, sz, sz, sz, sz, vp, vp, vp]
    L.b200bf_create.argtypes = [C.POINTER(_Params), C.POINTER(vp)]
    L.b200bf_load.argtypes = [C.c_char_p, C.POINTER(_Params), C.POINTER(vp)]
    L.b200bf_save.argtypes = [vp, C.c_char_p]
    L.b200bf_destroy.argtypes = [vp]
    L.b200bf_destroy.restype = None
    L.b200bf_add_batch.argtypes = [vp, vp, vp, sz]
    L.b200bf_remove.argtypes = [vp, C.c_uint64]
    L.b200bf_search_batch.argtypes = [vp, vp, sz, sz, vp, vp, vp]
    L.b200bf_search_batch_device.argtypes = [vp, vp, sz, sz, vp, vp, vp, vp]
    L.b200bf_search_batch_filtered.argtypes = [vp, vp, sz, sz, vp, vp, vp, vp]
    L.b200bf_get_labels.argtypes = [vp, vp, sz]
    L.b200bf_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.b200bf_get_stats.argtypes = [vp, C.POINTER(_Stats)]
    _LIB = L
    return L


def _chk(rc):
    if rc != 0:
        raise B200Error(rc, load_library().b200hnsw_last_error().decode())


def device_count():
    n = load_library().b200hnsw_device_count()
    if n < 0:
        _chk(n)
    return n


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class _Space:
    metric = L2

    def __init__(self, dim):
        self.dim = int(dim)

    def get_data_size(self):  # space_l2.h:240-242
        return self.dim * 4


class L2Space(_Space):
    metric = L2


class InnerProductSpace(_Space):
    metric = IP


def _params(space, max_elements=0, M=16, ef_construction=200, random_seed=100, allow_replace_deleted=False,
            storage=F32, device=-1):
    return _Params(space.metric, storage, device, int(allow_replace_deleted), space.dim, max_elements, M,
                   ef_construction, random_seed)


class HierarchicalNSW:
    """hnswlib::HierarchicalNSW<float> (hnswalg.h:17).  ``HierarchicalNSW(space, location)`` loads,
    ``HierarchicalNSW(space, max_elements, M, ef_construction, random_seed)`` builds."""

    def __init__(self, space, arg, M=16, ef_construction=200, random_seed=100, allow_replace_deleted=False,
                 max_elements=0, storage=F32, device=-1):
        self._L = load_library()
        self.space = space
        self._h = C.c_void_p()
        if isinstance(arg, (str, bytes, os.PathLike)):
            p = _params(space, max_elements, 0, 0, 0, allow_replace_deleted, storage, device)
            _chk(self._L.b200hnsw_load(os.fsencode(arg), C.byref(p), C.byref(self._h)))
        else:
            p = _params(space, int(arg), M, ef_construction, random_seed, allow_replace_deleted, storage, device)
            _chk(self._L.b200hnsw_create(C.byref(p), C.byref(self._h)))

    @classmethod
    def _borrowed(cls, space, handle, owner):
        """view of a shard owned by a ShardedHierarchicalNSW (never destroyed from here)"""
        self = cls.__new__(cls)
        self._L, self.space, self._h, self._owner = load_library(), space, handle, owner
        return self

    def __del__(self):
        if getattr(self, "_h", None) and not getattr(self, "_owner", None):
            self._L.b200hnsw_destroy(self._h)
        self._h = None

    # ---- public fields of the reference class -------------------------------------------------
    def info(self):
        o = _Info()
        _chk(self._L.b200hnsw_get_info(self._h, C.byref(o)))
        return {n: getattr(o, n) for n, _ in _Info._fields_ if n != "reserved"}

    @property
    def cur_element_count(self):
        return self.info()["cur_element_count"]

    @property
    def maxlevel_(self):
        return self.info()["maxlevel"]

    @property
    def enterpoint_node_(self):
        return self.info()["enterpoint_node"]

    @property
    def element_levels_(self):
        p = C.POINTER(C.c_int32)()
        _chk(self._L.b200hnsw_get_levels(self._h, C.byref(p)))
        n = self.cur_element_count
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.int32)

    # ---- methods ------------------------------------------------------------------------------
    def setEf(self, ef):
        _chk(self._L.b200hnsw_set_ef(self._h, ef))

    def addPoint(self, datapoint, label, replace_deleted=False):
        self.addPoints(np.asarray(datapoint, np.float32).reshape(1, -1), np.array([label], np.uint64), replace_deleted)

    def addPoints(self, X, labels=None, replace_deleted=False):
        X = np.ascontiguousarray(X, np.float32)
        assert X.ndim == 2 and X.shape[1] == self.space.dim
        if labels is not None:
            labels = np.ascontiguousarray(labels, np.uint64)
        fn = self._L.b200hnsw_add_batch_replace_deleted if replace_deleted else self._L.b200hnsw_add_batch
        _chk(fn(self._h, _ptr(X), _ptr(labels), X.shape[0]))

    def flush(self):
        _chk(self._L.b200hnsw_flush(self._h))

    def saveIndex(self, location):
        _chk(self._L.b200hnsw_save(self._h, os.fsencode(location)))

    def searchKnn(self, query, k):
        """One query -> list of (dist, label), FURTHEST first like the reference's max-heap pops (main.cpp:71-75)."""
        r = self.searchKnnBatch(np.asarray(query, np.float32).reshape(1, -1), k)
        n = int(r["counts"][0])
        return [(float(r["dists"][0, j]), int(r["labels"][0, j])) for j in range(n - 1, -1, -1)]

    def searchKnnCloserFirst(self, query, k):  # hnswlib.h:205-225
        return self.searchKnn(query, k)[::-1]

    def searchKnnBatch(self, Q, k, ef=0, work=False, out=None):
        """Batched searchKnn through the host-pointer C ABI -> dict(labels[nq,k], dists[nq,k], counts[nq], [work]).
        `out` may carry preallocated (e.g. page-locked) `labels`, `dists`, `counts` arrays; with page-locked Q and
        outputs the library overlaps the copies with the kernels of a large batch."""
        Q = np.ascontiguousarray(Q, np.float32)
        assert Q.ndim == 2 and Q.shape[1] == self.space.dim
        nq = Q.shape[0]
        if out is not None:
            labels, dists, counts = out["labels"], out["dists"], out["counts"]
            assert labels.shape == (nq, k) and labels.dtype == np.uint64 and labels.flags.c_contiguous
            assert dists.shape == (nq, k) and dists.dtype == np.float32 and dists.flags.c_contiguous
            assert counts.shape == (nq,) and counts.dtype == np.uint32
        else:
            labels = np.empty((nq, k), np.uint64)
            dists = np.empty((nq, k), np.float32)
            counts = np.zeros(nq, np.uint32)
        w = np.zeros((nq, 4), np.uint32) if work else None
        _chk(self._L.b200hnsw_search_batch(self._h, _ptr(Q), nq, k, ef, _ptr(labels), _ptr(dists), _ptr(counts),
                                           _ptr(w)))
        out = dict(labels=labels, dists=dists, counts=counts)
        if work:
            out.update(D=w[:, 0].copy(), H0=w[:, 1].copy(), Hup=w[:, 2].copy(), resets=w[:, 3].copy())
        return out

    def searchKnnBatchSubmit(self, Q, k, out, ef=0):
        """Asynchronous searchKnnBatch (b200hnsw_search_batch_submit): `Q` and out["labels"/"dists"/"counts"] must be
        page-locked numpy views that stay alive and untouched until searchKnnBatchWait(ticket) returns."""
        assert Q.dtype == np.float32 and Q.flags.c_contiguous and Q.shape[1] == self.space.dim
        nq = Q.shape[0]
        labels, dists, counts = out["labels"], out["dists"], out["counts"]
        assert labels.shape == (nq, k) and labels.dtype == np.uint64 and dists.shape == (nq, k) and dists.dtype == np.float32
        t = C.c_uint64()
        _chk(self._L.b200hnsw_search_batch_submit(self._h, _ptr(Q), nq, k, ef, _ptr(labels), _ptr(dists), _ptr(counts),
                                                  C.byref(t)))
        return t.value

    def searchKnnBatchWait(self, ticket):
        _chk(self._L.b200hnsw_search_batch_wait(self._h, ticket))

    def searchKnnFiltered(self, Q, k, is_id_allowed, ef=0):
        """searchKnn(query, k, isIdAllowed) batched: `is_id_allowed(label) -> bool` plays BaseFilterFunctor
        (hnswlib.h:128-132); it is evaluated once per stored label on the host."""
        Q = np.ascontiguousarray(Q, np.float32)
        nq, n = Q.shape[0], self.cur_element_count
        lab = np.empty(max(n, 1), np.uint64)
        _chk(self._L.b200hnsw_get_labels(self._h, _ptr(lab), lab.size))
        allowed = np.fromiter((1 if is_id_allowed(int(l)) else 0 for l in lab[:n]), np.uint8, n)
        labels = np.empty((nq, k), np.uint64)
        dists = np.empty((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        _chk(self._L.b200hnsw_search_batch_filtered(self._h, _ptr(Q), nq, k, ef, _ptr(allowed), _ptr(labels),
                                                    _ptr(dists), _ptr(counts)))
        return dict(labels=labels, dists=dists, counts=counts)

    def searchKnnDevice(self, dQ, nq, k, ef, d_labels, d_dists, d_counts=0, d_work=0, stream=0):
        """Device-pointer search (raw addresses, e.g. torch ``tensor.data_ptr()``), asynchronous on ``stream``."""
        _chk(self._L.b200hnsw_search_batch_device(self._h, dQ, nq, k, ef, d_labels, d_dists, d_counts or None,
                                                  d_work or None, stream or None))

    def get_linklist_at_level(self, internal_id, level):
        """-> neighbour ids (hnswalg.h:501-503 + getListCount :940)."""
        p = C.POINTER(C.c_uint32)()
        _chk(self._L.b200hnsw_get_linklist(self._h, internal_id, level, C.byref(p)))
        cnt = p[0] & 0xFFFF
        return np.array([p[1 + j] for j in range(cnt)], np.uint32)

    def getExternalLabel(self, internal_id):
        v = C.c_uint64()
        _chk(self._L.b200hnsw_get_label(self._h, internal_id, C.byref(v)))
        return v.value

    def getDataByLabel(self, label):
        out = np.empty(self.space.dim, np.float32)
        _chk(self._L.b200hnsw_get_data_by_label(self._h, label, _ptr(out)))
        return out

    def markDelete(self, label):
        _chk(self._L.b200hnsw_mark_delete(self._h, label))

    def unmarkDelete(self, label):
        _chk(self._L.b200hnsw_unmark_delete(self._h, label))

    def resizeIndex(self, new_max_elements):
        _chk(self._L.b200hnsw_resize(self._h, new_max_elements))

    def indexFileSize(self):
        v = C.c_uint64()
        _chk(self._L.b200hnsw_index_file_size(self._h, C.byref(v)))
        return v.value

    def getMaxElements(self):
        return self.info()["max_elements"]

    def getCurrentElementCount(self):
        return self.info()["cur_element_count"]

    def getDeletedCount(self):
        return self.info()["num_deleted"]

    def stats(self):
        s = _Stats()
        _chk(self._L.b200hnsw_get_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in _Stats._fields_}


class ShardedHierarchicalNSW:
    """One process, one HierarchicalNSW sub-index per device (include/b200hnsw.h, b200hnsw_sharded_*):
    every query is searched on every shard, per-shard top-k are merged on devices[0].  ``arg`` is the per-shard capacity
    (build) or a list of per-shard saveIndex files (load)."""

    def __init__(self, space, arg, devices, M=16, ef_construction=200, random_seed=100, storage=F32):
        self._L = load_library()
        self.space = space
        self._h = C.c_void_p()
        dv = (C.c_int * len(devices))(*devices)
        if isinstance(arg, (list, tuple)):
            assert len(arg) == len(devices)
            p = _params(space, 0, 0, 0, 0, False, storage)
            paths = (C.c_char_p * len(arg))(*[os.fsencode(a) for a in arg])
            _chk(self._L.b200hnsw_sharded_load(paths, C.byref(p), dv, len(devices), C.byref(self._h)))
        else:
            p = _params(space, int(arg), M, ef_construction, random_seed, False, storage)
            _chk(self._L.b200hnsw_sharded_create(C.byref(p), dv, len(devices), C.byref(self._h)))
        self.n_shards = len(devices)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200hnsw_sharded_destroy(self._h)
            self._h = None

    def shard(self, s):
        h = C.c_void_p()
        _chk(self._L.b200hnsw_sharded_get_shard(self._h, s, C.byref(h)))
        return HierarchicalNSW._borrowed(self.space, h, self)

    @property
    def cur_element_count(self):
        v = C.c_uint64()
        _chk(self._L.b200hnsw_sharded_count(self._h, C.byref(v)))
        return v.value

    def addPoints(self, X, labels=None):
        X = np.ascontiguousarray(X, np.float32)
        assert X.ndim == 2 and X.shape[1] == self.space.dim
        if labels is not None:
            labels = np.ascontiguousarray(labels, np.uint64)
        _chk(self._L.b200hnsw_sharded_add_batch(self._h, _ptr(X), _ptr(labels), X.shape[0]))

    def flush(self):
        _chk(self._L.b200hnsw_sharded_flush(self._h))

    def saveIndex(self, locations):
        paths = (C.c_char_p * len(locations))(*[os.fsencode(a) for a in locations])
        _chk(self._L.b200hnsw_sharded_save(self._h, paths))

    def searchKnnBatch(self, Q, k, ef=0):
        Q = np.ascontiguousarray(Q, np.float32)
        assert Q.ndim == 2 and Q.shape[1] == self.space.dim
        nq = Q.shape[0]
        labels = np.empty((nq, k), np.uint64)
        dists = np.empty((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        _chk(self._L.b200hnsw_sharded_search_batch(self._h, _ptr(Q), nq, k, ef, _ptr(labels), _ptr(dists), _ptr(counts)))
        return dict(labels=labels, dists=dists, counts=counts)

    def last_ms(self):
        v = C.c_double()
        _chk(self._L.b200hnsw_sharded_last_ms(self._h, C.byref(v)))
        return v.value


class BruteforceSearch:
    """hnswlib::BruteforceSearch<float> (bruteforce.h:10)."""

    def __init__(self, space, arg, device=-1):
        self._L = load_library()
        self.space = space
        self._h = C.c_void_p()
        if isinstance(arg, (str, bytes, os.PathLike)):
            p = _params(space, device=device)
            _chk(self._L.b200bf_load(os.fsencode(arg), C.byref(p), C.byref(self._h)))
        else:
            p = _params(space, int(arg), device=device)
            _chk(self._L.b200bf_create(C.byref(p), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200bf_destroy(self._h)
            self._h = None

    @property
    def cur_element_count(self):
        v = C.c_uint64()
        _chk(self._L.b200bf_count(self._h, C.byref(v)))
        return v.value

    def addPoint(self, datapoint, label):
        self.addPoints(np.asarray(datapoint, np.float32).reshape(1, -1), np.array([label], np.uint64))

    def addPoints(self, X, labels=None):
        X = np.ascontiguousarray(X, np.float32)
        assert X.ndim == 2 and X.shape[1] == self.space.dim
        if labels is not None:
            labels = np.ascontiguousarray(labels, np.uint64)
        _chk(self._L.b200bf_add_batch(self._h, _ptr(X), _ptr(labels), X.shape[0]))

    def removePoint(self, label):
        _chk(self._L.b200bf_remove(self._h, label))

    def saveIndex(self, location):
        _chk(self._L.b200bf_save(self._h, os.fsencode(location)))

    def searchKnnBatch(self, Q, k):
        Q = np.ascontiguousarray(Q, np.float32)
        assert Q.ndim == 2 and Q.shape[1] == self.space.dim
        nq = Q.shape[0]
        labels = np.empty((nq, k), np.uint64)
        dists = np.empty((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        _chk(self._L.b200bf_search_batch(self._h, _ptr(Q), nq, k, _ptr(labels), _ptr(dists), _ptr(counts)))
        return dict(labels=labels, dists=dists, counts=counts)

    def searchKnn(self, query, k):
        r = self.searchKnnBatch(np.asarray(query, np.float32).reshape(1, -1), k)
        n = int(r["counts"][0])
        return [(float(r["dists"][0, j]), int(r["labels"][0, j])) for j in range(n - 1, -1, -1)]

    def searchKnnFiltered(self, Q, k, is_id_allowed):
        """searchKnn(query, k, isIdAllowed) batched (bruteforce.h:106-135): `is_id_allowed(label) -> bool` plays
        BaseFilterFunctor; it is evaluated once per stored row on the host, the kernels take the verdicts as a row mask."""
        Q = np.ascontiguousarray(Q, np.float32)
        nq, n = Q.shape[0], self.cur_element_count
        lab = np.empty(max(n, 1), np.uint64)
        _chk(self._L.b200bf_get_labels(self._h, _ptr(lab), lab.size))
        allowed = np.fromiter((1 if is_id_allowed(int(l)) else 0 for l in lab[:n]), np.uint8, n)
        labels = np.empty((nq, k), np.uint64)
        dists = np.empty((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        _chk(self._L.b200bf_search_batch_filtered(self._h, _ptr(Q), nq, k, _ptr(allowed), _ptr(labels), _ptr(dists),
                                                  _ptr(counts)))
        return dict(labels=labels, dists=dists, counts=counts)

    def searchKnnDevice(self, dQ, nq, k, d_labels, d_dists, d_counts=0, stream=0):
        _chk(self._L.b200bf_search_batch_device(self._h, dQ, nq, k, d_labels, d_dists, d_counts or None,
                                                stream or None))

    def stats(self):
        s = _Stats()
        _chk(self._L.b200bf_get_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in _Stats._fields_}


EXCHANGE_DESC_BYTES = 128


class P2PExchange:
    """b200hnsw_exchange_*: result blocks pushed peer to peer by the copy engines, stream memory operations as flags
    (csrc/exchange.cu).  `gather_descs(my_desc: bytes) -> bytes` must return the descriptors of all ranks in rank order."""

    def __init__(self, device, world, rank, block_bytes, gather_descs):
        self._L = load_library()
        self._h = C.c_void_p()
        desc = C.create_string_buffer(EXCHANGE_DESC_BYTES)
        _chk(self._L.b200hnsw_exchange_create(device, world, rank, block_bytes, C.byref(self._h), desc))
        alld = gather_descs(desc.raw)
        assert len(alld) == world * EXCHANGE_DESC_BYTES
        buf = C.create_string_buffer(alld, len(alld))
        _chk(self._L.b200hnsw_exchange_connect(self._h, buf))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200hnsw_exchange_destroy(self._h)
            self._h = None

    def slot(self, step):
        """-> (address of this rank's block, address of the [world][block] area) for `step` (numbered from 1)"""
        mine, allb = C.c_void_p(), C.c_void_p()
        _chk(self._L.b200hnsw_exchange_slot(self._h, step, C.byref(mine), C.byref(allb)))
        return mine.value, allb.value

    def step(self, step, stream=0):
        _chk(self._L.b200hnsw_exchange_step(self._h, step, stream or None))


def merge_topk_device(d_labels_in, d_dists_in, shards, nq, k, d_labels_out, d_dists_out, stream=0):
    """Merge [shards][nq][k] per-shard results (device pointers) into [nq][k]."""
    _chk(load_library().b200hnsw_merge_topk_device(d_labels_in, d_dists_in, shards, nq, k, d_labels_out, d_dists_out,
                                                   stream or None))


def merge_topk_packed_device(d_blocks, block_bytes, shards, nq, k, d_labels_out, d_dists_out, stream=0):
    """Merge per-shard blocks packed as [labels | dists] (one all_gather moves both arrays)."""
    _chk(load_library().b200hnsw_merge_topk_packed_device(d_blocks, block_bytes, shards, nq, k, d_labels_out,
                                                          d_dists_out, stream or None))
