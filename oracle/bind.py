"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

* ``Oracle``  : oracle/liboracle.so, the own restatement (oracle/hnsw_oracle.cpp).
* ``Ref``     : oracle/_ref/libhnswref_{sse,avx2,avx512}.so, the unmodified reference headers
                compiled from /root/reference behind oracle/ref_harness.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
L2, IP = 0, 1
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(verbose=False):
    """Compile liboracle.so and (when /root/reference is present) oracle/_ref/*.so."""
    r = subprocess.run(["make", "-C", _HERE, "-j4"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_available(level="sse"):
    return os.path.exists(os.path.join(_HERE, "_ref", "libhnswref_%s.so" % level))


def best_ref_level():
    """Strongest reference build this host can run."""
    fl = cpu_flags()
    if "avx512f" in fl and "avx512dq" in fl and "avx512bw" in fl and "avx512vl" in fl and ref_available("avx512"):
        return "avx512"
    if "avx2" in fl and "fma" in fl and ref_available("avx2"):
        return "avx2"
    return "sse"


class _Lib:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)


class Ref(_Lib):
    """The real reference (hnswlib headers under /root/reference) behind ref_harness.cpp."""

    def __init__(self, level="sse"):
        super().__init__(os.path.join(_HERE, "_ref", "libhnswref_%s.so" % level))
        L = self.lib
        L.ref_last_error.restype = C.c_char_p
        L.ref_simd_level.restype = C.c_char_p
        L.ref_gen_gaussian.argtypes = [C.c_uint64, C.c_size_t, C.c_size_t, _f32p]
        L.ref_dist.restype = C.c_float
        L.ref_dist.argtypes = [C.c_int, C.c_size_t, _f32p, _f32p]
        L.ref_hnsw_new.restype = C.c_void_p
        L.ref_hnsw_new.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        L.ref_hnsw_load.restype = C.c_void_p
        L.ref_hnsw_load.argtypes = [C.c_int, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
        L.ref_hnsw_free.argtypes = [C.c_void_p]
        L.ref_hnsw_add.argtypes = [C.c_void_p, _f32p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        L.ref_hnsw_save.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_hnsw_info.argtypes = [C.c_void_p, _i64p]
        L.ref_hnsw_levels.argtypes = [C.c_void_p, _i32p]
        L.ref_hnsw_links.argtypes = [C.c_void_p, C.c_uint32, C.c_int, _u32p, C.c_int]
        L.ref_hnsw_mark_delete.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_hnsw_new_replace.restype = C.c_void_p
        L.ref_hnsw_new_replace.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
        L.ref_hnsw_add_replace_deleted.argtypes = [C.c_void_p, _f32p, C.c_void_p, C.c_size_t]
        L.ref_hnsw_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                      _u64p, _f32p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_bf_new.restype = C.c_void_p
        L.ref_bf_new.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
        L.ref_bf_load.restype = C.c_void_p
        L.ref_bf_load.argtypes = [C.c_int, C.c_size_t, C.c_char_p]
        L.ref_bf_free.argtypes = [C.c_void_p]
        L.ref_bf_add.argtypes = [C.c_void_p, _f32p, C.c_void_p, C.c_size_t]
        L.ref_bf_remove.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_bf_save.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_bf_count.restype = C.c_int64
        L.ref_bf_count.argtypes = [C.c_void_p]
        L.ref_bf_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_int, _u64p, _f32p,
                                    C.c_void_p, C.POINTER(C.c_double)]
        self.prefix = "ref"

    def simd_level(self):
        return self.lib.ref_simd_level().decode()

    def err(self):
        return self.lib.ref_last_error().decode()

    def gen_gaussian(self, seed, n, d):
        out = np.empty((n, d), np.float32)
        self.lib.ref_gen_gaussian(seed, n, d, out)
        return out

    def dist(self, metric, a, b):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        return float(self.lib.ref_dist(metric, a.size, a, b))

    def hnsw_new(self, metric, dim, max_elements, M=16, efc=200, seed=100, counting=False, allow_replace_deleted=False):
        if allow_replace_deleted:
            h = self.lib.ref_hnsw_new_replace(metric, dim, max_elements, M, efc, seed)
        else:
            h = self.lib.ref_hnsw_new(metric, dim, max_elements, M, efc, seed, int(counting))
        if not h:
            raise RuntimeError(self.err())
        return RefHnsw(self, h, dim)

    def hnsw_load(self, metric, dim, path, max_elements=0, counting=False):
        h = self.lib.ref_hnsw_load(metric, dim, path.encode(), max_elements, int(counting))
        if not h:
            raise RuntimeError(self.err())
        return RefHnsw(self, h, dim)

    def bf_new(self, metric, dim, max_elements):
        h = self.lib.ref_bf_new(metric, dim, max_elements)
        if not h:
            raise RuntimeError(self.err())
        return RefBF(self, h, dim)

    def bf_load(self, metric, dim, path):
        h = self.lib.ref_bf_load(metric, dim, path.encode())
        if not h:
            raise RuntimeError(self.err())
        return RefBF(self, h, dim)


class RefHnsw:
    def __init__(self, ref, h, dim):
        self.ref, self.h, self.dim = ref, h, dim

    def __del__(self):
        if getattr(self, "h", None):
            self.ref.lib.ref_hnsw_free(self.h)
            self.h = None

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.ref.err())

    def add(self, X, labels=None, threads=1):
        X = np.ascontiguousarray(X, np.float32)
        if labels is not None:
            labels = np.ascontiguousarray(labels, np.uint64)
        sec = C.c_double(0)
        self._chk(self.ref.lib.ref_hnsw_add(self.h, X, _opt(labels), X.shape[0], threads, C.byref(sec)))
        return sec.value

    def add_replace_deleted(self, X, labels):
        """addPoint(data, label, replace_deleted=True), serial"""
        X = np.ascontiguousarray(X, np.float32)
        labels = np.ascontiguousarray(labels, np.uint64)
        self._chk(self.ref.lib.ref_hnsw_add_replace_deleted(self.h, X, _opt(labels), X.shape[0]))

    def save(self, path):
        self._chk(self.ref.lib.ref_hnsw_save(self.h, path.encode()))

    def info(self):
        a = np.zeros(8, np.int64)
        self.ref.lib.ref_hnsw_info(self.h, a)
        keys = ["cur_element_count", "max_elements", "maxlevel", "enterpoint", "M", "maxM0", "ef_construction",
                "size_data_per_element"]
        return dict(zip(keys, (int(x) for x in a)))

    def levels(self):
        out = np.zeros(self.info()["cur_element_count"], np.int32)
        self.ref.lib.ref_hnsw_levels(self.h, out)
        return out

    def links(self, i, level):
        out = np.zeros(self.info()["maxM0"], np.uint32)
        n = self.ref.lib.ref_hnsw_links(self.h, i, level, out, out.size)
        if n < 0:
            raise IndexError((i, level))
        return out[:n].copy()

    def mark_delete(self, label):
        self._chk(self.ref.lib.ref_hnsw_mark_delete(self.h, label))

    def search(self, Q, k, ef, threads=1, counters=False):
        """-> dict(labels[nq,k] u64, dists[nq,k] f32, counts[nq], seconds, [D, Hup])."""
        Q = np.ascontiguousarray(Q, np.float32)
        nq = Q.shape[0]
        labels = np.empty((nq, k), np.uint64)
        dists = np.empty((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        D = np.zeros(nq, np.uint32) if counters else None
        Hup = np.zeros(nq, np.uint32) if counters else None
        sec = C.c_double(0)
        self._chk(self.ref.lib.ref_hnsw_search(self.h, Q, nq, k, ef, threads, labels, dists, _opt(counts), _opt(D),
                                               _opt(Hup), C.byref(sec)))
        out = dict(labels=labels, dists=dists, counts=counts, seconds=sec.value)
        if counters:
            out.update(D=D, Hup=Hup)
        return out


class RefBF:
    def __init__(self, ref, h, dim):
        self.ref, self.h, self.dim = ref, h, dim

    def __del__(self):
        if getattr(self, "h", None):
            self.ref.lib.ref_bf_free(self.h)
            self.h = None

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.ref.err())

    def add(self, X, labels=None):
        X = np.ascontiguousarray(X, np.float32)
        if labels is not None:
            labels = np.ascontiguousarray(labels, np.uint64)
        self._chk(self.ref.lib.ref_bf_add(self.h, X, _opt(labels), X.shape[0]))

    def remove(self, label):
        self._chk(self.ref.lib.ref_bf_remove(self.h, label))

    def save(self, path):
        self._chk(self.ref.lib.ref_bf_save(self.h, path.encode()))

    def count(self):
        return int(self.ref.lib.ref_bf_count(self.h))

    def search(self, Q, k, threads=1):
        Q = np.ascontiguousarray(Q, np.float32)
        nq = Q.shape[0]
        labels = np.empty((nq, k), np.uint64)
        dists = np.empty((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        sec = C.c_double(0)
        self._chk(self.ref.lib.ref_bf_search(self.h, Q, nq, k, threads, labels, dists, _opt(counts), C.byref(sec)))
        return dict(labels=labels, dists=dists, counts=counts, seconds=sec.value)


class Oracle(_Lib):
    """The own CPU restatement (oracle/hnsw_oracle.cpp)."""

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        super().__init__(path)
        L = self.lib
        L.orc_dist.restype = C.c_float
        L.orc_dist.argtypes = [C.c_int, C.c_size_t, _f32p, _f32p]
        L.orc_hnsw_new.restype = C.c_void_p
        L.orc_hnsw_new.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
        L.orc_hnsw_load.restype = C.c_void_p
        L.orc_hnsw_load.argtypes = [C.c_int, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        L.orc_hnsw_free.argtypes = [C.c_void_p]
        L.orc_hnsw_add.argtypes = [C.c_void_p, _f32p, C.c_void_p, C.c_size_t]
        L.orc_hnsw_save.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_hnsw_info.argtypes = [C.c_void_p, _i64p]
        L.orc_hnsw_levels.argtypes = [C.c_void_p, _i32p]
        L.orc_hnsw_links.argtypes = [C.c_void_p, C.c_uint32, C.c_int, _u32p, C.c_int]
        L.orc_hnsw_mark_delete.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_hnsw_allow_replace_deleted.argtypes = [C.c_void_p, C.c_int]
        L.orc_hnsw_add_replace_deleted.argtypes = [C.c_void_p, _f32p, C.c_void_p, C.c_size_t]
        L.orc_hnsw_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, _u64p, _f32p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.orc_bf_new.restype = C.c_void_p
        L.orc_bf_new.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
        L.orc_bf_free.argtypes = [C.c_void_p]
        L.orc_bf_add.argtypes = [C.c_void_p, _f32p, C.c_void_p, C.c_size_t]
        L.orc_bf_remove.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_bf_save.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_bf_count.restype = C.c_int64
        L.orc_bf_count.argtypes = [C.c_void_p]
        L.orc_bf_search.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, _u64p, _f32p, C.c_void_p,
                                    C.POINTER(C.c_double)]

    def dist(self, metric, a, b):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        return float(self.lib.orc_dist(metric, a.size, a, b))

    def hnsw_new(self, metric, dim, max_elements, M=16, efc=200, seed=100, allow_replace_deleted=False):
        h = OrcHnsw(self, self.lib.orc_hnsw_new(metric, dim, max_elements, M, efc, seed), dim)
        if allow_replace_deleted:
            self.lib.orc_hnsw_allow_replace_deleted(h.h, 1)
        return h

    def hnsw_load(self, metric, dim, path, max_elements=0, allow_replace_deleted=False):
        rc = C.c_int(0)
        h = self.lib.orc_hnsw_load(metric, dim, path.encode(), max_elements, C.byref(rc))
        if not h:
            raise RuntimeError({-1: "Cannot open file", -2: "Index seems to be corrupted or unsupported"}[rc.value])
        if allow_replace_deleted:
            self.lib.orc_hnsw_allow_replace_deleted(h, 1)
        return OrcHnsw(self, h, dim)

    def bf_new(self, metric, dim, max_elements):
        return OrcBF(self, self.lib.orc_bf_new(metric, dim, max_elements), dim)


class OrcHnsw:
    def __init__(self, orc, h, dim):
        self.orc, self.h, self.dim = orc, h, dim

    def __del__(self):
        if getattr(self, "h", None):
            self.orc.lib.orc_hnsw_free(self.h)
            self.h = None

    def add(self, X, labels=None):
        X = np.ascontiguousarray(X, np.float32)
        if labels is not None:
            labels = np.ascontiguousarray(labels, np.uint64)
        rc = self.orc.lib.orc_hnsw_add(self.h, X, _opt(labels), X.shape[0])
        if rc == -1:
            raise RuntimeError("The number of elements exceeds the specified limit")
        if rc:
            raise RuntimeError("oracle add_point rc=%d" % rc)

    def add_replace_deleted(self, X, labels):
        X = np.ascontiguousarray(X, np.float32)
        labels = np.ascontiguousarray(labels, np.uint64)
        rc = self.orc.lib.orc_hnsw_add_replace_deleted(self.h, X, _opt(labels), X.shape[0])
        if rc == -3:
            raise RuntimeError("Replacement of deleted elements is disabled in constructor")
        if rc:
            raise RuntimeError("oracle add_point_replace rc=%d" % rc)

    def save(self, path):
        if self.orc.lib.orc_hnsw_save(self.h, path.encode()):
            raise RuntimeError("cannot write " + path)

    def info(self):
        a = np.zeros(8, np.int64)
        self.orc.lib.orc_hnsw_info(self.h, a)
        keys = ["cur_element_count", "max_elements", "maxlevel", "enterpoint", "M", "maxM0", "ef_construction",
                "size_data_per_element"]
        return dict(zip(keys, (int(x) for x in a)))

    def levels(self):
        out = np.zeros(self.info()["cur_element_count"], np.int32)
        self.orc.lib.orc_hnsw_levels(self.h, out)
        return out

    def links(self, i, level):
        out = np.zeros(self.info()["maxM0"], np.uint32)
        n = self.orc.lib.orc_hnsw_links(self.h, i, level, out, out.size)
        if n < 0:
            raise IndexError((i, level))
        return out[:n].copy()

    def mark_delete(self, label):
        rc = self.orc.lib.orc_hnsw_mark_delete(self.h, label)
        if rc:
            raise RuntimeError("Label not found" if rc == -1 else "The requested to delete element is already deleted")

    def search(self, Q, k, ef):
        """Serial search with work counters -> dict(labels, dists, counts, D, H0, Hup, seconds)."""
        Q = np.ascontiguousarray(Q, np.float32)
        nq = Q.shape[0]
        labels = np.empty((nq, k), np.uint64)
        dists = np.empty((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        D = np.zeros(nq, np.uint32)
        H0 = np.zeros(nq, np.uint32)
        Hup = np.zeros(nq, np.uint32)
        sec = C.c_double(0)
        self.orc.lib.orc_hnsw_search(self.h, Q, nq, k, ef, labels, dists, _opt(counts), _opt(D), _opt(H0), _opt(Hup),
                                     C.byref(sec))
        return dict(labels=labels, dists=dists, counts=counts, D=D, H0=H0, Hup=Hup, seconds=sec.value)


class OrcBF:
    def __init__(self, orc, h, dim):
        self.orc, self.h, self.dim = orc, h, dim

    def __del__(self):
        if getattr(self, "h", None):
            self.orc.lib.orc_bf_free(self.h)
            self.h = None

    def add(self, X, labels=None):
        X = np.ascontiguousarray(X, np.float32)
        if labels is not None:
            labels = np.ascontiguousarray(labels, np.uint64)
        if self.orc.lib.orc_bf_add(self.h, X, _opt(labels), X.shape[0]):
            raise RuntimeError("The number of elements exceeds the specified limit\n")

    def remove(self, label):
        self.orc.lib.orc_bf_remove(self.h, label)

    def save(self, path):
        if self.orc.lib.orc_bf_save(self.h, path.encode()):
            raise RuntimeError("cannot write " + path)

    def count(self):
        return int(self.orc.lib.orc_bf_count(self.h))

    def search(self, Q, k):
        Q = np.ascontiguousarray(Q, np.float32)
        nq = Q.shape[0]
        labels = np.empty((nq, k), np.uint64)
        dists = np.empty((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        sec = C.c_double(0)
        self.orc.lib.orc_bf_search(self.h, Q, nq, k, labels, dists, _opt(counts), C.byref(sec))
        return dict(labels=labels, dists=dists, counts=counts, seconds=sec.value)


# Synthetic data laws live with the product (research_new_hnsw_b200/synth.py); re-exported for the tests.
def lowrank_data(*args, **kw):
    from research_new_hnsw_b200.synth import lowrank_data as f
    return f(*args, **kw)
