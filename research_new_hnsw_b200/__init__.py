"""research_new_hnsw_b200 -- B200-native (sm_100a) HNSW / brute-force engine behind the hnswlib API.

The product is ``libb200hnsw.so`` (C ABI: ``include/b200hnsw.h``) plus the drop-in C++ header shim under
``research_new_hnsw_b200/hnswlib/``.  This Python package is a thin ctypes mirror of the same interface
(``HierarchicalNSW``, ``BruteforceSearch``, ``L2Space``, ``InnerProductSpace`` with the reference's method names)
used by the tests and the benchmark; it contains no compute and there is no CPU fallback: every search/build call
runs hand-written CUDA kernels and raises if the library or a GPU is missing.
"""
from .capi import (  # noqa: F401
    B200Error,
    BruteforceSearch,
    HierarchicalNSW,
    InnerProductSpace,
    L2Space,
    P2PExchange,
    ShardedHierarchicalNSW,
    build_library,
    device_count,
    lib_path,
    load_library,
    merge_topk_device,
    merge_topk_packed_device,
)

__all__ = [
    "B200Error", "BruteforceSearch", "HierarchicalNSW", "InnerProductSpace", "L2Space", "P2PExchange", "ShardedHierarchicalNSW", "build_library",
    "device_count", "lib_path", "load_library", "merge_topk_device", "merge_topk_packed_device",
]
