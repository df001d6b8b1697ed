"""CPU: the C-ABI library builds, loads and exports every symbol include/b200hnsw.h declares; without a GPU the
product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "b200hnsw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200(?:hnsw|bf)_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 30
    L = ctypes.CDLL(lib.lib_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    from research_new_hnsw_b200 import capi
    assert sorted(capi.EXPORTS) == names
    assert L.b200hnsw_abi_version() == 2


def test_no_cpu_fallback(lib):
    """On a box without a CUDA device every constructor fails with the CUDA status -- nothing runs on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.B200Error) as e:
        lib.HierarchicalNSW(lib.L2Space(16), 100)
    assert e.value.code == -1 and "no CPU fallback" in str(e.value)
    with pytest.raises(lib.B200Error):
        lib.BruteforceSearch(lib.L2Space(16), 100)
    with pytest.raises(lib.B200Error):
        lib.HierarchicalNSW(lib.L2Space(16), os.path.join(ROOT, "tests", "golden", "l2_n2000_d16_M8.bin"))


def test_load_errors_keep_reference_messages(lib, tmp_path):
    with pytest.raises(lib.B200Error, match="Cannot open file"):
        lib.HierarchicalNSW(lib.L2Space(16), str(tmp_path / "nope.bin"))
    data = open(os.path.join(ROOT, "tests", "golden", "l2_n2000_d16_M8.bin"), "rb").read()
    p = tmp_path / "trunc.bin"
    p.write_bytes(data[:-5])
    with pytest.raises(lib.B200Error, match="Index seems to be corrupted or unsupported"):
        lib.HierarchicalNSW(lib.L2Space(16), str(p))


def test_product_does_not_touch_oracle():
    """The shipped path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "research_new_hnsw_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), (dirpath, f)


def test_headers_compile_standalone(tmp_path):
    """The boundary is a C ABI: include/b200hnsw.h is valid C99 (plain pointers and sizes, no C++ or torch types), and the
    drop-in header directory compiles on its own as C++17 without CUDA headers."""
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "b200hnsw.h"\nint main(void) { return (int)sizeof(b200hnsw_params) == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I" + os.path.join(ROOT, "include"), str(c)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "hnswlib/hnswlib.h"\n#include "hnswlib/stop_condition.h"\n#include "hnswlib/hnswalg.h"\n'
                   '#include "hnswlib/bruteforce.h"\n#include "hnswlib/space_l2.h"\n#include "hnswlib/space_ip.h"\n'
                   'int main() { hnswlib::L2Space a(8); hnswlib::InnerProductSpace b(8); (void)a; (void)b; return 0; }\n')
    r = subprocess.run(["g++", "-std=gnu++17", "-Wall", "-fsyntax-only", "-I" + os.path.join(ROOT, "research_new_hnsw_b200"),
                        "-I" + os.path.join(ROOT, "include"), str(cpp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
