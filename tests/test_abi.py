"""CPU tests: the C-ABI library loads without a GPU, exports every symbol include/rxgpu.h declares, and refuses to compute
without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import reindexer_b200 as rx
from reindexer_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rxgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rxgpu_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(binding.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/rxgpu.h but not exported by librxgpu.so"
    assert set(syms) == set(binding._SIGNATURES), set(syms) ^ set(binding._SIGNATURES)
    assert rx.lib().rxgpu_abi_version() == 4


def test_library_does_not_link_oracle_or_libcuda():
    import subprocess

    out = subprocess.check_output(["ldd", binding.LIB_PATH]).decode()
    assert "oracle" not in out and "libcuda.so" not in out
    syms = subprocess.check_output(["nm", "-D", "--defined-only", binding.LIB_PATH]).decode()
    assert "port_bf" not in syms and "ref_bf" not in syms


@pytest.mark.skipif(rx.device_count() > 0, reason="box has a GPU")
def test_no_cpu_fallback_without_device():
    with pytest.raises(rx.RxGpuError) as e:
        rx.GpuBruteforceSearch(rx.L2, 16, 100)
    assert e.value.code == 37 and "no CPU fallback" in e.value.what
    with pytest.raises(rx.RxGpuError) as e:  # the in-process communicators of the sharded ft_fast merge need their devices as well
        rx.ShardComm.local_group(2)
    assert e.value.code == 37 and "no CPU fallback" in e.value.what


def test_product_package_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "reindexer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                body = open(os.path.join(dirpath, f)).read()
                code = "\n".join(l for l in body.splitlines() if not l.strip().startswith(("//", "#", "*", "/*", '"""')))
                assert "import oracle" not in code and "from oracle" not in code and "liboracle" not in code, os.path.join(dirpath, f)


def test_header_is_plain_c():
    """the boundary is a C ABI: include/rxgpu.h must compile as C99 with no C++ or torch types in it"""
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "hdr_check.c")
        with open(src, "w") as f:
            f.write('#include "rxgpu.h"\nint main(void) { rxgpu_search_stats s; (void)s; return 0; }\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                               "-I", os.path.join(ROOT, "include"), src])
