"""CPU: libcm2.so builds for sm_100a, loads, and exports every symbol include/cm2.h declares
(no compute calls -- there is no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from centermask2_b200 import build
    path, _ = build.build()
    return path


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cm2.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cm2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported(libpath):
    from centermask2_b200 import lib
    names = declared_symbols()
    assert len(names) >= 25
    assert set(names) == set(lib.SYMBOLS), set(names) ^ set(lib.SYMBOLS)
    handle = lib.load()
    for n in names:
        assert hasattr(handle, n), n
    assert handle.cm2_version() >= 101


def test_library_is_sm100a_only_and_has_no_undefined_cuda_driver_deps(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    ldd = subprocess.run(["ldd", libpath], capture_output=True, text=True).stdout
    assert "libcuda.so" not in ldd and "not found" not in ldd


def test_tensor_core_engine_emits_tcgen05_and_tma_sass(libpath):
    sass = subprocess.run(["cuobjdump", "-sass", libpath], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass, "no tcgen05.mma in SASS"
    assert "UTMALDG" in sass, "no TMA tensor load in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"


def test_error_reporting_without_device(libpath):
    from centermask2_b200 import lib
    import ctypes as C
    h = lib.load()
    d = lib.ConvDesc()
    assert h.cm2_conv2d(C.byref(d), None) == -1
    assert b"num_src" in h.cm2_last_error()
    a = lib.Act()
    assert h.cm2_maxpool3x3s2_ceil(C.byref(a), C.byref(a), 0, None) == -1


def test_struct_layout_matches_header():
    """ctypes mirror of cm2_conv_desc must have the C layout (checked against sizeof from a tiny C program)."""
    from centermask2_b200 import lib
    import ctypes as C
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "cm2.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu",' \
          'sizeof(cm2_act),sizeof(cm2_conv_desc),offsetof(cm2_conv_desc,weight),offsetof(cm2_conv_desc,out),' \
          'offsetof(cm2_conv_desc,stats),offsetof(cm2_conv_desc,src_phase),offsetof(cm2_conv_desc,seg),' \
          'sizeof(cm2_seg),offsetof(cm2_conv_desc,stats_mode));return 0;}'
    exe = "/tmp/cm2_layout_check"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True, check=True)
    vals = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert vals == [C.sizeof(lib.Act), C.sizeof(lib.ConvDesc), lib.ConvDesc.weight.offset, lib.ConvDesc.out.offset,
                    lib.ConvDesc.stats.offset, lib.ConvDesc.src_phase.offset, lib.ConvDesc.seg.offset, C.sizeof(lib.Seg),
                    lib.ConvDesc.stats_mode.offset]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No fallback path: without the built libcm2.so every product entry raises (nothing silently runs on torch / CPU)."""
    import pytest as _pytest
    from centermask2_b200 import lib as cmlib
    monkeypatch.setattr(cmlib, "_lib", None)
    monkeypatch.setattr(cmlib, "LIB_PATH", str(tmp_path / "libcm2.so"))
    with _pytest.raises(RuntimeError, match="no fallback"):
        cmlib.load()
    with _pytest.raises(RuntimeError, match="no fallback"):
        cmlib.last_error()                          # every wrapper goes through load()
