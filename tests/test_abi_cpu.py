"""CPU tests (no GPU, no compute calls): the C-ABI library loads, exports every symbol include/rolo_gicp.h declares,
and the product refuses to run without a device instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from rolo_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    L = capi.load()
    names = capi.declared_symbols()
    assert len(names) >= 55
    missing = [s for s in names if not hasattr(L, s)]
    assert not missing, missing
    assert L.rg_abi_version() == 1


def test_exported_symbols_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", capi.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (rg_[a-z0-9_]+)\b", out))
    assert set(capi.declared_symbols()) <= exported


def test_sm100a_cubin_with_tma_and_no_other_arch():
    out = subprocess.run(["cuobjdump", "-lelf", capi.lib_path()], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", capi.lib_path()], capture_output=True, text=True, check=True).stdout
    assert "UBLKCP" in sass            # 1-D TMA bulk copies in the fused kernel
    assert "SYNCS" in sass             # mbarrier transactions
    assert "DFMA" in sass              # fp64 accumulation
    assert "HMMA" not in sass and "UTCHMMA" not in sass   # no tensor-core path by design


def test_no_device_is_a_loud_error_not_a_fallback():
    L = capi.load()
    if L.rg_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert L.rg_create(0, C.byref(h)) == capi.RG_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.rg_last_error(None)
    from rolo_b200.registration import RotVGICP
    with pytest.raises(capi.RoloError):
        RotVGICP()


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/: nothing under rolo_b200/ imports, includes, links or loads it."""
    pkg = os.path.join(ROOT, "rolo_b200")
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#\s*include\s*[<\"].*oracle)|librolo_oracle|rolo_oracle\.h|oracle/_build", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(txt), os.path.join(dirpath, f)
    out = subprocess.run(["ldd", capi.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_c_examples_compile_and_link_as_plain_c(tmp_path):
    """Both examples are C99 against include/rolo_gicp.h only (no GPU needed to compile and link)."""
    import subprocess
    libdir = os.path.dirname(capi.lib_path())
    capi.load()
    for name in ("lidar_odometry_sequence", "back_end_sequence"):
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", name + ".c"),
                               "-o", str(tmp_path / name), "-L", libdir, "-lrolo_gicp_b200", f"-Wl,-rpath,{libdir}", "-lm"])
