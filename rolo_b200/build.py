"""Builds librolo_gicp_b200.so (hand-written CUDA for sm_100a + the C-ABI) in-tree with nvcc.

    python -m rolo_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB = os.path.join(OUT_DIR, "librolo_gicp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v", "-I", os.path.join(HERE, "..", "include")]
EXTRA = os.environ.get("ROLO_NVCC_FLAGS", "").split()
SOURCES = ["linearize.cu", "voxelmap.cu", "covariance.cu", "frontend.cu", "scan2map.cu", "scancontext.cu", "context.cu", "abi_frontend.cu", "abi_mapping.cu", "testing_hooks.cu"]
# scan2map.cu restates float arithmetic of the reference operation by operation: no FMA contraction there
PER_FILE = {"scan2map.cu": ["-fmad=false"]}


def _deps():
    return [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "..", "include", "rolo_gicp.h"), __file__]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
    cmd = [NVCC, *ARCH, *COMMON, *PER_FILE.get(src, []), *EXTRA, "-c", os.path.join(SRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    if verbose:
        print(log)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(SRC, s))]
    if not force and not _stale(LIB, _deps()):
        if not os.path.exists(ALIAS):
            _alias()
        return LIB
    with cf.ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    cmd = [NVCC, *ARCH, "-shared", "-ccbin", HOST_CXX, "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    _alias()
    return LIB


ALIAS = os.path.join(OUT_DIR, "librot_gicp.so")   # the reference's library name (CMakeLists.txt:160): `target_link_libraries(... rot_gicp)` resolves to it


def _alias():
    try:
        if os.path.islink(ALIAS) or os.path.exists(ALIAS):
            os.remove(ALIAS)
        os.symlink(os.path.basename(LIB), ALIAS)
    except OSError:
        pass


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
