"""In-tree build of the sm_100a C-ABI library (`lib/libfloodplanet_b200.so`).

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the resulting .so is
git-ignored but travels with the repo snapshot to the GPU box.  Rebuilds only when a source
is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libfloodplanet_b200.so"
INCLUDE = PKG_DIR.parent / "include"

SOURCES = ["conv_igemm.cu", "conv_wgrad.cu", "elementwise.cu", "head_ce.cu", "fusion.cu",
           "augment.cu", "comm.cu"]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the floodplanet_b200 CUDA library cannot be built")


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    lib_m = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(
        INCLUDE.glob("*.h"))
    return any(d.stat().st_mtime > lib_m for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    LIB_DIR.mkdir(exist_ok=True)
    objdir = LIB_DIR / "obj"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = objdir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(CSRC / src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = (objdir / (Path(src).stem + ".log"))
        log.write_text(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-cudart", "static", "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


TEST_SRC = PKG_DIR.parent / "tests" / "csrc" / "conv_pertap_crosscheck.cu"
TEST_LIB_PATH = PKG_DIR.parent / "tests" / "lib" / "libfpb200_crosscheck.so"


def build_test_library(force: bool = False) -> Path:
    """TEST-ONLY: the first-generation per-tap conv kernel as its own shared library
    (tests/lib/libfpb200_crosscheck.so).  The product library neither contains nor exports it."""
    if not TEST_SRC.exists():
        raise RuntimeError(f"{TEST_SRC} missing")
    deps = [TEST_SRC, CSRC / "ptx.cuh", CSRC / "host_common.h", INCLUDE / "floodplanet_b200.h"]
    if (not force and TEST_LIB_PATH.exists()
            and all(d.stat().st_mtime <= TEST_LIB_PATH.stat().st_mtime for d in deps)):
        return TEST_LIB_PATH
    TEST_LIB_PATH.parent.mkdir(exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(INCLUDE), "-I", str(CSRC), "-shared", str(TEST_SRC), "-o",
           str(TEST_LIB_PATH), "-cudart", "static"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {TEST_SRC.name}:\n{res.stdout}\n{res.stderr}")
    return TEST_LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_test_library(force="--force" in sys.argv))
