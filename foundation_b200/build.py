"""In-tree build of libfoundation_pt.so (sm_100a).  nvcc cross-compiles without a GPU; the .so travels to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.environ.get("FOUNDATION_PT_LIB") or os.path.join(LIB_DIR, "libfoundation_pt.so")   # override: A/B experiments only
SOURCES = ["foundation_pt.cu"]
HEADERS = ["pt_math.h", "pt_layout.h", "pt_shading.h", "pt_host_shared.h", "pt_build.h", "pt_traverse.h", "pt_kernels.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17", "-Xcompiler",
              "-fPIC,-ffp-contract=off,-mfma,-fvisibility=hidden", "-shared"]


def nvcc_path() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_HERE, "..", "include", "foundation_pt.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the C-ABI library for sm_100a if it is missing or older than its sources.
    Safe under concurrent callers (one rank per GPU all importing at once): an exclusive file lock serialises the build and the
    result is moved into place atomically, so no process can ever dlopen a half-written library."""
    if not force and (os.environ.get("FOUNDATION_PT_LIB") or not is_stale()):   # an A/B variant is used as it is, never rebuilt from the tree
        return LIB_PATH
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():          # another process built it while we waited
                return LIB_PATH
            tmp = LIB_PATH + f".tmp{os.getpid()}"
            cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
