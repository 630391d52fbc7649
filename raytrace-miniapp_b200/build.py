"""In-tree build of librtb200.so (hand-written sm_100a CUDA + C++ host layer + C ABI).

nvcc cross-compiles for sm_100a without a GPU.  The built library lives next to this file
(git-ignored, but it travels to the GPU box with the snapshot).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librtb200.so")
SOURCES = ["rtb200_kernels.cu", "rtb200_host.cu", "rtb200_multi.cu", "rtb200_dat.cpp"]
HEADERS = ["rtb200_march_flat.cuh", "rtb200_fp64.cuh", "rtb200_exptab.h", "rtb200_math.cuh", "rtb200_march.cuh", "rtb200_device.cuh", "rtb200_kernels.cuh",
           "rtb200_pack.h", os.path.join("..", "..", "include", "rtb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared", "-cudart", "static", "-ldl"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the rtb200 CUDA library cannot be built")


def source_hash():
    """sha256 (first 16 hex digits) over the library's sources and build flags: identifies a build
    of the kernels across machines (the .so itself is not bit-reproducible).  Profile figures
    are only quoted next to a run of the very sources they were captured from."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in sorted(SOURCES + HEADERS):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    """Compile csrc/ into librtb200.so for sm_100a.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    out = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (out.stdout, out.stderr))
    if verbose:
        print(out.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
