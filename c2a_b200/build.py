"""Builds the in-tree native library ``c2a_b200/csrc/libc2a_b200.so`` (CUDA kernels + C ABI +
host-side C++ drop-in API) for sm_100a with nvcc.  nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libc2a_b200.so")
CU_SOURCES = ["c2a_kernels.cu", "c2a_host_model.cpp", "c2a_dropin.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-ffp-contract=off",  # host half (motion constants) must not fuse either
              "-fmad=false",  # bit-parity with the reference's -ffp-contract=off CPU build
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    deps.append(os.path.join(CSRC, "..", "..", "include", "c2a_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in CU_SOURCES]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout, out.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
