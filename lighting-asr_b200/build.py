"""Builds the C-ABI CUDA library (libb200fe.so) in-tree with nvcc for sm_100a.

There is no CPU fallback: if the library is missing and cannot be built, importing the
bindings raises."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "b200fe.cu")
LIB = os.environ.get("B200FE_LIB") or os.path.join(HERE, "libb200fe.so")   # B200FE_LIB: A/B experiments with alternative builds
DEPS = [os.path.join(HERE, "csrc", f) for f in ("b200fe.cu", "fbank_kernel.cuh", "fbank_ws_kernel.cuh", "aux_kernels.cuh", "b200fe_common.cuh", "mel_static_default.inc")] + \
       [os.path.join(os.path.dirname(HERE), "include", "b200fe.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("B200FE_NVCC_EXTRA", "").split()        # e.g. -DB200FE_WARPS=6 for A/B experiments
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
