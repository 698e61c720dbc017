"""Builds the C-ABI CUDA library (libb200fe.so) in-tree with nvcc for sm_100a.

The library is several translation units compiled in parallel: ``csrc/b200fe.cu`` (the C ABI, plan
construction, auxiliary kernels, host thread pool) and ``csrc/fbank_inst.cu`` once per instantiation
group of the fused kernel (``csrc/fbank_instances.h``).  Objects go to ``lighting-asr_b200/obj/``
(git-ignored), the linked library next to this file.

There is no CPU fallback: if the library is missing and cannot be built, importing the bindings raises."""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("B200FE_LIB") or os.path.join(HERE, "libb200fe.so")   # B200FE_LIB: A/B experiments with alternative builds
OBJ = (LIB + ".obj") if os.environ.get("B200FE_LIB") else os.path.join(HERE, "obj")
KERNEL_HDRS = ["fbank_kernel.cuh", "b200fe_common.cuh", "mel_static_default.inc", "fbank_instances.h"]
MAIN_HDRS = KERNEL_HDRS + ["aux_kernels.cuh", "stream_kernels.cuh", "resample_kernels.cuh", "host_pool.h", "fbank_ws_kernel.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O3,-pthread"]


def _groups():
    txt = open(os.path.join(CSRC, "fbank_instances.h")).read()
    return int(re.search(r"#define\s+B200FE_INST_GROUPS\s+(\d+)", txt).group(1))


def _units():
    """(object path, source, extra flags, dependency list)"""
    inc = os.path.join(os.path.dirname(HERE), "include", "b200fe.h")
    units = [(os.path.join(OBJ, "b200fe.o"), os.path.join(CSRC, "b200fe.cu"), [],
              [os.path.join(CSRC, "b200fe.cu"), inc] + [os.path.join(CSRC, h) for h in MAIN_HDRS])]
    # host-only SIMD loops of the staging pool: a .cpp, i.e. nvcc hands it to the host compiler untouched (function-level target attributes)
    units.append((os.path.join(OBJ, "host_simd.o"), os.path.join(CSRC, "host_simd.cpp"), [], [os.path.join(CSRC, "host_simd.cpp")]))
    for g in range(_groups()):
        units.append((os.path.join(OBJ, "fbank_inst_%d.o" % g), os.path.join(CSRC, "fbank_inst.cu"), ["-DB200FE_INST_GROUP=%d" % g],
                      [os.path.join(CSRC, "fbank_inst.cu")] + [os.path.join(CSRC, h) for h in KERNEL_HDRS]))
    return units


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _extra():
    return os.environ.get("B200FE_NVCC_EXTRA", "").split()        # e.g. -DB200FE_WARPS=6 for A/B experiments


def needs_build():
    if os.environ.get("B200FE_LIB") and __name__ != "__main__":
        return not os.path.exists(LIB)              # a variant library is used as built (the GPU box must not rebuild it)
    return any(_stale(LIB, deps) for _, _, _, deps in _units())


def build(force=False, verbose=False, jobs=None):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "flags.txt")
    flags = " ".join(NVCC_FLAGS + _extra())
    if not os.path.exists(stamp) or open(stamp).read() != flags:
        force = True
    todo = [(o, s, x) for o, s, x, deps in _units() if force or _stale(o, deps)]

    def compile_one(u):
        o, s, x = u
        cmd = [nvcc] + NVCC_FLAGS + _extra() + x + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", o, s]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=jobs or min(len(todo) or 1, os.cpu_count() or 4)) as ex:
        for cmd, r in ex.map(compile_one, todo):
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
            if verbose:
                sys.stderr.write(r.stderr)
    open(stamp, "w").write(flags)
    objs = [o for o, _, _, _ in _units()]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC,-pthread", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
