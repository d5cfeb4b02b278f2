"""Builds libfulgor_gpu.so (CUDA kernels + C ABI) and the `fulgor_b200_pseudoalign` CLI in-tree with nvcc for
sm_100a. No JIT cache: the artefacts live next to the sources so they travel with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfulgor_gpu.so")
CLI = os.path.join(HERE, "fulgor_b200_pseudoalign")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wextra,-Wno-unused-parameter", "--cudart", "static"]

LIB_SOURCES = [os.path.join(CSRC, f) for f in ("engine.cu", "fur_reader.cpp")]
LIB_DEPS = LIB_SOURCES + [os.path.join(CSRC, f) for f in ("kernels.cuh", "pipeline_kernels.cuh", "image.h", "fur_reader.h")] + [
    os.path.join(os.path.dirname(HERE), "include", "fulgor_gpu.h")]
CLI_SOURCES = [os.path.join(CSRC, "pseudoalign_cli.cpp")]
CLI_DEPS = CLI_SOURCES + [os.path.join(CSRC, "fastx_io.h"), os.path.join(os.path.dirname(HERE), "include", "fulgor_gpu.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if force or _stale(LIB, LIB_DEPS):
        cmd = [NVCC] + ARCH + COMMON + os.environ.get("FG_NVCC_FLAGS", "").split() + ["-shared", "-o", LIB] + LIB_SOURCES
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    if all(os.path.exists(s) for s in CLI_SOURCES) and (force or _stale(CLI, CLI_DEPS + [LIB])):
        cmd = [NVCC] + ARCH + COMMON + ["-o", CLI] + CLI_SOURCES + ["-L" + HERE, "-lfulgor_gpu", "-Xlinker", "-rpath=$ORIGIN", "-lz", "-lpthread"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
