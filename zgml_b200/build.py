"""Build recipes: the sm_100a CUDA library (product) and, separately, the CPU oracle
(test infrastructure under oracle/).  nvcc cross-compiles without a GPU."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "zgml_b200", "csrc")
LIB_DIR = os.path.join(ROOT, "zgml_b200", "lib")
LIB = os.path.join(LIB_DIR, "libzgml_cuda.so")
ORACLE_SRC = os.path.join(ROOT, "oracle", "zgml_oracle.c")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libzgml_oracle.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-cudart", "static"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_cuda(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "zgml_cuda.h")]
    if not force and _newer(LIB, deps):
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    subprocess.run(cmd, check=True)
    return LIB


def build_oracle(force=False):
    deps = [ORACLE_SRC, os.path.join(ROOT, "include", "zgml_cuda.h")]
    if not force and _newer(ORACLE_LIB, deps):
        return ORACLE_LIB
    os.makedirs(os.path.dirname(ORACLE_LIB), exist_ok=True)
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-fvisibility=hidden", "-Wall", "-o", ORACLE_LIB, ORACLE_SRC, "-lm", "-lpthread"]
    subprocess.run(cmd, check=True)
    return ORACLE_LIB


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracle(force="--force" in sys.argv)
    print(LIB, ORACLE_LIB)
