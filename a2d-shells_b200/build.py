"""In-tree build of the CUDA library (sm_100a only).  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "liba2ds_b200.so")
SRCS = [os.path.join(HERE, "csrc", "a2ds.cu"), os.path.join(HERE, "csrc", "mesh_io.cpp"),
        os.path.join(HERE, "csrc", "partition.cpp")]
DEPS = SRCS + [os.path.join(HERE, "csrc", "mitc4_math.h"), os.path.join(HERE, "csrc", "mitc4_tying.h"),
               os.path.join(HERE, "csrc", "assemble_kernels.cuh"),
               os.path.join(HERE, "csrc", "mitc9_math.h"), os.path.join(HERE, "csrc", "assemble9_kernels.cuh"),
               os.path.join(HERE, "csrc", "aux_kernels.cuh"),
               os.path.join(HERE, "..", "include", "a2ds.h")]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False):
    if up_to_date() and not force:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SRCS + ["-lnccl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
