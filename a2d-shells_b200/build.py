"""In-tree build of the CUDA library (sm_100a only).  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "liba2ds_b200.so")
SRCS = [os.path.join(HERE, "csrc", "a2ds.cu")]
DEPS = SRCS + [os.path.join(HERE, "csrc", "mitc4_math.h"),
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


SHIM = os.path.join(HERE, "lib", "libtacs_a2ds_shim.so")


def build_shim(reference="/root/reference", ref_lib_dir=None):
    """The link-time replacement of TACSAssembler::assembleRes/Jacobian/MatType
    (host/tacs_shim.cpp).  Needs the reference's headers, so it is only built where the
    reference tree is present (the build container); the .so travels with the snapshot.
    It is linked against the reference library it interposes on (here: the single-rank
    build under oracle/_ref) — in a deployment that is the user's liba2dshells.so."""
    src = os.path.join(HERE, "host", "tacs_shim.cpp")
    if not os.path.exists(os.path.join(reference, "src", "TACSAssembler.h")):
        return None
    ref_lib_dir = ref_lib_dir or os.path.join(HERE, "..", "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_lib_dir, "liba2dshells_ref.so")):
        return None
    if os.path.exists(SHIM) and os.path.getmtime(SHIM) >= max(os.path.getmtime(src),
                                                              os.path.getmtime(LIB)):
        return SHIM
    inc = ["-I" + os.path.join(HERE, "..", "oracle", "stubs"), "-I" + os.path.join(HERE, "..", "include")]
    for d in ("", "bpmat", "elements", "elements/basis", "elements/shell", "constitutive", "io", "utils"):
        inc.append("-I" + os.path.join(reference, "src", d))
    cmd = (["g++", "-std=c++11", "-O2", "-fPIC", "-w", "-fno-access-control", "-shared"] + inc +
           ["-o", SHIM, src, "-L" + os.path.dirname(LIB), "-la2ds_b200",
            "-L" + ref_lib_dir, "-la2dshells_ref",
            "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../../oracle/_ref"])
    subprocess.check_call(cmd)
    return SHIM


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_shim())
