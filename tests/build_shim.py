"""Build the link-time drop-in (a2d-shells_b200/host/tacs_shim.cpp) against THIS
container's copy of the reference: headers from /root/reference, library = the single-rank
build under oracle/_ref.  Test infrastructure: in a deployment the maintainer compiles the
same source against their own liba2dshells.so.  The result goes to
tests/_shim/ and travels to the GPU box with the snapshot."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "a2d-shells_b200")
LIB = os.path.join(PKG, "lib", "liba2ds_b200.so")
SHIM = os.path.join(ROOT, "tests", "_shim", "libtacs_a2ds_shim.so")


def build_shim(reference="/root/reference"):
    src = os.path.join(PKG, "host", "tacs_shim.cpp")
    ref_lib_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(reference, "src", "TACSAssembler.h")):
        return SHIM if os.path.exists(SHIM) else None
    if not os.path.exists(os.path.join(ref_lib_dir, "liba2dshells_ref.so")):
        return None
    if os.path.exists(SHIM) and os.path.getmtime(SHIM) >= max(os.path.getmtime(src),
                                                              os.path.getmtime(LIB)):
        return SHIM
    os.makedirs(os.path.dirname(SHIM), exist_ok=True)
    inc = ["-I" + os.path.join(ROOT, "oracle", "stubs"), "-I" + os.path.join(ROOT, "include")]
    for d in ("", "bpmat", "elements", "elements/basis", "elements/shell", "constitutive", "io",
              "utils"):
        inc.append("-I" + os.path.join(reference, "src", d))
    cmd = (["g++", "-std=c++11", "-O2", "-fPIC", "-w", "-fno-access-control", "-shared"] + inc +
           ["-o", SHIM, src, "-L" + os.path.dirname(LIB), "-la2ds_b200",
            "-L" + ref_lib_dir, "-la2dshells_ref",
            "-Wl,-rpath,$ORIGIN/../../a2d-shells_b200/lib", "-Wl,-rpath,$ORIGIN/../../oracle/_ref"])
    subprocess.check_call(cmd)
    return SHIM


if __name__ == "__main__":
    print(build_shim())
