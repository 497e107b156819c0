"""small driver for ncu: a few fused residual+K+G assemblies of an nx x nx plate"""
import importlib, sys
import numpy as np
sys.path.insert(0, ".")
a2ds = importlib.import_module("a2d-shells_b200")
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 300
what = sys.argv[2] if len(sys.argv) > 2 else "all"
conn, X, bcn = a2ds.meshes.plate(nx, nx, bump=0.0)
n = len(X)
asm = a2ds.Assembler(0)
asm.set_mesh(conn, n); asm.set_nodes(X)
Cs, eth = a2ds.iso_shell_tables()
asm.set_components(Cs[None], eth[None]); asm.set_state(a2ds.meshes.seeded_state(np.arange(n), 1e-5))
asm.set_bcs(bcn, 63)
k = asm.create_mat(); g = asm.create_mat()
for _ in range(4):
    if what == "all":
        asm.assembleAll(k, g, False)
    elif what == "K":
        asm.assembleMatType(0, k)
    elif what == "G":
        asm.assembleMatType(1, g)
    else:
        asm.assembleRes(False)
asm.synchronize()
print("kernel ms", asm.last_kernel_ms(), "elements", len(conn))
