"""quick device-time probe of the assembly kernels (development aid, not bench.py)"""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
a2ds = importlib.import_module("a2d-shells_b200")
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 500
conn, X, bcn = a2ds.meshes.plate(nx, nx, bump=1e-3)
n = len(X)
u = a2ds.meshes.seeded_state(np.arange(n), 1e-5)
Cs, eth = a2ds.iso_shell_tables()
asm = a2ds.Assembler(0)
asm.set_mesh(conn, n); asm.set_nodes(X); asm.set_components(Cs[None], eth[None]); asm.set_state(u)
asm.set_bcs(bcn, 63)
if len(sys.argv) > 2 and sys.argv[2] == "colored":
    asm.set_scatter_mode(a2ds.SCATTER_COLORED)   # one launch per element colour, bit-reproducible
asm.set_mass_moments(a2ds.iso_mass_moments()[None])
import torch
xd = torch.randn(n, 6, dtype=torch.float64, device="cuda"); yd = torch.zeros_like(xd)
t0 = time.time(); k = asm.create_mat(); g = asm.create_mat(); print("mat create s", time.time() - t0)
ne = len(conn)
for name, fn in (("res", lambda: asm.assembleRes(False)),
                 ("jac(res+K)", lambda: asm.assembleJacobian(1.0, 0, 0, k, False)),
                 ("K", lambda: asm.assembleMatType(0, k)),
                 ("G", lambda: asm.assembleMatType(1, g)),
                 ("all(res+K+G)", lambda: asm.assembleAll(k, g, False)),
                 ("M", lambda: asm.assembleMatType(2, g)),
                 ("jac+gammaM", lambda: asm.assembleJacobian(1.0, 0, 2.0, k, False)),
                 ("K x (matrix free)", lambda: asm.addJacobianVecProduct_dev(1.0, 1.0, xd.data_ptr(), yd.data_ptr()))):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(5):
        fn(); ts.append(asm.last_timing()[0])
    ms = np.median(ts)
    print(f"{name:14s} {ms:8.3f} ms  {ne / ms * 1e-3:8.2f} M elem/s  launches {asm.last_timing()[1]}")
