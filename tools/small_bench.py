"""latency probe on a mesh of the shipped example's size (3 200 elements): device time between
the first and last operation of a call, and host wall time per call"""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
a2ds = importlib.import_module("a2d-shells_b200")
conn, X, bcn = a2ds.meshes.cylinder(80, 40)
n = len(X)
Cs, eth = a2ds.iso_shell_tables()
asm = a2ds.Assembler(0)
asm.set_mesh(conn, n); asm.set_nodes(X); asm.set_components(Cs[None], eth[None])
asm.set_state(a2ds.meshes.seeded_state(np.arange(n), 1e-5)); asm.set_bcs(bcn, 63)
k = asm.create_mat(); g = asm.create_mat()
for name, fn in (("res", lambda: asm.assembleRes(False)),
                 ("jac(res+K)", lambda: asm.assembleJacobian(1.0, 0, 0, k, False)),
                 ("G", lambda: asm.assembleMatType(1, g)),
                 ("all(res+K+G)", lambda: asm.assembleAll(k, g, False))):
    for _ in range(5):
        fn()
    asm.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
    asm.synchronize()
    wall = (time.perf_counter() - t0) / 200 * 1e6
    dev = asm.last_timing()[0] * 1e3
    print(f"{name:14s} device {dev:7.1f} us   wall/call {wall:7.1f} us   {len(conn) / wall:6.1f} M elem/s  launches {asm.last_timing()[1]}")
