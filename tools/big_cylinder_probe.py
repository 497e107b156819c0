"""The 16 M-element cylinder of BASELINE configs[4] (4000 x 4000 MITC4, 16.0 M nodes,
144 M blocks = 41.5 GB per matrix) on ONE B200: fused residual + Kmat + Gmat, device resident,
with size-independent checks through the device SpMV (K u = r for the linear element without
BCs, symmetry of K and G).  Exercises every 64-bit offset (block values beyond 2^32 doubles).
  python tools/big_cylinder_probe.py [ntheta] [nx]"""
import importlib, json, sys, time
import numpy as np
sys.path.insert(0, ".")
a2ds = importlib.import_module("a2d-shells_b200")
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
t0 = time.time()
conn, X, ends = a2ds.meshes.cylinder(nt, nx)
n = len(X); ne = len(conn)
u = a2ds.meshes.seeded_state(np.arange(n), 1e-5)
Cs, eth = a2ds.iso_shell_tables()
t1 = time.time()
asm = a2ds.Assembler(0)
asm.set_mesh(conn, n); asm.set_nodes(X); asm.set_components(Cs[None], eth[None]); asm.set_state(u)
k = asm.create_mat(); g = asm.create_mat()
t2 = time.time()
nnz = asm.mat_nnz(k)
ms = []
for _ in range(3):
    asm.assembleAll(k, g, download=False); asm.synchronize()
    ms.append((asm.last_timing()[0], asm.last_kernel_ms()))
r = asm.assembleAll(k, g)
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
ku = asm.mat_mult(k, u)
e_kur = rel(ku, r)
rng = np.random.default_rng(0)
x = rng.normal(size=(n, 6)); y = rng.normal(size=(n, 6))
sym = []
for m in (k, g):
    a = float(np.sum(x * asm.mat_mult(m, y))); b = float(np.sum(y * asm.mat_mult(m, x)))
    sym.append(abs(a - b) / max(abs(a), abs(b)))
asm.close()
step_ms, kern_ms = min(ms)
out = dict(elements=ne, nodes=n, blocks_per_matrix=nnz, matrix_GB=nnz * 288 / 1e9,
           host_mesh_s=t1 - t0, pattern_and_offsets_s=t2 - t1, step_ms=step_ms, kernel_ms=kern_ms,
           elements_per_s=ne / (step_ms * 1e-3), Ku_equals_r=e_kur, K_symmetry=sym[0], G_symmetry=sym[1])
print("BIG_CYLINDER " + json.dumps(out))
assert e_kur < 1e-11 and max(sym) < 1e-9
print("BIG_CYLINDER_OK")
