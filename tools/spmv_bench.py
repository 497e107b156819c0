"""device-time probe of the 6x6 BCSR SpMV and the matrix axpy (HBM-bound next-row kernels)"""
import importlib, sys
import numpy as np
import torch
sys.path.insert(0, ".")
a2ds = importlib.import_module("a2d-shells_b200")
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
conn, X, bcn = a2ds.meshes.plate(nx, nx, bump=0.0)
n = len(X)
asm = a2ds.Assembler(0)
asm.set_mesh(conn, n); asm.set_nodes(X)
Cs, eth = a2ds.iso_shell_tables()
asm.set_components(Cs[None], eth[None]); asm.set_state(a2ds.meshes.seeded_state(np.arange(n), 1e-5))
k = asm.create_mat(); g = asm.create_mat()
asm.assembleAll(k, g, False)
nnz = asm.mat_nnz(k)
x = torch.randn(n, 6, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
for name, fn, nbytes in (("spmv6", lambda: asm.mat_mult_dev(k, x.data_ptr(), y.data_ptr()), nnz * (288 + 4) + 2 * n * 48),
                         ("axpy", lambda: asm.mat_axpy(0.5, g, k), nnz * 288 * 3),
                         ("copy", lambda: asm.mat_copy(k, g), nnz * 288 * 2)):
    for _ in range(3):
        fn()
    asm.region_begin()
    for _ in range(10):
        fn()
    ms = asm.region_end() / 10
    print(f"{name:6s} {ms:7.3f} ms  {nbytes / ms * 1e-6:8.1f} GB/s  ({nnz} blocks)")
# torch's own copy / axpy of the same size on the same box, for scale
a = torch.empty(nnz * 36, dtype=torch.float64, device="cuda").normal_(); b = torch.empty_like(a)
for name, fn, nbytes in (("torch copy_", lambda: b.copy_(a), nnz * 288 * 2),
                         ("torch add_", lambda: b.add_(a, alpha=0.5), nnz * 288 * 3)):
    for _ in range(3):
        fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:12s} {ms:7.3f} ms  {nbytes / ms * 1e-6:8.1f} GB/s")
