"""Can a large memset run WHILE the fused element kernel occupies every SM?  (development aid)
Times: the fused assembly alone, a 5.2 GB zero-fill alone (torch, second stream), and both
started together.  If together ~ max(a, b) the zeroing of a later element range could hide
behind the kernel of an earlier one; if ~ a + b it cannot (the memset kernel finds no room
next to 2 x 128 threads x 244 registers + 222 KB of shared memory per SM)."""
import importlib, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
a2ds = importlib.import_module("a2d-shells_b200")
nx = 1000
conn, X, bcn = a2ds.meshes.plate(nx, nx, bump=0.0)
n = len(X)
asm = a2ds.Assembler(0)
asm.set_mesh(conn, n); asm.set_nodes(X)
Cs, eth = a2ds.iso_shell_tables()
asm.set_components(Cs[None], eth[None]); asm.set_state(a2ds.meshes.seeded_state(np.arange(n), 1e-5))
asm.set_bcs(bcn, 63)
k, g = asm.create_mat(), asm.create_mat()
big = torch.empty(int(5.2e9) // 8, dtype=torch.float64, device="cuda")
side = torch.cuda.Stream()
def run(assemble, zero):
    torch.cuda.synchronize(); asm.synchronize()
    t0 = time.perf_counter()
    if assemble:
        asm.assembleAll(k, g, download=False)
    if zero:
        with torch.cuda.stream(side):
            big.zero_()
    asm.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3
for _ in range(3):
    run(True, True)
for name, a, z in (("assembly alone", True, False), ("zero-fill alone", False, True), ("both together", True, True)):
    ts = [run(a, z) for _ in range(5)]
    print(f"{name:18s} {np.median(ts):8.3f} ms")
# the same with a device-to-device copy from a zero buffer (copy engine?)
zsrc = torch.zeros(int(2.6e9) // 8, dtype=torch.float64, device="cuda")
def run2(assemble, copy):
    torch.cuda.synchronize(); asm.synchronize()
    t0 = time.perf_counter()
    if assemble:
        asm.assembleAll(k, g, download=False)
    if copy:
        with torch.cuda.stream(side):
            big[:zsrc.numel()].copy_(zsrc)
    asm.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3
for name, a, z in (("d2d copy alone", False, True), ("assembly + d2d copy", True, True)):
    ts = [run2(a, z) for _ in range(5)]
    print(f"{name:18s} {np.median(ts):8.3f} ms")
