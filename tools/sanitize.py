"""Small driver for compute-sanitizer: every kernel family once on a tiny mesh (fused and
single-output assemblies in both scatter modes, both element classes, mass path, matrix-free
product, matrix algebra, SpMV).   compute-sanitizer --tool memcheck python tools/sanitize.py"""
import importlib, sys
import numpy as np
sys.path.insert(0, ".")
a2ds = importlib.import_module("a2d-shells_b200")
conn, X, comp, root = a2ds.meshes.wingbox(3, 2, 3, 2)
n = len(X); ncomp = int(comp.max()) + 1
Cs, eth = a2ds.iso_shell_tables(t_offset=0.2)
Csn = np.stack([Cs * (1 + 0.1 * k) for k in range(ncomp)]); ethn = np.stack([eth] * ncomp)
cls = (np.arange(ncomp) % 2).astype(np.int32)           # both element classes in one mesh
for mode in (a2ds.SCATTER_ATOMIC, a2ds.SCATTER_COLORED):
    asm = a2ds.Assembler(0)
    asm.set_mesh(conn, n, elem_comp=comp); asm.set_nodes(X)
    asm.set_components(Csn, ethn, temperature=np.full(ncomp, 3.0), elem_class=cls)
    asm.set_mass_moments(np.tile(a2ds.iso_mass_moments(2700.0, 0.01, 0.2), (ncomp, 1)))
    asm.set_bcs(root, 63); asm.set_scatter_mode(mode)
    u = a2ds.meshes.seeded_state(np.arange(n), 1e-4)
    asm.set_state(u); asm.set_state_rates(None, a2ds.meshes.seeded_state(np.arange(n) + 9, 1.0))
    k, g, m = asm.create_mat(), asm.create_mat(), asm.create_mat()
    r = asm.assembleAll(k, g)
    asm.assembleRes(); asm.assembleJacobian(1.0, 0.0, 2.0, k)
    for t in (0, 1, 2):
        asm.assembleMatType(t, m)
    asm.assembleMatCombo([0, 1, 2], [1.0, 0.5, -3.0], m)
    y = asm.addJacobianVecProduct(1.0, 1.0, u, np.zeros((n, 6)))
    asm.mat_copy(m, k); asm.mat_axpy(0.3, g, m); asm.mat_apply_bcs(m)
    z = asm.mat_mult(m, u)
    print("mode", mode, float(np.abs(r).max()), float(np.abs(y).max()), float(np.abs(z).max()))
    asm.close()
# uncoupled components as well (the tying-level kernel k_assemble_t; the wing box above has
# coupled sections and runs the first formulation), and a TACSParallelMat-style matrix: two
# BCSR blocks with row maps, the second one EMPTY (Bext on one rank), boundary conditions on
asm = a2ds.Assembler(0)
conn2, X2, ends2 = a2ds.meshes.cylinder(12, 5)
n2 = len(X2)
Cs2, eth2 = a2ds.iso_shell_tables()
asm.set_mesh(conn2, n2); asm.set_nodes(X2)
asm.set_components(np.stack([Cs2, Cs2]), np.stack([eth2, eth2]), elem_class=[0, 1])
asm.set_bcs(ends2, 0b100111)
asm.set_state(a2ds.meshes.seeded_state(np.arange(n2), 1e-4))
rowp, cols = asm.mat_pattern(asm.create_mat())
ident = np.arange(n2, dtype=np.int32); none = np.full(n2, -1, dtype=np.int32)
blocks = [dict(nrows=n2, rowp=rowp, cols=cols, row_map=ident, col_map=ident, ident=1),
          dict(nrows=0, rowp=np.zeros(1, np.int32), cols=np.zeros(0, np.int32), row_map=none,
               col_map=none, ident=0)]
pk = asm.create_mat_from_pattern(blocks); pg = asm.create_mat_from_pattern(blocks)
r2 = asm.assembleAll(pk, pg)
asm.assembleJacobian(1.0, 0.0, 0.0, pk); asm.assembleMatType(1, pg)
for _ in range(3):   # double-buffered matrices: swap with the spare array the previous kernel zeroed
    r2 = asm.assembleAll(pk, pg)
y2 = asm.addJacobianVecProduct(1.0, 1.0, np.ones((n2, 6)), np.zeros((n2, 6)))
print("two-block", float(np.abs(r2).max()), float(np.abs(asm.mat_values(pk)).max()), float(np.abs(y2).max()))
asm.close()
# streamed assembly (state chunks up, element ranges, residual chunks back on a third stream)
import os
os.environ["A2DS_STREAM_CHUNKS"] = "3"; os.environ["A2DS_STREAM_MIN_ELEMS"] = "1"
asm = a2ds.Assembler(0)
conn3, X3, bc3 = a2ds.meshes.plate(24, 17, bump=2e-2)
n3 = len(X3)
asm.set_mesh(conn3, n3); asm.set_nodes(X3)
asm.set_components(Cs2[None], eth2[None]); asm.set_bcs(bc3, 63)
k3, g3 = asm.create_mat(), asm.create_mat()
for _ in range(2):
    asm.set_state(a2ds.meshes.seeded_state(np.arange(n3), 1e-4))
    r3 = asm.assembleAll(k3, g3)
print("streamed", float(np.abs(r3).max()), asm.last_timing()[1])
asm.close()
# 9-node shells: k_assemble9 (producer warp + seven consumer warps over two shared-memory records)
asm = a2ds.Assembler(0)
conn9, X9, bc9 = a2ds.meshes.plate9(7, 5, bump=2e-2)
n9 = len(X9)
ec9 = (np.arange(len(conn9)) % 2).astype(np.int32)   # linear and nonlinear class in one mesh
asm.set_mesh(conn9, n9, elem_comp=ec9, order=3); asm.set_nodes(X9)
asm.set_components(np.stack([Cs2, Cs2]), np.stack([eth2, eth2]), temperature=[4.0, 4.0], elem_class=[0, 1])
asm.set_bcs(bc9, 63)
asm.set_state(a2ds.meshes.seeded_state(np.arange(n9), 1e-4))
k9, g9 = asm.create_mat(), asm.create_mat()
r9 = asm.assembleJacobian(1.0, 0.0, 0.0, k9); asm.assembleRes(); asm.assembleMatType(0, k9)
asm.assembleMatType(1, g9); asm.assembleAll(k9, g9)
print("quad9", float(np.abs(r9).max()), float(np.abs(asm.mat_values(k9)).max()), float(np.abs(asm.mat_values(g9)).max()))
asm.close()
# dependent nodes: two interior nodes of a plate replaced by the weighted mean of the 8 nodes
# around them (k_dep_gather / k_dep_scatter / k_dep_fold, scratch blocks behind the matrix)
os.environ["A2DS_STREAM_CHUNKS"] = "1"
asm = a2ds.Assembler(0)
connd, Xd, bcd = a2ds.meshes.plate(8, 6, bump=2e-2)
nd_ = len(Xd)
picks = [2 * 9 + 2, 3 * 9 + 5]
keep = np.ones(nd_, bool); keep[picks] = False
new = -np.ones(nd_, dtype=np.int64); new[keep] = np.arange(keep.sum())
dp, dc, dw = [0], [], []
for d, v in enumerate(picks):
    nb = sorted(set(int(w) for e in connd if v in e for w in e) - {v})
    dc += [int(new[w]) for w in nb]; dw += [1.0 / len(nb)] * len(nb); dp.append(len(dc))
    new[v] = -(d + 1)
asm.set_dependent_nodes(dp, dc, dw)
asm.set_mesh(new[connd].astype(np.int32), int(keep.sum()), elem_comp=(np.arange(len(connd)) % 2).astype(np.int32))
asm.set_nodes(Xd[keep])
asm.set_components(np.stack([Cs2, Cs]), np.stack([eth2, eth]), elem_class=[1, 0])   # both kernel families
asm.set_mass_moments(np.tile(a2ds.iso_mass_moments(2700.0, 0.01, 0.2), (2, 1)))
asm.set_bcs(new[bcd].astype(np.int32), 63)
asm.set_state(a2ds.meshes.seeded_state(np.arange(int(keep.sum())), 1e-4))
kd, gd = asm.create_mat(), asm.create_mat()
rd = asm.assembleAll(kd, gd); asm.assembleJacobian(1.0, 0.0, 2.0, kd); asm.assembleRes()
asm.assembleMatCombo([0, 1, 2], [1.0, 0.5, -3.0], gd)
print("dependent nodes", float(np.abs(rd).max()), float(np.abs(asm.mat_values(kd)).max()))
asm.close()
print("SANITIZE_DRIVER_DONE")
