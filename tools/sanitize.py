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
print("SANITIZE_DRIVER_DONE")
