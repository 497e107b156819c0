"""Run the reference's buckling flow (TACSLinearBuckling::solve) on a synthetic cylinder and
print the lowest eigenvalues.  With LD_PRELOAD=libtacs_a2ds_shim.so the three TACSAssembler
assembly entry points inside that flow run on the GPU; without it they are the reference's."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import refdrv  # noqa: E402

a2ds_meshes = importlib.import_module("a2d-shells_b200.meshes")
if len(sys.argv) > 1 and sys.argv[1] == "quad9":
    # the same flow with 9-node shells (TACSQuad9Shell): buckling (K and G through
    # assembleMatType) and one Jacobian assembly; no mass terms for this element on the device
    conn, X, ends = a2ds_meshes.cylinder9(20, 10)
    nt = 40
    bc_vars = [[0, 1, 2, 5]] * len(ends)
    bc_vals = [[-1e-3 if i >= nt else 0.0, 0.0, 0.0, 0.0] for i in range(len(ends))]
    ra = refdrv.RefAssembler(conn, X, np.zeros(len(conn), dtype=np.int32), refdrv.iso_props(kind=2)[None],
                             ends, bc_vars, bc_vals, nodes_per_elem=9)
    km, gm, am = ra.mat_create(1), ra.mat_create(1), ra.mat_create(1)
    u = np.zeros((len(X), 6)); u[ra.new_nodes] = a2ds_meshes.seeded_state(np.arange(len(X)), 1e-3)
    ra.set_state(u)
    # K and G into the TACSSchurMat blocks at a fixed state (weighted checksums over all four blocks)
    chk = {}
    for typ, mat, tag in ((0, km, "k"), (1, gm, "g")):
        ra.assemble_mat_type(typ, mat)
        tot, mx = 0.0, 0.0
        for which in range(4):
            blk = ra.mat_block(mat, which)
            if blk is None or blk["A"].size == 0:
                continue
            Ab = blk["A"]
            tot += float((Ab * np.cos(np.arange(Ab.size) + which).reshape(Ab.shape)).sum())
            mx = max(mx, float(np.abs(Ab).max()))
        chk[tag + "_chk"] = tot; chk[tag + "_max"] = mx
    # shift below the lowest eigenvalue: with a shift inside the cluster at 11.7 .. 12 the
    # shifted operator is nearly singular and the Lanczos result moves with the last bit of G
    eig, err = ra.buckling(km, gm, am, 0, sigma=10.0, num_eigs=50, max_lanczos=100, u0=None)
    path = ra.path.copy()
    chk["path_chk"] = float((path * np.cos(np.arange(path.size)).reshape(path.shape)).sum())
    chk["path_max"] = float(np.abs(path).max())
    gblk = ra.mat_block(gm, 0)["A"]
    chk["gpath_chk"] = float((gblk * np.cos(np.arange(gblk.size)).reshape(gblk.shape)).sum())
    chk["gpath_max"] = float(np.abs(gblk).max())
    pm = ra.mat_create(0)
    u = np.zeros((len(X), 6)); u[ra.new_nodes] = a2ds_meshes.seeded_state(np.arange(len(X)), 1e-5)
    ra.set_state(u)
    r = ra.assemble_jacobian(pm)
    A = ra.mat_block(pm, 0)["A"]
    w = np.cos(np.arange(A.size)).reshape(A.shape)
    print("SHIM_PROBE " + json.dumps(dict(eig=eig[:6].tolist(), err=err[:6].tolist(), **chk,
                                         res_norm=float(np.abs(r).max()), res_sum=float(r.sum()),
                                         a_max=float(np.abs(A).max()), a_sum=float(A.sum()),
                                         a_chk=float((A * w).sum()))))
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "dep":
    # a cylinder with DEPENDENT nodes (TACSCreator::setDependentNodes): K and G into the four
    # TACSSchurMat blocks, the buckling flow, and a Jacobian into a TACSParallelMat
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import with_dependent_nodes
    conn, X, ends = a2ds_meshes.cylinder(40, 20)
    conn2, Xi, ends2, dep = with_dependent_nodes(conn, X, ends, 30, seed=3)
    bc_vars = [[0, 1, 2, 5]] * len(ends2)
    bc_vals = [[-1e-3 if i >= 40 else 0.0, 0.0, 0.0, 0.0] for i in range(len(ends2))]
    ra = refdrv.RefAssembler(conn2, Xi, np.zeros(len(conn2), dtype=np.int32), refdrv.iso_props()[None],
                             ends2, bc_vars, bc_vals, dep=dep)
    n = len(Xi)
    km, gm, am = ra.mat_create(1), ra.mat_create(1), ra.mat_create(1)
    u = np.zeros((n, 6)); u[ra.new_nodes] = a2ds_meshes.seeded_state(np.arange(n), 1e-3)
    ra.set_state(u)
    chk = {}
    for typ, mat, tag in ((0, km, "k"), (1, gm, "g")):
        ra.assemble_mat_type(typ, mat)
        tot, mx = 0.0, 0.0
        for which in range(4):
            blk = ra.mat_block(mat, which)
            if blk is None or blk["A"].size == 0:
                continue
            Ab = blk["A"]
            tot += float((Ab * np.cos(np.arange(Ab.size) + which).reshape(Ab.shape)).sum())
            mx = max(mx, float(np.abs(Ab).max()))
        chk[tag + "_chk"] = tot; chk[tag + "_max"] = mx
    eig, err = ra.buckling(km, gm, am, 0, sigma=12.0, num_eigs=50, max_lanczos=100, u0=None)
    pm = ra.mat_create(0)
    u = np.zeros((n, 6)); u[ra.new_nodes] = a2ds_meshes.seeded_state(np.arange(n), 1e-5)
    ra.set_state(u)
    r = ra.assemble_jacobian(pm)
    A = ra.mat_block(pm, 0)["A"]
    w = np.cos(np.arange(A.size)).reshape(A.shape)
    print("SHIM_PROBE " + json.dumps(dict(eig=eig[:6].tolist(), err=err[:6].tolist(), **chk,
                                         res_norm=float(np.abs(r).max()),
                                         res_chk=float((r * w.ravel()[:r.size].reshape(r.shape)).sum()),
                                         a_max=float(np.abs(A).max()), a_chk=float((A * w).sum()))))
    sys.exit(0)
conn, X, ends = a2ds_meshes.cylinder(40, 20)
bc_vars = [[0, 1, 2, 5]] * len(ends)
bc_vals = [[-1e-3 if i >= 40 else 0.0, 0.0, 0.0, 0.0] for i in range(len(ends))]
ra = refdrv.RefAssembler(conn, X, np.zeros(len(conn), dtype=np.int32), refdrv.iso_props()[None],
                         ends, bc_vars, bc_vals)
km, gm, am = ra.mat_create(1), ra.mat_create(1), ra.mat_create(1)
eig, err = ra.buckling(km, gm, am, 0, sigma=12.0, num_eigs=50, max_lanczos=100, u0=None)
# also one plain Jacobian assembly into a TACSParallelMat
pm = ra.mat_create(0)
u = np.zeros((len(X), 6)); u[ra.new_nodes] = a2ds_meshes.seeded_state(np.arange(len(X)), 1e-5)
ra.set_state(u)
r = ra.assemble_jacobian(pm)
A = ra.mat_block(pm, 0)["A"]
# the dynamic terms: mass matrix, and a Jacobian with gamma and second time derivatives
ra.assemble_mat_type(2, pm)
M = ra.mat_block(pm, 0)["A"]
udd = np.zeros((len(X), 6)); udd[ra.new_nodes] = a2ds_meshes.seeded_state(np.arange(len(X)) + 5, 1.0)
ra.set_state(u, None, udd)
rd = ra.assemble_jacobian(pm, alpha=1.0, beta=0.0, gamma=1e4)
Ad = ra.mat_block(pm, 0)["A"]
# TACS_MAT_TRANSPOSE: the same Jacobian assembled with the element matrices transposed
rt = ra.assemble_jacobian(pm, alpha=1.0, beta=0.0, gamma=1e4, transpose=True)
At = ra.mat_block(pm, 0)["A"]
# TACSFrequencyAnalysis::solve (K and M through assembleMatType, shift-invert Lanczos)
fk, fm = ra.mat_create(1), ra.mat_create(1)
feig, ferr = ra.frequency(fk, fm, sigma=1e6, num_eigs=6, max_lanczos=60)
w = np.cos(np.arange(A.size)).reshape(A.shape)   # fixed weights: a checksum that sees every entry
print("SHIM_PROBE " + json.dumps(dict(eig=eig[:6].tolist(), err=err[:6].tolist(),
                                     feig=feig.tolist(), ferr=ferr.tolist(),
                                     res_norm=float(np.abs(r).max()), res_sum=float(r.sum()),
                                     a_max=float(np.abs(A).max()), a_sum=float(A.sum()),
                                     m_max=float(np.abs(M[M != 1.0]).max()), m_chk=float(((M - (M == 1.0)) * w).sum()),
                                     rd_max=float(np.abs(rd).max()), rd_chk=float((rd * w.ravel()[:rd.size].reshape(rd.shape)).sum()),
                                     ad_max=float(np.abs(Ad).max()), ad_chk=float((Ad * w).sum()),
                                     at_max=float(np.abs(At).max()), at_chk=float((At * w).sum()),
                                     at_vs_ad=float(np.abs(At - Ad).max() / np.abs(Ad).max()))))
