"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference
(oracle/_ref, built by oracle/Makefile from /root/reference).  Run in the build
container:  python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these are outputs
of the reference itself on seeded inputs:
  elements.npz : 12 random elements x {linear, nonlinear} x {natural, ref-axis}
                 x {T=0 & no offset, T=10 & offset}: addJacobian res/K and, for the
                 linear class, getMatType(G)
  plate.npz, cylinder.npz : small assembled meshes (reference numbering): pattern,
                 res, K, G of assembleJacobian / assembleMatType, with BCs
  quad9.npz    : the 9-node element (TACSQuad9Shell): random elements and a small assembled plate
  dep.npz      : a small curved panel with DEPENDENT nodes (TACSCreator::setDependentNodes), both
                 element classes: pattern, res, K, G, M and the dynamic Jacobian, reference numbering
  buckling.npz : lowest 6 buckling eigenvalues of a 40x20 cylinder
  bdf.npz      : what the reference's TACSMeshLoader reads from the decks in this
                 directory (mixed.bdf: hand-written, every card family and field format;
                 cyl_large.bdf / cyl_small.bdf / cyl_free.bdf: written here by
                 meshes.write_bdf with shuffled file numbers) and from the two decks the
                 reference ships (examples/cylinder-buckling, kept as array digests only)
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import refdrv  # noqa: E402
from helpers import random_elements, random_elements9, with_dependent_nodes  # noqa: E402

a2ds = importlib.import_module("a2d-shells_b200")
AXIS = np.array([0.3, 1.0, 0.2])


def elements():
    X, q = random_elements(12, seed=2024)
    out = dict(X=X, q=q, axis=AXIS)
    for kind in (0, 1):
        for tr in (0, 1):
            for ci, (T, off) in enumerate(((0.0, 0.0), (10.0, 0.3))):
                p = refdrv.iso_props(kind=kind, temperature=T, t_offset=off)
                Cs, eth, _ = refdrv.con_tables(p)
                r, k, _ = refdrv.element_batch(p, 1, X.reshape(-1, 12), q.reshape(-1, 24),
                                               transform=tr, axis=AXIS)
                key = f"k{kind}_t{tr}_c{ci}"
                out[key + "_res"] = r; out[key + "_K"] = k
                out[key + "_Cs"] = Cs; out[key + "_eth"] = eth; out[key + "_T"] = T
                if kind == 0:
                    _, g, _ = refdrv.element_batch(p, 3, X.reshape(-1, 12), q.reshape(-1, 24),
                                                   transform=tr, axis=AXIS)
                    out[key + "_G"] = g
    np.savez_compressed(os.path.join(HERE, "elements.npz"), **out)


def quad9():
    """quad9.npz: the 9-node element (TACSQuad9Shell / TACSQuad9NonlinearShell, props kind 2 / 3):
    4 random elements x {linear, nonlinear} x {natural, ref-axis} x {T=0, T=10 & offset}
    addJacobian res/K (+ getMatType(G) for the linear class), and a small assembled plate of
    9-node elements with boundary conditions (reference numbering)"""
    X, q = random_elements9(4, seed=909)
    out = dict(X=X, q=q, axis=AXIS)
    for kind in (0, 1):
        for tr in (0, 1):
            for ci, (T, off) in enumerate(((0.0, 0.0), (10.0, 0.3))):
                p = refdrv.iso_props(kind=2 + kind, temperature=T, t_offset=off)
                Cs, eth, _ = refdrv.con_tables(p)
                key = f"k{kind}_t{tr}_c{ci}"
                rs, ks, gs = [], [], []
                for e in range(X.shape[0]):
                    r, k = refdrv.element(p, 1, X[e].ravel(), q[e].ravel(), transform=tr, axis=AXIS)
                    rs.append(r); ks.append(k)
                    if kind == 0:
                        gs.append(refdrv.element(p, 3, X[e].ravel(), q[e].ravel(), transform=tr,
                                                 axis=AXIS)[1])
                out[key + "_res"] = np.array(rs); out[key + "_K"] = np.array(ks)
                out[key + "_Cs"] = Cs; out[key + "_eth"] = eth; out[key + "_T"] = T
                if kind == 0:
                    out[key + "_G"] = np.array(gs)
    conn, Xm, bcn = a2ds.meshes.plate9(3, 2, bump=2e-2)
    n = len(Xm)
    p = refdrv.iso_props(kind=2)
    Cs, eth, _ = refdrv.con_tables(p)
    bc_vars = [list(range(6)) if i % 2 else [0, 1, 2] for i in range(len(bcn))]
    bc_vals = [[-1e-5] + [0.0] * (len(v) - 1) for v in bc_vars]
    ra = refdrv.RefAssembler(conn, Xm, np.zeros(len(conn), dtype=np.int32), p[None], bcn, bc_vars,
                             bc_vals, nodes_per_elem=9)
    conn_r, X_r = ra.conn(), ra.nodes()
    nodes_b, vars_b, vals_b = ra.bcs()
    u = a2ds.meshes.seeded_state(np.arange(n), 1e-5)
    ra.set_state(u)
    m = ra.mat_create(0)
    res = ra.assemble_jacobian(m)
    blk = ra.mat_block(m, 0)
    out.update(m_conn=conn_r, m_X=X_r, m_u=u, m_bc_nodes=nodes_b, m_bc_vars=vars_b, m_bc_vals=vals_b,
               m_rowp=blk["rowp"], m_cols=blk["cols"], m_res=res, m_K=blk["A"], m_Cs=Cs, m_eth=eth)
    ra.close()
    np.savez_compressed(os.path.join(HERE, "quad9.npz"), **out)


def mesh(name):
    if name == "plate":
        conn, X, bcn = a2ds.meshes.plate(7, 5, bump=2e-2)
    else:
        conn, X, bcn = a2ds.meshes.cylinder(12, 4)
    n = len(X)
    bc_vars = [list(range(6)) if i % 2 else [0, 1, 2] for i in range(len(bcn))]
    bc_vals = [[-1e-5] + [0.0] * (len(v) - 1) for v in bc_vars]
    props = refdrv.iso_props()
    ra = refdrv.RefAssembler(conn, X, np.zeros(len(conn), dtype=np.int32), props[None], bcn,
                             bc_vars, bc_vals)
    u = np.zeros((n, 6))
    u[ra.new_nodes] = a2ds.meshes.seeded_state(np.arange(n), scale=1e-5)
    ra.set_state(u)
    pm = ra.mat_create(0)
    res = ra.assemble_jacobian(pm)
    blk = ra.mat_block(pm, 0)
    ra.assemble_mat_type(1, pm)
    G = ra.mat_block(pm, 0)["A"]
    res_only = ra.assemble_res()
    nodes_b, vars_b, vals_b = ra.bcs()
    Cs, eth, _ = refdrv.con_tables(props)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), conn=ra.conn(), X=ra.nodes(), u=u,
                        bc_nodes=nodes_b, bc_vars=vars_b, bc_vals=vals_b, rowp=blk["rowp"],
                        cols=blk["cols"], K=blk["A"], G=G, res=res, res_only=res_only, Cs=Cs,
                        eth=eth, new_nodes=ra.new_nodes)
    ra.close()


def dep():
    """dependent nodes through the reference's own path (TACSAssembler.h:469-510, TACSBVec.cpp:855-975)"""
    conn, X, bcn = a2ds.meshes.plate(6, 5, bump=0.05)
    conn2, Xi, bc2, dp = with_dependent_nodes(conn, X, bcn, 3, seed=1)
    n = len(Xi)
    out = {}
    for kind in (0, 1):
        p = refdrv.iso_props(kind=kind, t_offset=0.25)
        Cs, eth, mom = refdrv.con_tables(p)
        bc_vars = [list(range(6)) if i % 2 else [0, 1, 2] for i in range(len(bc2))]
        bc_vals = [[-1e-5] + [0.0] * (len(v) - 1) for v in bc_vars]
        ra = refdrv.RefAssembler(conn2, Xi, np.zeros(len(conn2), dtype=np.int32), p[None], bc2, bc_vars,
                                 bc_vals, dep=dp)
        u = a2ds.meshes.seeded_state(np.arange(n), 1e-4)
        udd = a2ds.meshes.seeded_state(np.arange(n) + 1000, 1.0)
        ra.set_state(u, None, udd)
        m = ra.mat_create(0)
        ra.assemble_mat_type(2, m)
        blk = ra.mat_block(m, 0)
        res_dyn = ra.assemble_jacobian(m, alpha=1.0, gamma=3.0)
        J = ra.mat_block(m, 0)["A"]
        ra.assemble_mat_type(1, m)
        G = ra.mat_block(m, 0)["A"]
        ra.set_state(u)
        res = ra.assemble_jacobian(m)
        K = ra.mat_block(m, 0)["A"]
        nodes_b, vars_b, vals_b = ra.bcs()
        dr = ra.dep()
        if kind == 0:
            out.update(conn=ra.conn(), X=ra.nodes(), u=u, udd=udd, bc_nodes=nodes_b, bc_vars=vars_b,
                       bc_vals=vals_b, rowp=blk["rowp"], cols=blk["cols"], dep_ptr=dr[0], dep_conn=dr[1],
                       dep_w=dr[2], Cs=Cs, eth=eth, mom=mom)
        out.update({"M%d" % kind: blk["A"], "J%d" % kind: J, "res_dyn%d" % kind: res_dyn,
                    "G%d" % kind: G, "K%d" % kind: K, "res%d" % kind: res})
        ra.close()
    np.savez_compressed(os.path.join(HERE, "dep.npz"), **out)


def bdf():
    out = {}
    rng = np.random.default_rng(7)
    conn, X, bcn = a2ds.meshes.cylinder(16, 5)
    nid = rng.permutation(len(X)) * 3 + 5
    eid = rng.permutation(len(conn)) * 2 + 11
    comp = rng.integers(0, 3, len(conn))
    for fmt in ("large", "small", "free"):
        a2ds.meshes.write_bdf(os.path.join(HERE, f"cyl_{fmt}.bdf"), conn, X, bcn,
                              ["123456" if i % 2 else "13" for i in range(len(bcn))],
                              [(-1e-5 if i % 3 else 0.0) for i in range(len(bcn))], comp, fmt, nid,
                              eid, comp_names=["SKIN", "RIB.001", "SPAR"])
    decks = {f"cyl_{f}": os.path.join(HERE, f"cyl_{f}.bdf") for f in ("large", "small", "free")}
    decks["mixed"] = os.path.join(HERE, "mixed.bdf")
    for name, path in decks.items():
        fail, ref, _ = refdrv.bdf_scan(path)
        assert fail == 0, path
        for k, v in ref.items():
            out[f"{name}_{k}"] = np.array(v)
    # the shipped decks cannot travel with the repository: keep what the reference reads from
    # them as sizes and checksums
    ex = "/root/reference/examples/cylinder-buckling"
    for name in ("mech-cylinder", "therm-cylinder"):
        fail, ref, _ = refdrv.bdf_scan(os.path.join(ex, name + ".bdf"))
        assert fail == 0
        out[name + "_sizes"] = np.array([len(ref["X"]), len(ref["elem_comp"]), len(ref["bc_nodes"])])
        out[name + "_digest"] = np.array([
            float(ref["X"].sum()), float(np.abs(ref["X"]).sum()),
            float((ref["elem_conn"].astype(np.int64) * (1 + np.arange(len(ref["elem_conn"])) % 7)).sum()),
            float((ref["bc_nodes"].astype(np.int64) * (1 + ref["bc_vars"])).sum()),
            float(ref["bc_vals"].sum())])
    np.savez_compressed(os.path.join(HERE, "bdf.npz"), **out)


def buckling():
    """cylinder 40x20 under end shortening; the reference solves for the load path itself
    (u0 = NULL) and runs Lanczos with the shipped example's settings (100 vectors, 50
    eigenvalues, mechBuckling.cpp:136-138); shift 12 is close to the lowest cluster."""
    conn, X, ends = a2ds.meshes.cylinder(40, 20)
    bc_vars = [[0, 1, 2, 5]] * len(ends)
    bc_vals = [[-1e-3 if i >= 40 else 0.0, 0.0, 0.0, 0.0] for i in range(len(ends))]
    ra = refdrv.RefAssembler(conn, X, np.zeros(len(conn), dtype=np.int32),
                             refdrv.iso_props()[None], ends, bc_vars, bc_vals)
    km, gm, am = ra.mat_create(1), ra.mat_create(1), ra.mat_create(1)
    eig, err = ra.buckling(km, gm, am, 0, sigma=12.0, num_eigs=50, max_lanczos=100, u0=None)
    np.savez_compressed(os.path.join(HERE, "buckling.npz"), eig=eig[:8], err=err[:8],
                        path=ra.path)
    print("buckling eigenvalues", eig[:8], err[:8])
    ra.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["elements", "plate", "cylinder", "buckling", "bdf", "quad9", "dep"]
    for w in which:   # e.g. `python make_golden.py bdf` regenerates only bdf.npz
        {"elements": elements, "plate": lambda: mesh("plate"), "cylinder": lambda: mesh("cylinder"),
         "buckling": buckling, "bdf": bdf, "quad9": quad9, "dep": dep}[w]()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
