"""The oracle (oracle/shell_oracle.c) against the golden fixtures generated from the
unmodified reference (tests/golden/make_golden.py) and, where oracle/_ref is present,
against the reference itself.  CPU only."""
import os

import numpy as np
import pytest

from helpers import random_elements, random_elements9, relmax

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cases():
    for kind in (0, 1):
        for tr in (0, 1):
            for ci in (0, 1):
                yield kind, tr, ci


@pytest.mark.parametrize("kind,tr,ci", list(_cases()))
def test_elements_against_golden(orc, kind, tr, ci):
    g = np.load(os.path.join(GOLD, "elements.npz"))
    key = f"k{kind}_t{tr}_c{ci}"
    T = float(g[key + "_T"])
    comp = orc.make_comp(kind, g[key + "_Cs"], g[key + "_eth"], (0, 0, 0), T, tr, g["axis"])
    for e in range(g["X"].shape[0]):
        X = g["X"][e].ravel(); q = g["q"][e].ravel()
        r, k = orc.jacobian(comp, X, q)
        assert relmax(r, g[key + "_res"][e]) < 1e-13
        assert relmax(k, g[key + "_K"][e]) < 1e-13
        assert relmax(orc.residual(comp, X, q), g[key + "_res"][e]) < 1e-13
        assert relmax(orc.mat_type(comp, 0, X, q), g[key + "_K"][e]) < 1e-13
        if kind == 0:
            # G is a finite difference in the reference: agreement is limited by its
            # cancellation noise (1e-12 at T = 0, ~1e-7 with a temperature)
            tol = 1e-10 if T == 0.0 else 1e-6
            assert relmax(orc.mat_type(comp, 1, X, q), g[key + "_G"][e]) < tol


@pytest.mark.parametrize("name", ["plate", "cylinder"])
def test_assembly_against_golden(orc, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    conn, X, u = g["conn"], g["X"], g["u"]
    n = len(X)
    rowp, cols = orc.pattern(n, conn)
    # bit-exact sparsity pattern (TACSAssembler::createMat, natural order)
    assert rowp.tobytes() == g["rowp"].tobytes() and cols.tobytes() == g["cols"].tobytes()
    comp = orc.make_comp(0, g["Cs"], g["eth"])
    ec = np.zeros(len(conn), dtype=np.int32)
    args = (conn, ec, [comp], X, u, rowp, cols, g["bc_nodes"], g["bc_vars"], g["bc_vals"])
    r, K = orc.assemble(1, *args)
    assert relmax(r, g["res"]) < 1e-13 and relmax(K, g["K"]) < 1e-13
    r0, _ = orc.assemble(0, *args)
    assert relmax(r0, g["res_only"]) < 1e-13
    _, G = orc.assemble(3, *args)
    assert relmax(G, g["G"]) < 1e-10
    # BC rows: zero except 1 on the diagonal; residual rows u - ubar
    for b, nd in enumerate(g["bc_nodes"]):
        for k in range(6):
            if g["bc_vars"][b] & (1 << k):
                assert r[nd, k] == u[nd, k] - g["bc_vals"][b, k]
                for j in range(rowp[nd], rowp[nd + 1]):
                    row = K[j, k].copy()
                    if cols[j] == nd:
                        assert row[k] == 1.0
                        row[k] = 0.0
                    assert not row.any()


def test_oracle_against_reference_live(orc, ref):
    """random elements beyond the fixtures, all model/transform combinations"""
    X, q = random_elements(40, seed=99)
    axis = np.array([0.3, 1.0, 0.2])
    for kind in (0, 1):
        for tr in (0, 1):
            p = ref.iso_props(kind=kind, temperature=0.0, t_offset=0.2)
            Cs, eth, mom = ref.con_tables(p)
            comp = orc.make_comp(kind, Cs, eth, mom, 0.0, tr, axis)
            r_ref, k_ref, _ = ref.element_batch(p, 1, X.reshape(-1, 12), q.reshape(-1, 24),
                                                transform=tr, axis=axis)
            _, g_ref, _ = ref.element_batch(p, 3, X.reshape(-1, 12), q.reshape(-1, 24),
                                            transform=tr, axis=axis)
            for e in range(X.shape[0]):
                r, k = orc.jacobian(comp, X[e].ravel(), q[e].ravel())
                assert relmax(r, r_ref[e]) < 1e-13 and relmax(k, k_ref[e]) < 1e-13
                assert relmax(orc.mat_type(comp, 1, X[e].ravel(), q[e].ravel()), g_ref[e]) < 1e-10


def test_empty_and_degenerate_inputs(orc):
    rowp, cols = orc.pattern(3, np.zeros((0, 4), dtype=np.int32))
    assert rowp.tolist() == [0, 0, 0, 0] and len(cols) == 0
    # zero state: zero residual; G of a zero state is exactly zero in the oracle's
    # central difference (both evaluations coincide)
    X, _ = random_elements(1, seed=3)
    Cs = np.ones(22); Cs[6:12] = 0.0
    comp = orc.make_comp(0, Cs, np.zeros(9))
    r, k = orc.jacobian(comp, X[0].ravel(), np.zeros(24))
    assert not r.any() and np.abs(k - k.T).max() <= 1e-12 * np.abs(k).max()


def test_oracle_mass_and_inertia_against_reference_live(orc, ref):
    """TACS_MASS_MATRIX and the gamma / qddot terms of addJacobian (TACSShellElement.h:410-447,
    614-648) — element level, both element classes and transforms, offset mid-surface so
    that the first mass moment is not zero"""
    X, q = random_elements(12, seed=7)
    rng = np.random.default_rng(5)
    qdd = rng.uniform(-1.0, 1.0, size=(X.shape[0], 24))
    axis = np.array([0.3, 1.0, 0.2])
    for kind in (0, 1):
        for tr in (0, 1):
            p = ref.iso_props(kind=kind, temperature=0.0, t_offset=0.3)
            Cs, eth, mom = ref.con_tables(p)
            assert mom[0] > 0 and mom[1] != 0 and mom[2] > 0
            comp = orc.make_comp(kind, Cs, eth, mom, 0.0, tr, axis)
            for e in range(X.shape[0]):
                Xe, qe = X[e].ravel(), q[e].ravel()
                _, m_ref = ref.element(p, 4, Xe, qe, transform=tr, axis=axis)
                assert relmax(orc.mat_type(comp, 2, Xe, qe), m_ref) < 1e-13
                r_ref, j_ref = ref.element(p, 1, Xe, qe, ddvars=qdd[e], alpha=0.7, beta=0.3,
                                           gamma=2.5, transform=tr, axis=axis)
                r, j = orc.jacobian_dyn(comp, Xe, qe, qdd[e], alpha=0.7, gamma=2.5)
                assert relmax(r, r_ref) < 1e-13 and relmax(j, j_ref) < 1e-13


def test_oracle_dynamic_assembly_against_reference_live(orc, ref):
    """assembleMatType(MASS) and assembleJacobian(alpha, beta, gamma) with second time
    derivatives set, on a small curved panel with boundary conditions"""
    import importlib
    a2ds = importlib.import_module("a2d-shells_b200")
    conn, X, bcn = a2ds.meshes.plate(6, 5, bump=0.05)
    n = len(X)
    p = ref.iso_props(kind=0, t_offset=0.25)
    Cs, eth, mom = ref.con_tables(p)
    bc_vars = [list(range(6))] * len(bcn)
    bc_vals = [[0.0] * 6] * len(bcn)
    ra = ref.RefAssembler(conn, X, np.zeros(len(conn), dtype=np.int32), p[None], bcn, bc_vars,
                          bc_vals)
    try:
        # everything below in the reference's own node numbering
        conn_r, X_r = ra.conn(), ra.nodes()
        nodes_b, vars_b, vals_b = ra.bcs()
        u = a2ds.meshes.seeded_state(np.arange(n), 1e-4)
        udd = a2ds.meshes.seeded_state(np.arange(n) + 1000, 1.0)
        ra.set_state(u, None, udd)
        m = ra.mat_create(0)
        ra.assemble_mat_type(2, m)
        blk = ra.mat_block(m, 0)
        rowp, cols, M_ref = blk["rowp"], blk["cols"], blk["A"]
        r_ref = ra.assemble_jacobian(m, alpha=1.0, beta=0.0, gamma=3.0)
        J_ref = ra.mat_block(m, 0)["A"]
        r0_ref = ra.assemble_res()
    finally:
        ra.close()
    comp = orc.make_comp(0, Cs, eth, mom)
    ec = np.zeros(len(conn_r), dtype=np.int32)
    args = (conn_r, ec, [comp], X_r, u, rowp, cols, nodes_b, vars_b, vals_b)
    _, M = orc.assemble(4, *args)
    assert relmax(M, M_ref) < 1e-13
    r, J = orc.assemble(1, *args, alpha=1.0, gamma=3.0, udd=udd)
    assert relmax(r, r_ref) < 1e-13 and relmax(J, J_ref) < 1e-13
    r0, _ = orc.assemble(0, *args, udd=udd)
    assert relmax(r0, r0_ref) < 1e-13


# ---- 9-node element (TACSQuad9Shell, oracle9_*: shell_oracle.c compiled with ORACLE_ORDER 3) ----
@pytest.mark.parametrize("kind,tr,ci", list(_cases()))
def test_quad9_elements_against_golden(orc, kind, tr, ci):
    g = np.load(os.path.join(GOLD, "quad9.npz"))
    key = f"k{kind}_t{tr}_c{ci}"
    T = float(g[key + "_T"])
    comp = orc.make_comp(kind, g[key + "_Cs"], g[key + "_eth"], (0, 0, 0), T, tr, g["axis"])
    for e in range(g["X"].shape[0]):
        X = g["X"][e].ravel(); q = g["q"][e].ravel()
        r, k = orc.jacobian(comp, X, q, order=3)
        assert relmax(r, g[key + "_res"][e]) < 1e-13
        assert relmax(k, g[key + "_K"][e]) < 1e-13
        assert relmax(orc.residual(comp, X, q, order=3), g[key + "_res"][e]) < 1e-13
        if kind == 0:
            tol = 1e-10 if T == 0.0 else 1e-6   # the reference's finite-difference noise, as above
            assert relmax(orc.mat_type(comp, 1, X, q, order=3), g[key + "_G"][e]) < tol


def test_quad9_assembly_against_golden(orc):
    g = np.load(os.path.join(GOLD, "quad9.npz"))
    conn, X, u = g["m_conn"], g["m_X"], g["m_u"]
    n = len(X)
    rowp, cols = orc.pattern(n, conn, order=3)
    assert rowp.tobytes() == g["m_rowp"].tobytes() and cols.tobytes() == g["m_cols"].tobytes()
    comp = orc.make_comp(0, g["m_Cs"], g["m_eth"])
    ec = np.zeros(len(conn), dtype=np.int32)
    r, K = orc.assemble(1, conn, ec, [comp], X, u, rowp, cols, g["m_bc_nodes"], g["m_bc_vars"],
                        g["m_bc_vals"], order=3)
    assert relmax(r, g["m_res"]) < 1e-13 and relmax(K, g["m_K"]) < 1e-13


def test_quad9_oracle_against_reference_live(orc, ref):
    """random 9-node elements beyond the fixture: residual, tangent, geometric stiffness and
    mass matrix of TACSQuad9Shell / TACSQuad9NonlinearShell"""
    X, q = random_elements9(10, seed=31)
    axis = np.array([0.3, 1.0, 0.2])
    for kind in (0, 1):
        for tr in (0, 1):
            p = ref.iso_props(kind=2 + kind, temperature=0.0, t_offset=0.2)
            Cs, eth, mom = ref.con_tables(p)
            assert mom[0] > 0
            comp = orc.make_comp(kind, Cs, eth, mom, 0.0, tr, axis)
            for e in range(X.shape[0]):
                Xe, qe = X[e].ravel(), q[e].ravel()
                r_ref, k_ref = ref.element(p, 1, Xe, qe, transform=tr, axis=axis)
                r, k = orc.jacobian(comp, Xe, qe, order=3)
                assert relmax(r, r_ref) < 1e-13 and relmax(k, k_ref) < 1e-13
                g_ref = ref.element(p, 3, Xe, qe, transform=tr, axis=axis)[1]
                assert relmax(orc.mat_type(comp, 1, Xe, qe, order=3), g_ref) < 1e-10
                m_ref = ref.element(p, 4, Xe, qe, transform=tr, axis=axis)[1]
                assert np.abs(m_ref).max() > 0
                assert relmax(orc.mat_type(comp, 2, Xe, qe, order=3), m_ref) < 1e-13


def test_dependent_nodes_against_golden(orc):
    """mesh with dependent nodes (TACSAssembler::setDependentNodes): pattern bit-exact, res / K /
    G / M / dynamic Jacobian of both element classes against the reference's outputs (dep.npz)"""
    g = np.load(os.path.join(GOLD, "dep.npz"))
    conn, X, u, udd = g["conn"], g["X"], g["u"], g["udd"]
    assert (conn < 0).any()
    dep = (g["dep_ptr"], g["dep_conn"], g["dep_w"])
    rowp, cols = orc.pattern_dep(len(X), conn, dep[0], dep[1])
    assert rowp.tobytes() == g["rowp"].tobytes() and cols.tobytes() == g["cols"].tobytes()
    ec = np.zeros(len(conn), dtype=np.int32)
    for kind in (0, 1):
        comp = orc.make_comp(kind, g["Cs"], g["eth"], g["mom"])
        args = (conn, ec, [comp], X, u, dep, rowp, cols, g["bc_nodes"], g["bc_vars"], g["bc_vals"])
        r, K = orc.assemble_dep(1, *args)
        assert relmax(r, g["res%d" % kind]) < 1e-13 and relmax(K, g["K%d" % kind]) < 1e-13
        _, G = orc.assemble_dep(3, *args)
        assert relmax(G, g["G%d" % kind]) < 1e-10
        _, M = orc.assemble_dep(4, *args)
        assert relmax(M, g["M%d" % kind]) < 1e-13
        r, J = orc.assemble_dep(1, *args, alpha=1.0, gamma=3.0, udd=udd)
        assert relmax(r, g["res_dyn%d" % kind]) < 1e-13 and relmax(J, g["J%d" % kind]) < 1e-13


def test_dependent_nodes_against_reference_live(orc, ref):
    """the same on another mesh / seed against the unmodified reference run here; without
    dependent nodes the dep entry points reduce to the plain ones"""
    import importlib
    from helpers import with_dependent_nodes
    a2ds = importlib.import_module("a2d-shells_b200")
    conn, X, bcn = a2ds.meshes.cylinder(10, 6)
    conn2, Xi, bc2, dp = with_dependent_nodes(conn, X, bcn, 5, seed=4)
    n = len(Xi)
    p = ref.iso_props(kind=1, t_offset=0.1)
    Cs, eth, mom = ref.con_tables(p)
    ra = ref.RefAssembler(conn2, Xi, np.zeros(len(conn2), dtype=np.int32), p[None], bc2,
                          [list(range(6))] * len(bc2), [[0.0] * 6] * len(bc2), dep=dp)
    try:
        conn_r, X_r, dep_r = ra.conn(), ra.nodes(), ra.dep()
        nodes_b, vars_b, vals_b = ra.bcs()
        u = a2ds.meshes.seeded_state(np.arange(n), 1e-4)
        ra.set_state(u)
        m = ra.mat_create(0)
        r_ref = ra.assemble_jacobian(m)
        blk = ra.mat_block(m, 0)
    finally:
        ra.close()
    rowp, cols = orc.pattern_dep(n, conn_r, dep_r[0], dep_r[1])
    assert np.array_equal(rowp, blk["rowp"]) and np.array_equal(cols, blk["cols"])
    comp = orc.make_comp(1, Cs, eth, mom)
    ec = np.zeros(len(conn_r), dtype=np.int32)
    r, K = orc.assemble_dep(1, conn_r, ec, [comp], X_r, u, dep_r, rowp, cols, nodes_b, vars_b, vals_b)
    assert relmax(r, r_ref) < 1e-13 and relmax(K, blk["A"]) < 1e-13
    # no dependent nodes: identical to the plain entry points
    none = (np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
    rp0, cl0 = orc.pattern(len(X), conn)
    rp1, cl1 = orc.pattern_dep(len(X), conn, none[0], none[1])
    assert np.array_equal(rp0, rp1) and np.array_equal(cl0, cl1)
    uu = a2ds.meshes.seeded_state(np.arange(len(X)), 1e-4)
    ec = np.zeros(len(conn), dtype=np.int32)
    ra_, Ka = orc.assemble(1, conn, ec, [comp], X, uu, rp0, cl0)
    rb_, Kb = orc.assemble_dep(1, conn, ec, [comp], X, uu, none, rp0, cl0)
    assert np.array_equal(ra_, rb_) and np.array_equal(Ka, Kb)
