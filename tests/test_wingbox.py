"""BASELINE configs[3] stand-in (the uCRM wing-box mesh is not in the reference checkout): a
synthetic wing box — skins, spars and ribs sharing their junction nodes, one component per
panel with a full 22-entry (composite-like, coupled) tangent, reference-axis transform along
the span — through pattern, colouring, partition (CPU) and the device assembly (GPU)."""
import importlib

import numpy as np
import pytest

from helpers import relmax

RES_TOL, MAT_TOL = 1e-12, 1e-10


def _components(a2ds, ncomp, seed=17):
    rng = np.random.default_rng(seed)
    Cs = np.zeros((ncomp, 22)); eth = np.zeros((ncomp, 9))
    for c in range(ncomp):
        base, e = a2ds.iso_shell_tables(E=70e9 * rng.uniform(0.5, 2.0), nu=rng.uniform(0.2, 0.4),
                                        t=rng.uniform(0.004, 0.02), t_offset=rng.uniform(-0.4, 0.4))
        base[2] = 0.08 * base[0] * rng.uniform(-1, 1)      # A16: off-axis plies
        base[7] *= 1.0 + 0.1 * rng.uniform()               # a little anisotropy in B
        base[14] = 0.08 * base[12] * rng.uniform(-1, 1)    # D16
        base[19] = 0.05 * base[18] * rng.uniform(-1, 1)    # As[1] != 0
        Cs[c] = base; eth[c] = e
    return Cs, eth


def test_wingbox_topology_pattern_colouring_partition(a2ds, orc):
    conn, X, comp, root = a2ds.meshes.wingbox(8, 2, 4, 2)
    n = len(X)
    assert comp.max() + 1 == 4 * 8 + 9 and len(np.unique(conn)) == n
    # every element is a proper quad (no repeated node, positive area both ways round)
    assert all(len(set(e)) == 4 for e in conn.tolist())
    a, b = X[conn[:, 1]] - X[conn[:, 0]], X[conn[:, 2]] - X[conn[:, 0]]
    c, d = X[conn[:, 2]] - X[conn[:, 3]], X[conn[:, 1]] - X[conn[:, 3]]
    assert np.linalg.norm(np.cross(a, b), axis=1).min() > 1e-4
    assert (np.einsum("ij,ij->i", np.cross(a, b), np.cross(c, d)) > 0).all()
    rp, cl = a2ds.host_pattern(n, conn)
    rpo, clo = orc.pattern(n, conn)
    assert rp.tobytes() == rpo.tobytes() and cl.tobytes() == clo.tobytes()
    # junction nodes: more than the 9 blocks per row of a structured surface
    assert np.diff(rp).max() == 12 and np.diff(rp).min() >= 6
    valence = np.bincount(conn.ravel(), minlength=n)
    assert valence.max() == 6
    color, nc = a2ds.host_color_elements(n, conn)
    assert nc >= 6
    for k in range(nc):                       # no two elements of a colour share a node
        nodes = conn[color == k].ravel()
        assert len(nodes) == len(np.unique(nodes))
    # span-wise partition through the native planner: every node owned once, pairs agree
    N = 3
    er = np.minimum((X[conn].mean(axis=1)[:, 0] / X[:, 0].max() * N).astype(int), N - 1)
    P = [a2ds.Partition(conn, n, er, N, r) for r in range(N)]
    owned = np.concatenate([p.glob[:p.n_owned] for p in P])
    assert len(owned) == n and len(np.unique(owned)) == n
    for r, p in enumerate(P):
        for k, q in enumerate(p.peers):
            o = P[int(q)]
            ko = list(o.peers).index(r)
            assert np.array_equal(p.glob[p.send_lists[k]], o.glob[o.recv_lists[ko]])


def test_rcb_partition_is_balanced_compact_and_deterministic(a2ds):
    """a2ds_partition_rcb (in place of the reference's METIS call): element counts differ by at
    most one per bisection, interfaces an order of magnitude shorter than a random partition,
    the result does not depend on anything but the mesh"""
    conn, X, comp, root = a2ds.meshes.wingbox(24, 4, 12, 3)
    n = len(X)
    for N in (1, 2, 3, 5, 8):
        er = a2ds.partition_rcb(conn, X, N)
        assert np.array_equal(er, a2ds.partition_rcb(conn, X, N))
        cnt = np.bincount(er, minlength=N)
        assert cnt.sum() == len(conn) and cnt.max() - cnt.min() <= int(np.ceil(np.log2(max(N, 2))))
        if N == 1:
            continue
        P = [a2ds.Partition(conn, n, er, N, r) for r in range(N)]
        rnd = np.random.default_rng(0).integers(0, N, len(conn))
        R = [a2ds.Partition(conn, n, rnd, N, r) for r in range(N)]
        ghosts = sum(p.n_nodes - p.n_owned for p in P)
        assert 10 * ghosts < sum(p.n_nodes - p.n_owned for p in R)
        assert sum(p.n_owned for p in P) == n
    # elements of a part are neighbours in space: every part's bounding box along the span is
    # a fraction of the wing for a span-dominated box
    er = a2ds.partition_rcb(conn, X, 8)
    cen = X[conn].mean(axis=1)
    widths = [np.ptp(cen[er == r, 0]) for r in range(8)]
    assert max(widths) < 0.3 * np.ptp(cen[:, 0])
    with pytest.raises(a2ds.A2dsError, match="outside"):
        a2ds.partition_rcb(conn + n, X, 2)


def test_wingbox_oracle_against_reference_live(a2ds, orc, ref):
    """pins the oracle on this topology and constitutive family: residual, K and G of the
    UNMODIFIED reference (general 22-entry tangent per component, reference-axis transform)"""
    conn, X, comp, root = a2ds.meshes.wingbox(4, 2, 3, 2)
    n = len(X); ncomp = int(comp.max()) + 1
    Cs, eth = _components(a2ds, ncomp)
    axis = np.array([1.0, 0.35, 0.0])
    props = np.stack([ref.general_props(0, Cs[c], eth[c, :3], (27.0, 0.0, 2e-4)) for c in range(ncomp)])
    ra = ref.RefAssembler(conn, X, comp, props, root, [list(range(6))] * len(root),
                          [[0.0] * 6] * len(root), transform=1, axis=axis)
    try:
        conn_r, X_r = ra.conn(), ra.nodes()
        nodes_b, vars_b, vals_b = ra.bcs()
        u = a2ds.meshes.seeded_state(np.arange(n), 1e-4)
        ra.set_state(u)
        m = ra.mat_create(0)
        r_ref = ra.assemble_jacobian(m)
        blk = ra.mat_block(m, 0)
        rowp, cols, K_ref = blk["rowp"], blk["cols"], blk["A"]
        ra.assemble_mat_type(1, m)
        G_ref = ra.mat_block(m, 0)["A"]
    finally:
        ra.close()
    comps = [orc.make_comp(0, Cs[c], eth[c], (27.0, 0.0, 2e-4), 0.0, 1, axis) for c in range(ncomp)]
    # elements keep their order through TACSCreator on one rank, so comp carries over
    r, K = orc.assemble(1, conn_r, comp, comps, X_r, u, rowp, cols, nodes_b, vars_b, vals_b)
    _, G = orc.assemble(3, conn_r, comp, comps, X_r, u, rowp, cols, nodes_b, vars_b, vals_b)
    assert relmax(r, r_ref) < 1e-13 and relmax(K, K_ref) < 1e-13 and relmax(G, G_ref) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_wingbox_device_assembly_vs_oracle(a2ds, orc, mode):
    conn, X, comp, root = a2ds.meshes.wingbox(24, 2, 4, 2)       # 121 components
    n = len(X); ncomp = int(comp.max()) + 1
    assert ncomp == 121
    Cs, eth = _components(a2ds, ncomp)
    axis = np.array([1.0, 0.35, 0.0])
    u = a2ds.meshes.seeded_state(np.arange(n), 1e-4)
    asm = a2ds.Assembler(0)
    asm.set_mesh(conn, n, elem_comp=comp); asm.set_nodes(X)
    asm.set_components(Cs, eth, transform=a2ds.TRANSFORM_REF_AXIS, ref_axis=axis)
    asm.set_bcs(root, 63); asm.set_state(u)
    asm.set_scatter_mode(a2ds.SCATTER_COLORED if mode else a2ds.SCATTER_ATOMIC)
    k, g = asm.create_mat(), asm.create_mat()
    rowp, cols = asm.mat_pattern(k)
    assert np.diff(rowp).max() == 12
    res = asm.assembleAll(k, g)
    K, G = asm.mat_values(k), asm.mat_values(g)
    # the separate entry points (buckling flow: K, then G) give the same matrices
    asm.assembleMatType(a2ds.STIFFNESS_MATRIX, k)
    asm.assembleMatType(a2ds.GEOMETRIC_STIFFNESS_MATRIX, g)
    K2, G2 = asm.mat_values(k), asm.mat_values(g)
    asm.close()
    comps = [orc.make_comp(0, Cs[c], eth[c], (0, 0, 0), 0.0, 1, axis) for c in range(ncomp)]
    bc_vars = np.full(len(root), 63, dtype=np.int32); bc_vals = np.zeros((len(root), 6))
    r_o, k_o = orc.assemble(1, conn, comp, comps, X, u, rowp, cols, root, bc_vars, bc_vals)
    _, g_o = orc.assemble(3, conn, comp, comps, X, u, rowp, cols, root, bc_vars, bc_vals)
    assert relmax(res, r_o) < RES_TOL and relmax(K, k_o) < MAT_TOL and relmax(G, g_o) < MAT_TOL
    assert relmax(K2, k_o) < MAT_TOL and relmax(G2, g_o) < MAT_TOL
