"""Sampled-row parity of a device assembly against the plain-C oracle — TEST INFRASTRUCTURE.

Used by tests/ and by bench.py's `parity` block (outside every timed region).  For a set of
sampled owned nodes the oracle assembles just the elements around them (a sub-mesh with an
order-preserving renumbering, so block columns come out in the same order as on the device)
and the node rows of the residual, of K and of G are compared with the device's.  This is a
VALUE check at sizes where the whole mesh is out of the oracle's reach (1 M ... 16 M
elements); the full-mesh comparisons live in tests/test_gpu_parity.py at small sizes.

Multi-rank (matrices without a matrix halo, the TACSSchurMat convention): interface rows of
K and G hold the LOCAL element contributions only and are compared as such; the residual of
an interface node is complete after the reverse halo exchange, so the contributions of the
neighbour's elements (evaluated by the neighbour's oracle on its ghost copy of the node) are
collected with all_gather_object and added before the comparison.
"""
import numpy as np

import oracle_py as orc


def _node_to_elems(n_nodes, conn):
    flat = conn.ravel()
    order = np.argsort(flat, kind="stable")
    ptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(ptr, flat + 1, 1)
    np.cumsum(ptr, out=ptr)
    return ptr, (order // 4).astype(np.int64)


def _submesh(nodes, ptr, adj, conn):
    """elements touching `nodes`, their nodes (sorted: order preserving) and local connectivity"""
    el = np.unique(np.concatenate([adj[ptr[n]:ptr[n + 1]] for n in nodes])) if len(nodes) else np.zeros(0, np.int64)
    sub_nodes = np.unique(conn[el].ravel())
    sub_conn = np.searchsorted(sub_nodes, conn[el]).astype(np.int32)
    return el, sub_nodes, sub_conn


def _oracle_rows(nodes, ptr, adj, conn, elem_comp, comps, X, u, bc_nodes, bc_mask, nonlinear,
                 want_g, fma=False):
    """oracle residual / K / G rows of `nodes` from the sub-mesh of their elements.
    bc_nodes None: no boundary conditions (raw contributions).  fma: the oracle build with fused
    multiply-adds — the same restatement of the reference's arithmetic under another legal
    compilation; the spread between the two is what the reference itself reproduces."""
    el, sub_nodes, sub_conn = _submesh(nodes, ptr, adj, conn)
    ec = np.ascontiguousarray(elem_comp[el], dtype=np.int32)
    Xs, us = X[sub_nodes], u[sub_nodes]
    rowp, cols = orc.pattern(len(sub_nodes), sub_conn)
    if bc_nodes is not None and len(bc_nodes):
        inb = np.isin(sub_nodes, bc_nodes)
        bn = np.nonzero(inb)[0].astype(np.int32)
        bv = np.full(len(bn), bc_mask, dtype=np.int32)
        bx = np.zeros((len(bn), 6))
    else:
        bn = bv = bx = None
    r, K = orc.assemble(1, sub_conn, ec, comps, Xs, us, rowp, cols, bn, bv, bx, fma=fma)
    G = None
    if want_g:
        _, G = orc.assemble(3, sub_conn, ec, comps, Xs, us, rowp, cols, bn, bv, bx, fma=fma)
    idx = np.searchsorted(sub_nodes, nodes)
    out = []
    for i in idx:
        a, b = rowp[i], rowp[i + 1]
        out.append((r[i], sub_nodes[cols[a:b]], K[a:b], None if G is None else G[a:b]))
    return out


def pick_nodes(n_owned, n_nodes, conn, bc_nodes, interface_nodes=(), n_interior=160, n_each=48,
               seed=0):
    """owned sample nodes: random interior ones, BC nodes and their neighbours, nodes next to
    ghost nodes, and (separately) interface nodes that other ranks hold as ghosts"""
    rng = np.random.default_rng(seed)
    interface = np.asarray(interface_nodes, dtype=np.int64)
    bc = np.asarray(bc_nodes, dtype=np.int64)
    touches_ghost = np.unique(conn[(conn >= n_owned).any(axis=1)].ravel()) if n_nodes > n_owned else np.zeros(0, np.int64)
    touches_ghost = touches_ghost[touches_ghost < n_owned]
    bc_adj = np.unique(conn[np.isin(conn, bc[:: max(1, len(bc) // 64)]).any(axis=1)].ravel()) if len(bc) else np.zeros(0, np.int64)
    bc_adj = bc_adj[bc_adj < n_owned]

    def some(a, k):
        a = np.asarray(a, dtype=np.int64)
        return a if len(a) <= k else rng.choice(a, size=k, replace=False)
    regular = np.unique(np.concatenate([
        rng.integers(0, n_owned, size=n_interior), some(bc, n_each), some(bc_adj, n_each),
        some(touches_ghost, n_each)])).astype(np.int64)
    regular = regular[~np.isin(regular, interface)]
    iface = some(interface[~np.isin(interface, bc)], n_each)
    return regular, np.sort(iface)


def check(asm, kmat, gmat, conn, X, u, elem_comp, comps, bc_nodes, n_owned, bc_mask=63,
          interface_nodes=(), glob=None, dist=None, nonlinear=False, n_interior=160, seed=0,
          res_dev=None):
    """Compare sampled node rows of the device residual / K / G with the oracle.
    asm: a2ds Assembler whose matrices kmat (and gmat, or None) and residual are ASSEMBLED
    for the state u (all local nodes, ghosts included).  res_dev: the device residual of the
    owned nodes [n_owned, 6] as a host array.  Returns a dict for the bench line / asserts."""
    conn = np.ascontiguousarray(conn, dtype=np.int64).reshape(-1, 4)
    n_nodes = len(X)
    ec = np.zeros(len(conn), dtype=np.int32) if elem_comp is None else np.asarray(elem_comp)
    ptr, adj = _node_to_elems(n_nodes, conn)
    regular, iface = pick_nodes(n_owned, n_nodes, conn, bc_nodes, interface_nodes, n_interior,
                                seed=seed)
    want_g = gmat is not None
    rowp, _ = asm.mat_pattern(kmat) if asm.mat_nnz(kmat) < 60_000_000 else (None, None)
    if rowp is None:
        import ctypes as C
        nr = C.c_int()
        rowp = np.zeros(n_nodes + 1, dtype=np.int32)
        asm._chk(asm.L.a2ds_mat_pattern(asm.ctx, C.c_int(kmat), C.c_int(0), C.byref(nr),
                                        rowp.ctypes.data_as(C.c_void_p), None))
    worst = dict(res=0.0, K=0.0, G=0.0)
    scale = dict(res=0.0, K=0.0, G=0.0)
    n_rows = 0

    def compare(nodes, rows, with_res, extra_res=None):
        nonlocal n_rows
        if not len(nodes):
            return
        Kd = asm.mat_rows(kmat, nodes, rowp)
        Gd = asm.mat_rows(gmat, nodes, rowp) if want_g else None
        for k, (n, (r_o, cols_o, K_o, G_o)) in enumerate(zip(nodes, rows)):
            assert Kd[k].shape == K_o.shape, "row pattern differs from the oracle's"
            worst["K"] = max(worst["K"], float(np.abs(Kd[k] - K_o).max()))
            scale["K"] = max(scale["K"], float(np.abs(K_o).max()))
            if want_g:
                worst["G"] = max(worst["G"], float(np.abs(Gd[k] - G_o).max()))
                scale["G"] = max(scale["G"], float(np.abs(G_o).max()))
            if with_res and res_dev is not None:
                ref = r_o if extra_res is None else r_o + extra_res.get(int(n), 0.0)
                worst["res"] = max(worst["res"], float(np.abs(res_dev[n] - ref).max()))
                scale["res"] = max(scale["res"], float(np.abs(ref).max()))
            n_rows += 1

    bc = np.asarray(bc_nodes, dtype=np.int64)
    rows_reg = _oracle_rows(regular, ptr, adj, conn, ec, comps, X, u, bc, bc_mask, nonlinear, want_g)
    compare(regular, rows_reg, True)
    # noise floor of the reference's own arithmetic on THIS mesh: the oracle compiled with and
    # without FMA contraction (two legal builds of the same source).  An element of size h at
    # distance |X| from the origin has its coordinate differences exact to eps |X| / h only, and
    # on the fine bench meshes (h = 1e-3 ... 1e-4 at |X| ~ 1 ... 8) the two builds differ by
    # 3e-13 ... 3e-12 in the residual — at or above the north-star 1e-12 by themselves.
    noise = dict(res=0.0, K=0.0, G=0.0)
    sub = regular[:: max(1, len(regular) // 96)]
    if len(sub):
        a = _oracle_rows(sub, ptr, adj, conn, ec, comps, X, u, bc, bc_mask, nonlinear, want_g)
        b = _oracle_rows(sub, ptr, adj, conn, ec, comps, X, u, bc, bc_mask, nonlinear, want_g, fma=True)
        for (r0, _, k0, g0), (r1, _, k1, g1) in zip(a, b):
            noise["res"] = max(noise["res"], float(np.abs(r0 - r1).max()))
            noise["K"] = max(noise["K"], float(np.abs(k0 - k1).max()))
            if want_g:
                noise["G"] = max(noise["G"], float(np.abs(g0 - g1).max()))
    # interface nodes: raw local contributions; the neighbours' share of the residual comes
    # from their oracle on the ghost copies
    extra = None
    if dist is not None and glob is not None:
        want = [None] * dist.get_world_size()
        dist.all_gather_object(want, [int(glob[n]) for n in iface])
        ghost_glob = glob[n_owned:]
        mine = {}
        asked = np.array(sorted({g for lst in want for g in lst}), dtype=np.int64)
        held = asked[np.isin(asked, ghost_glob)]
        if len(held):
            loc = n_owned + np.searchsorted(ghost_glob, held) if np.all(np.diff(ghost_glob) > 0) else \
                np.array([n_owned + int(np.nonzero(ghost_glob == g)[0][0]) for g in held])
            rows = _oracle_rows(loc, ptr, adj, conn, ec, comps, X, u, None, bc_mask, nonlinear, False)
            for g, row in zip(held, rows):
                mine[int(g)] = row[0]
        got = [None] * dist.get_world_size()
        dist.all_gather_object(got, mine)
        extra = {}
        for n in iface:
            g = int(glob[n])
            tot = 0.0
            for d in got:
                if g in d:
                    tot = tot + d[g]
            extra[int(n)] = tot
    if len(iface):
        compare(iface, _oracle_rows(iface, ptr, adj, conn, ec, comps, X, u, None, bc_mask,
                                    nonlinear, want_g), extra is not None or not len(interface_nodes),
                extra)
    rel = {k: (worst[k] / scale[k] if scale[k] > 0 else 0.0) for k in worst}
    nrel = {k: (noise[k] / scale[k] if scale[k] > 0 else 0.0) for k in worst}
    if not want_g:
        rel.pop("G"); nrel.pop("G")
    tol = dict(res=1e-12, K=1e-10, G=1e-10)
    # gate: the north-star tolerance, or 4 x the reference's own build-to-build spread on this
    # mesh where that is larger (residual of very fine meshes far from the origin; the spread is
    # taken on ~100 rows, the comparison on all sampled rows of all ranks)
    used = {k: max(tol[k], 4.0 * nrel[k]) for k in rel}
    return dict(rows=n_rows, interface_rows=int(len(iface)), max_rel=rel,
                tol={k: tol[k] for k in rel}, reference_fma_spread=nrel, tol_used=used,
                ok=bool(all(rel[k] <= used[k] for k in rel)),
                against="plain-C oracle (oracle/shell_oracle.c) on the elements around each sampled node")
