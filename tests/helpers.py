"""Shared helpers for the parity tests."""
import ctypes as C

import numpy as np


def relmax(a, b):
    """norm-relative error max|a-b| / max|b| (SURVEY.md §7 hard part 2)"""
    d = np.abs(np.asarray(a) - np.asarray(b)).max()
    s = np.abs(np.asarray(b)).max()
    return d / s if s > 0 else d


def random_elements(n, seed, h_range=(-3, 0), state_range=(-6, -3), flat=False):
    """n random distorted, curved, arbitrarily oriented quads with random states.
    Returns X[n,4,3], q[n,4,6]."""
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], float)
    X = np.zeros((n, 4, 3)); q = np.zeros((n, 4, 6))
    for e in range(n):
        h = 10.0 ** rng.uniform(*h_range)
        pert = 0.2 * rng.uniform(-1, 1, (4, 3))
        if flat:
            pert[:, 2] = 0.0
        Xe = h * (base + pert)
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        X[e] = Xe @ R.T + rng.uniform(-1, 1, 3)
        sc = 10.0 ** rng.uniform(*state_range)
        q[e, :, :3] = sc * h * rng.uniform(-1, 1, (4, 3))
        q[e, :, 3:] = sc * rng.uniform(-1, 1, (4, 3))
    return X, q


def random_elements9(n, seed, h_range=(-3, 0), state_range=(-6, -3)):
    """n random curved, distorted, arbitrarily oriented 9-node quads (3 x 3 nodes, xi fastest)
    with random states.  Returns X[n,9,3], q[n,9,6]."""
    rng = np.random.default_rng(seed)
    g = np.array([0.0, 0.5, 1.0])
    base = np.array([[a, b, 0.0] for b in g for a in g])
    X = np.zeros((n, 9, 3)); q = np.zeros((n, 9, 6))
    for e in range(n):
        h = 10.0 ** rng.uniform(*h_range)
        curv = rng.uniform(-0.3, 0.3, 3)
        bend = np.zeros((9, 3))
        bend[:, 2] = curv[0] * base[:, 0] ** 2 + curv[1] * base[:, 1] ** 2 + curv[2] * base[:, 0] * base[:, 1]
        Xe = h * (base + bend + 0.06 * rng.uniform(-1, 1, (9, 3)))
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        X[e] = Xe @ R.T + rng.uniform(-1, 1, 3)
        sc = 10.0 ** rng.uniform(*state_range)
        q[e, :, :3] = sc * h * rng.uniform(-1, 1, (9, 3))
        q[e, :, 3:] = sc * rng.uniform(-1, 1, (9, 3))
    return X, q


def emul_element(L, Cs, eth, T, model, transform, axis, X, q, want_g, tying=False):
    """one element through the host emulation of the kernel math; tying=True: the tying-level
    formulation of k_assemble_t (uncoupled sections only)"""
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    res = np.zeros(24); K = np.zeros(576); G = np.zeros(576)
    ax = np.asarray(axis, float)
    ax = np.ascontiguousarray(ax / np.linalg.norm(ax))
    Cs = np.ascontiguousarray(Cs, dtype=np.float64); eth = np.ascontiguousarray(eth, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64).ravel()
    q = np.ascontiguousarray(q, dtype=np.float64).ravel()
    fn = L.emul_element_t if tying else L.emul_element
    rc = fn(p(Cs), p(eth), C.c_double(T), C.c_int(model), C.c_int(transform), p(ax), p(X),
            p(q), C.c_int(want_g), p(res), p(K), p(G))
    assert rc == 0
    return res, K.reshape(24, 24), G.reshape(24, 24)


def element_blocks(A, rowp, cols, nodes):
    """extract the 24x24 of one element (4 node ids) from BCSR values A[nnz,6,6]"""
    out = np.zeros((24, 24))
    for i, r in enumerate(nodes):
        row = cols[rowp[r]:rowp[r + 1]]
        for j, c in enumerate(nodes):
            k = rowp[r] + int(np.searchsorted(row, c))
            assert cols[k] == c
            out[6 * i:6 * i + 6, 6 * j:6 * j + 6] = A[k]
    return out


def bcsr_to_dense(A, rowp, cols, n):
    M = np.zeros((6 * n, 6 * n))
    for r in range(n):
        for k in range(rowp[r], rowp[r + 1]):
            c = cols[k]
            M[6 * r:6 * r + 6, 6 * c:6 * c + 6] = A[k]
    return M


def bcsr_matvec(A, rowp, cols, x):
    """y = A x for x[n,6] without forming the dense matrix"""
    n = len(rowp) - 1
    rows = np.repeat(np.arange(n), np.diff(rowp))
    y = np.zeros((n, 6))
    np.add.at(y, rows, np.einsum("kij,kj->ki", A, x[cols]))
    return y


def with_dependent_nodes(conn, X, bc_nodes, n_dep, seed=0, npe=4):
    """Turn n_dep nodes of a mesh into DEPENDENT nodes (TACSAssembler::setDependentNodes): each
    is replaced by a weighted mean of the other nodes of its elements (up to 8 of them), the
    remaining nodes are renumbered compactly.  Returns (conn with -(d + 1) entries, X of the
    independent nodes, their BC node list, (dep_ptr, dep_conn, dep_weights))."""
    rng = np.random.default_rng(seed)
    conn = np.asarray(conn).reshape(-1, npe)
    n = len(X)
    banned = set(int(b) for b in bc_nodes)
    chosen, nbrs_of = [], {}
    node_elems = [[] for _ in range(n)]
    for e, nd in enumerate(conn):
        for v in nd:
            node_elems[int(v)].append(e)
    for v in rng.permutation(n):
        v = int(v)
        if len(chosen) == n_dep:
            break
        if v in banned or len(node_elems[v]) < 4:   # interior corner nodes: the patch surrounds them
            continue
        nb = sorted(set(int(w) for e in node_elems[v] for w in conn[e]) - {v})
        if any(w in nbrs_of for w in nb):      # a dependent node never depends on another one
            continue
        pick = nb     # the whole patch around the node: its weighted mean stays close to the node
        chosen.append(v); nbrs_of[v] = pick
        banned.update(nb); banned.add(v)
    assert len(chosen) == n_dep, "mesh too small for that many dependent nodes"
    keep = np.ones(n, bool); keep[chosen] = False
    new = -np.ones(n, dtype=np.int64); new[keep] = np.arange(keep.sum())
    for d, v in enumerate(chosen):
        new[v] = -(d + 1)
    dep_ptr = [0]; dep_conn = []; dep_w = []
    for v in chosen:
        w = rng.uniform(0.8, 1.2, len(nbrs_of[v])); w /= w.sum()
        dep_conn += [int(new[x]) for x in nbrs_of[v]]; dep_w += list(w)
        dep_ptr.append(len(dep_conn))
    conn2 = new[conn].astype(np.int32)
    bc2 = new[np.asarray(bc_nodes, dtype=np.int64)].astype(np.int32)
    assert (bc2 >= 0).all() and min(dep_conn) >= 0
    return (conn2, np.ascontiguousarray(X[keep]), bc2,
            (np.array(dep_ptr, np.int32), np.array(dep_conn, np.int32), np.array(dep_w)))
