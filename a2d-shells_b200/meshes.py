"""Synthetic structured shell meshes and seeded states.

Node order inside an element is the reference's tensor order
[(i,j),(i+1,j),(i,j+1),(i+1,j+1)] (TACSShellElementQuadBasis.h:147-150; the BDF
reader maps CQUAD4 n1,n2,n3,n4 to this as [n1,n2,n4,n3], TACSMeshLoader.cpp:932-937).
"""
import numpy as np


def plate(nx, ny, lx=1.0, ly=1.0, bump=1e-3):
    """Flat (bump=0) or slightly doubly-curved plate of nx x ny elements.
    Returns conn[ne,4], X[nn,3], clamped-edge BC nodes (edge i = 0)."""
    i = np.arange(nx + 1); j = np.arange(ny + 1)
    ii, jj = np.meshgrid(i, j, indexing="xy")          # node id = j*(nx+1) + i
    x = ii.ravel() * (lx / nx); y = jj.ravel() * (ly / ny)
    z = bump * np.sin(2 * np.pi * x / lx) * np.sin(2 * np.pi * y / ly)
    X = np.stack([x, y, z], axis=1)
    ei, ej = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    n0 = (ej * (nx + 1) + ei).ravel()
    conn = np.stack([n0, n0 + 1, n0 + nx + 1, n0 + nx + 2], axis=1).astype(np.int32)
    bc_nodes = (np.arange(ny + 1) * (nx + 1)).astype(np.int32)
    return conn, X, bc_nodes


def cylinder(ntheta, nx, radius=0.2, length=0.4, x0=0.0):
    """Closed cylinder, periodic in theta: ntheta x nx elements, X = (x, -R sin t, -R cos t)
    as examples/cylinder-buckling/mech-cylinder.bdf:7-10.  Node id = ix*ntheta + it.
    Returns conn, X, nodes of the two end rings."""
    it = np.arange(ntheta); ix = np.arange(nx + 1)
    tt, xx = np.meshgrid(it, ix, indexing="xy")
    th = tt.ravel() * (2 * np.pi / ntheta)
    x = x0 + xx.ravel() * (length / nx)
    X = np.stack([x, -radius * np.sin(th), -radius * np.cos(th)], axis=1)
    et, ex = np.meshgrid(np.arange(ntheta), np.arange(nx), indexing="xy")
    et = et.ravel(); ex = ex.ravel()
    etp = (et + 1) % ntheta
    conn = np.stack([ex * ntheta + et, ex * ntheta + etp, (ex + 1) * ntheta + et,
                     (ex + 1) * ntheta + etp], axis=1).astype(np.int32)
    ends = np.concatenate([np.arange(ntheta), nx * ntheta + np.arange(ntheta)]).astype(np.int32)
    return conn, X, ends


def plate9(nx, ny, lx=1.0, ly=1.0, bump=1e-3):
    """The same plate meshed with nx x ny 9-node elements (TACSQuad9Shell): a (2 nx + 1) x
    (2 ny + 1) node grid, element node order = tensor order of TACSShellQuadBasis<3>
    (TACSShellElementQuadBasis.h:147-150: xi fastest, 3 x 3).  Returns conn[ne,9], X, BC nodes."""
    mx, my = 2 * nx + 1, 2 * ny + 1
    ii, jj = np.meshgrid(np.arange(mx), np.arange(my), indexing="xy")   # node id = j*mx + i
    x = ii.ravel() * (lx / (mx - 1)); y = jj.ravel() * (ly / (my - 1))
    z = bump * np.sin(2 * np.pi * x / lx) * np.sin(2 * np.pi * y / ly)
    X = np.stack([x, y, z], axis=1)
    ei, ej = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    n0 = (2 * ej * mx + 2 * ei).ravel()
    conn = np.stack([n0 + b * mx + a for b in range(3) for a in range(3)], axis=1).astype(np.int32)
    bc_nodes = (np.arange(my) * mx).astype(np.int32)
    return conn, X, bc_nodes


def cylinder9(ntheta, nx, radius=0.2, length=0.4):
    """Closed cylinder of ntheta x nx 9-node elements (2 ntheta nodes around, 2 nx + 1 along);
    nodes ON the cylinder (the mid-side nodes too).  Returns conn[ne,9], X, end-ring nodes."""
    mt, mx = 2 * ntheta, 2 * nx + 1
    tt, xx = np.meshgrid(np.arange(mt), np.arange(mx), indexing="xy")   # node id = ix*mt + it
    th = tt.ravel() * (2 * np.pi / mt)
    x = xx.ravel() * (length / (mx - 1))
    X = np.stack([x, -radius * np.sin(th), -radius * np.cos(th)], axis=1)
    et, ex = np.meshgrid(np.arange(ntheta), np.arange(nx), indexing="xy")
    et = et.ravel(); ex = ex.ravel()
    conn = np.stack([(2 * ex + b) * mt + (2 * et + a) % mt for b in range(3) for a in range(3)],
                    axis=1).astype(np.int32)
    ends = np.concatenate([np.arange(mt), (mx - 1) * mt + np.arange(mt)]).astype(np.int32)
    return conn, X, ends


def cubed_sphere(n, radius=0.3, shuffle_seed=None):
    """Closed sphere meshed from the six faces of a cube (n x n quads each): an UNSTRUCTURED
    quad mesh — the 8 cube corners have valence 3, no global (i, j) numbering exists.
    Element node order is the tensor order [(0,0),(1,0),(0,1),(1,1)] of each face's own
    (a, b) parametrisation, oriented so that every element normal points outwards.
    shuffle_seed: additionally permute node numbers and element order at random.
    Returns conn, X, and the nodes of one small patch to clamp."""
    faces = []  # (origin, da, db) on the cube [-1, 1]^3, da x db pointing outwards
    for axis in range(3):
        for sign in (-1.0, 1.0):
            a = np.zeros(3); b = np.zeros(3); o = np.zeros(3)
            a[(axis + 1) % 3] = 1.0; b[(axis + 2) % 3] = 1.0; o[axis] = sign
            if sign < 0:
                a, b = b, a
            faces.append((o - a - b, 2.0 * a, 2.0 * b))
    key = {}
    pts = []
    conn = []
    t = np.arange(n + 1) / n
    for o, da, db in faces:
        ids = np.zeros((n + 1, n + 1), dtype=np.int64)
        for j in range(n + 1):
            for i in range(n + 1):
                p = o + t[i] * da + t[j] * db
                k = tuple(np.round(p * n).astype(np.int64))   # exact on the cube lattice
                if k not in key:
                    key[k] = len(pts)
                    pts.append(p)
                ids[j, i] = key[k]
        for j in range(n):
            for i in range(n):
                conn.append([ids[j, i], ids[j, i + 1], ids[j + 1, i], ids[j + 1, i + 1]])
    P = np.array(pts)
    X = radius * P / np.linalg.norm(P, axis=1)[:, None]
    conn = np.array(conn, dtype=np.int32)
    if shuffle_seed is not None:
        rng = np.random.default_rng(shuffle_seed)
        perm = rng.permutation(len(X))            # new id of old node i
        Xn = np.zeros_like(X); Xn[perm] = X
        X = Xn
        conn = perm[conn].astype(np.int32)[rng.permutation(len(conn))]
    patch = np.unique(conn[:max(1, n // 2)].ravel()).astype(np.int32)
    return conn, X, patch


def wingbox(n_bays=6, nx=3, ny=4, nz=2, span=3.0, chord=1.0, height=0.25):
    """Synthetic wing box (BASELINE configs[3] stand-in: the uCRM mesh itself is not part of
    the reference checkout): upper and lower skin, front and rear spar, and a rib at every bay
    boundary, all sharing their nodes where they meet — skin/spar/rib junction nodes have 5 or
    6 elements around them, so matrix rows hold up to 15 blocks instead of the 9 of a
    structured surface.  Tapered, swept, skins slightly cambered.  n_bays bays of nx elements
    along the span, ny across the chord, nz through the height.
    Returns conn[ne,4] (tensor node order), X[nn,3], elem_comp[ne] (one component per skin /
    spar panel per bay and per rib: 4 n_bays + n_bays + 1), root nodes (clamped end)."""
    NX = n_bays * nx
    ids = {}
    X = []

    def node(i, j, k):
        key = (i, j, k)
        if key not in ids:
            ids[key] = len(X)
            s, c, h = i / NX, j / ny, k / nz
            taper = 1.0 - 0.5 * s
            camber = 0.04 * chord * np.sin(np.pi * c) * (1.0 if k == nz else -1.0 if k == 0 else 0.0)
            X.append((span * s,
                      0.35 * span * s + chord * taper * (c - 0.5),
                      0.03 * span * s * s + height * taper * (h - 0.5) + camber * taper))
        return ids[key]

    conn, comp = [], []

    def quad(a, b, c, d, cid):        # tensor order: (0,0) (1,0) (0,1) (1,1)
        conn.append((a, b, c, d)); comp.append(cid)

    for bay in range(n_bays):
        for i in range(bay * nx, (bay + 1) * nx):
            for j in range(ny):       # skins: k = nz (upper), k = 0 (lower)
                quad(node(i, j, nz), node(i + 1, j, nz), node(i, j + 1, nz), node(i + 1, j + 1, nz), 4 * bay)
                quad(node(i, j, 0), node(i, j + 1, 0), node(i + 1, j, 0), node(i + 1, j + 1, 0), 4 * bay + 1)
            for k in range(nz):       # spars: j = 0 (front), j = ny (rear)
                quad(node(i, 0, k), node(i + 1, 0, k), node(i, 0, k + 1), node(i + 1, 0, k + 1), 4 * bay + 2)
                quad(node(i, ny, k), node(i, ny, k + 1), node(i + 1, ny, k), node(i + 1, ny, k + 1), 4 * bay + 3)
    for rib in range(n_bays + 1):
        i = rib * nx
        for j in range(ny):
            for k in range(nz):
                quad(node(i, j, k), node(i, j + 1, k), node(i, j, k + 1), node(i, j + 1, k + 1),
                     4 * n_bays + rib)
    root = sorted(v for (i, j, k), v in ids.items() if i == 0)
    return (np.array(conn, dtype=np.int32), np.array(X), np.array(comp, dtype=np.int32),
            np.array(root, dtype=np.int32))


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def seeded_state(global_node_ids, scale=1e-5, seed=12345):
    """u[n,6] = scale * U(-1,1) from a counter-based hash keyed on the GLOBAL node id
    and dof (partition independent, no RNG state): splitmix64(seed ^ (6*id + dof))."""
    ids = np.asarray(global_node_ids, dtype=np.uint64)
    key = (ids[:, None] * np.uint64(6) + np.arange(6, dtype=np.uint64)[None, :]) ^ np.uint64(seed)
    with np.errstate(over="ignore"):
        h = _splitmix64(key)
    u01 = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return scale * (2.0 * u01 - 1.0)


def plate_slab(rank, n_ranks, nx, ny, lx=1.0, ly=1.0, bump=1e-3):
    """Rank `rank`'s slab of a plate of nx x (n_ranks*ny) elements partitioned by element
    rows, built directly (no global mesh): the weak-scaling workload of bench.py.
    Ownership is TACSCreator's first-touch rule (src/TACSCreator.cpp:1156-1205): the
    node row shared by slabs r-1 and r belongs to r-1, so rank r > 0 sees it as ghosts.
    Local numbering: owned nodes in global order, then ghosts.  Returns a dict with
    conn, X, n_owned, glob (global node ids), bc_nodes (edge i = 0), peers, send_lists,
    recv_lists (local indices)."""
    npr = nx + 1
    j0 = rank * ny
    first_owned_row = j0 if rank == 0 else j0 + 1
    owned_rows = np.arange(first_owned_row, j0 + ny + 1)
    n_owned = len(owned_rows) * npr
    glob_owned = (owned_rows[:, None] * npr + np.arange(npr)[None, :]).ravel()
    glob_ghost = (j0 * npr + np.arange(npr)) if rank > 0 else np.zeros(0, dtype=np.int64)
    glob = np.concatenate([glob_owned, glob_ghost]).astype(np.int64)

    def local(row, i):
        row = np.asarray(row); i = np.asarray(i)
        own = (row - first_owned_row) * npr + i
        return np.where(row >= first_owned_row, own, n_owned + i)

    ei, ej = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ei = ei.ravel(); ej = ej.ravel() + j0
    conn = np.stack([local(ej, ei), local(ej, ei + 1), local(ej + 1, ei), local(ej + 1, ei + 1)],
                    axis=1).astype(np.int32)
    gi = glob % npr; gj = glob // npr
    Ly = ly * n_ranks
    x = gi * (lx / nx); y = gj * (Ly / (ny * n_ranks))
    z = bump * np.sin(2 * np.pi * x / lx) * np.sin(2 * np.pi * y / ly)
    X = np.stack([x, y, z], axis=1)
    bc_nodes = np.nonzero(gi[:n_owned] == 0)[0].astype(np.int32)
    peers, send_lists, recv_lists = [], [], []
    if rank > 0:      # ghost row comes from the slab below; nothing is sent to it
        peers.append(rank - 1)
        send_lists.append(np.zeros(0, dtype=np.int32))
        recv_lists.append((n_owned + np.arange(npr)).astype(np.int32))
    if rank < n_ranks - 1:  # our top row is the next slab's ghost row
        peers.append(rank + 1)
        send_lists.append(local(np.full(npr, j0 + ny), np.arange(npr)).astype(np.int32))
        recv_lists.append(np.zeros(0, dtype=np.int32))
    return dict(conn=conn, X=X, n_owned=n_owned, n_nodes=len(glob), glob=glob, bc_nodes=bc_nodes,
                peers=np.asarray(peers, dtype=np.int32), send_lists=send_lists,
                recv_lists=recv_lists)


def cylinder_slab(rank, n_ranks, ntheta, nx, radius=0.2, length=0.4):
    """Rank `rank`'s slab of a closed cylinder of ntheta x (n_ranks*nx) elements cut into
    axial slabs (BASELINE configs[2]/[4]: 2000x2000 on 1-8 GPUs, 4000x4000 = 16 M elements
    on 8).  Same conventions as plate_slab / cylinder: node id = ring*ntheta + it, periodic in
    theta, first-touch ownership (the ring shared with the slab below belongs to it)."""
    i0 = rank * nx
    first_owned = i0 if rank == 0 else i0 + 1
    rings = np.arange(first_owned, i0 + nx + 1)
    n_owned = len(rings) * ntheta
    glob_owned = (rings[:, None] * ntheta + np.arange(ntheta)[None, :]).ravel()
    glob_ghost = (i0 * ntheta + np.arange(ntheta)) if rank > 0 else np.zeros(0, dtype=np.int64)
    glob = np.concatenate([glob_owned, glob_ghost]).astype(np.int64)

    def local(ring, it):
        ring = np.asarray(ring); it = np.asarray(it)
        own = (ring - first_owned) * ntheta + it
        return np.where(ring >= first_owned, own, n_owned + it)

    et, ex = np.meshgrid(np.arange(ntheta), np.arange(nx), indexing="xy")
    et = et.ravel(); ex = ex.ravel() + i0
    etp = (et + 1) % ntheta
    conn = np.stack([local(ex, et), local(ex, etp), local(ex + 1, et), local(ex + 1, etp)],
                    axis=1).astype(np.int32)
    it = glob % ntheta; ix = glob // ntheta
    th = it * (2 * np.pi / ntheta)
    x = ix * (length / (nx * n_ranks))
    X = np.stack([x, -radius * np.sin(th), -radius * np.cos(th)], axis=1)
    ends = np.nonzero((ix[:n_owned] == 0) | (ix[:n_owned] == nx * n_ranks))[0].astype(np.int32)
    peers, send_lists, recv_lists = [], [], []
    if rank > 0:
        peers.append(rank - 1)
        send_lists.append(np.zeros(0, dtype=np.int32))
        recv_lists.append((n_owned + np.arange(ntheta)).astype(np.int32))
    if rank < n_ranks - 1:
        peers.append(rank + 1)
        send_lists.append(local(np.full(ntheta, i0 + nx), np.arange(ntheta)).astype(np.int32))
        recv_lists.append(np.zeros(0, dtype=np.int32))
    return dict(conn=conn, X=X, n_owned=n_owned, n_nodes=len(glob), glob=glob, bc_nodes=ends,
                peers=np.asarray(peers, dtype=np.int32), send_lists=send_lists,
                recv_lists=recv_lists)


def partition_rows(conn, n_nodes, elem_rank, matrix_halo=False):
    """Element-wise partition -> per-rank local meshes with TACSCreator's node ownership
    rule: a node belongs to the rank of the first element (in global order) touching it
    (src/TACSCreator.cpp:1156-1205).  Returns a list of dicts per rank:
      conn_local, owned (global ids), ghosts (global ids), elems (global element ids),
      peers / send_lists / recv_lists (local indices) for the halo exchange.
    matrix_halo=True prepares the TACSParallelMat flavour as well: the ghosts of a rank then
    also include every node coupled to one of its owned nodes through ANOTHER rank's
    elements (the reference's external column map), and each dict gains
      rowp / cols                     local non-zero pattern: full global rows for the owned
                                      nodes, the local elements' couplings for ghost rows
      mat_send_lists / mat_recv_lists per peer (same order as peers): block indices of my
                                      ghost rows to send / of the blocks arriving ones add to,
                                      both sorted by (global row, global column)."""
    conn = np.asarray(conn)
    elem_rank = np.asarray(elem_rank)
    n_ranks = int(elem_rank.max()) + 1
    owner = np.full(n_nodes, -1, dtype=np.int64)
    # first touch in global element order
    flat_nodes = conn.ravel()
    flat_rank = np.repeat(elem_rank, 4)
    first = np.full(n_nodes, len(flat_nodes), dtype=np.int64)
    np.minimum.at(first, flat_nodes, np.arange(len(flat_nodes)))
    touched = first < len(flat_nodes)
    owner[touched] = flat_rank[first[touched]]
    parts = []
    for r in range(n_ranks):
        elems = np.nonzero(elem_rank == r)[0]
        used = np.unique(conn[elems].ravel())
        owned = used[owner[used] == r]
        # owned nodes this rank never references still belong to it
        extra = np.nonzero(owner == r)[0]
        owned = np.union1d(owned, extra)
        ghosts = used[owner[used] != r]
        if matrix_halo:
            # column nodes of the owned rows that only other ranks' elements bring in
            grp, gcl = _global_pattern(n_nodes, conn)
            coupled = np.unique(np.concatenate([gcl[grp[g]:grp[g + 1]] for g in owned])) \
                if len(owned) else np.zeros(0, dtype=np.int64)
            ghosts = np.union1d(ghosts, np.setdiff1d(coupled, owned))
        glob = np.concatenate([owned, ghosts])
        lookup = {}
        local_of = np.full(n_nodes, -1, dtype=np.int64)
        local_of[glob] = np.arange(len(glob))
        parts.append(dict(rank=r, elems=elems, owned=owned, ghosts=ghosts, glob=glob,
                          conn_local=local_of[conn[elems]].astype(np.int32),
                          local_of=local_of, ghost_owner=owner[ghosts],
                          ghost_owner_all=owner[ghosts]))
    # halo lists: for each (r, p): ghosts of r owned by p, sorted by global id on both sides
    for r in range(n_ranks):
        P = parts[r]
        peers = sorted(set(P["ghost_owner"].tolist()) |
                       {q for q in range(n_ranks) if q != r and np.any(parts[q]["ghost_owner"] == r)})
        send_lists, recv_lists = [], []
        for p in peers:
            mine = np.sort(P["ghosts"][P["ghost_owner"] == p])              # I receive these
            theirs = np.sort(parts[p]["ghosts"][parts[p]["ghost_owner"] == r])  # I send these
            recv_lists.append(P["local_of"][mine].astype(np.int32))
            send_lists.append(P["local_of"][theirs].astype(np.int32))
        P["peers"] = np.asarray(peers, dtype=np.int32)
        P["send_lists"] = send_lists
        P["recv_lists"] = recv_lists
    if matrix_halo:
        grp, gcl = _global_pattern(n_nodes, conn)
        for P in parts:
            # local pattern: couplings of the local elements, plus the full global row of
            # every owned node (all its columns are local by construction)
            nl = len(P["glob"])
            rows = [set() for _ in range(nl)]
            for e in P["conn_local"]:
                for a in e:
                    rows[a].update(int(b) for b in e)
            for l, g in enumerate(P["owned"]):
                rows[l].update(int(x) for x in P["local_of"][gcl[grp[g]:grp[g + 1]]])
            rowp = np.zeros(nl + 1, dtype=np.int32)
            rowp[1:] = np.cumsum([len(x) for x in rows])
            P["rowp"] = rowp
            P["cols"] = np.array([c for x in rows for c in sorted(x)], dtype=np.int32)
            # blocks of ghost rows that carry local element contributions, by owner
            P["_ghost_blocks"] = {}
            for l in range(len(P["owned"]), nl):
                elem_cols = set()
                for e in P["conn_local"]:
                    if l in e:
                        elem_cols.update(int(b) for b in e)
                if not elem_cols:
                    continue
                own = int(P["ghost_owner_all"][l - len(P["owned"])])
                lst = P["_ghost_blocks"].setdefault(own, [])
                for k in range(rowp[l], rowp[l + 1]):
                    if int(P["cols"][k]) in elem_cols:
                        lst.append((int(P["glob"][l]), int(P["glob"][P["cols"][k]]), k))
        for P in parts:
            P["mat_send_lists"] = []; P["mat_recv_lists"] = []
            for p in P["peers"]:
                mine = sorted(P["_ghost_blocks"].get(int(p), []))
                P["mat_send_lists"].append(np.array([k for _, _, k in mine], dtype=np.int32))
                Q = parts[int(p)]
                theirs = sorted(Q["_ghost_blocks"].get(P["rank"], []))
                dest = []
                for grow, gcol, _ in theirs:
                    lr, lc = int(P["local_of"][grow]), int(P["local_of"][gcol])
                    seg = P["cols"][P["rowp"][lr]:P["rowp"][lr + 1]]
                    k = int(np.searchsorted(seg, lc))
                    assert k < len(seg) and seg[k] == lc, "owner pattern misses an arriving block"
                    dest.append(P["rowp"][lr] + k)
                P["mat_recv_lists"].append(np.array(dest, dtype=np.int32))
    return parts


_pattern_cache = {}


def _global_pattern(n_nodes, conn):
    """node-to-node couplings of the whole mesh (rowp, cols), cached per call site"""
    key = (n_nodes, conn.shape[0], int(conn[:, 0].sum()), int(conn[-1, -1]))
    if key not in _pattern_cache:
        rows = [set() for _ in range(n_nodes)]
        for e in np.asarray(conn):
            for a in e:
                rows[a].update(int(b) for b in e)
        rowp = np.zeros(n_nodes + 1, dtype=np.int64)
        rowp[1:] = np.cumsum([len(x) for x in rows])
        cols = np.array([c for x in rows for c in sorted(x)], dtype=np.int64)
        _pattern_cache.clear()
        _pattern_cache[key] = (rowp, cols)
    return _pattern_cache[key]


# ---- NASTRAN bulk-data writer (the input format of the reference's examples) -------------
def _real8(x):
    """a real in 8 columns, NASTRAN compact form ("1.2345-3": exponent sign without a letter),
    with as many digits as fit"""
    if x == 0.0:
        return "      0."
    for digits in range(5, -1, -1):
        m, e = f"{x:.{digits}E}".split("E")
        if digits == 0:
            m += "."            # NASTRAN reals carry a decimal point
        s = f"{m}{int(e):+d}"
        if len(s) <= 8:
            return s.rjust(8)
    raise ValueError(f"{x!r} does not fit an 8-column field")


def write_bdf(path, conn, X, bc_nodes=(), bc_dofs=(), bc_vals=None, elem_comp=None, fmt="large",
              node_ids=None, elem_ids=None, comp_names=None, keyword="CQUAD4", order="file"):
    """Write a deck the reference's loader (and a2ds_mesh_read_bdf) reads.

    conn: (n_elems, 4) in the tensor order used everywhere in this package; the card lists the
    corners counter-clockwise, i.e. [n0, n1, n3, n2].  bc_dofs: per BC entry a string of DOF
    digits ("123456") or a list of 0-based DOFs; bc_vals: one value per entry.  fmt: "large"
    (GRID*, SPC*, 16-column fields, full double precision as the shipped decks), "small"
    (8-column fields) or "free" (comma separated GRID, small elements).  node_ids / elem_ids:
    the numbers written in the file (1-based, any order, gaps allowed)."""
    conn = np.asarray(conn).reshape(-1, 4)
    X = np.asarray(X, dtype=float).reshape(-1, 3)
    nid = np.arange(1, len(X) + 1) if node_ids is None else np.asarray(node_ids)
    eid = np.arange(1, len(conn) + 1) if elem_ids is None else np.asarray(elem_ids)
    comp = np.zeros(len(conn), dtype=int) if elem_comp is None else np.asarray(elem_comp)
    out = ["$ written by a2d-shells_b200.meshes.write_bdf", "SOL 103", "CEND", "BEGIN BULK"]
    for name in (comp_names or []):
        out.append("$       Shell element data for family".ljust(41) + str(name)[:32])
    for k in range(len(X)):
        x, y, z = X[k]
        if fmt == "large":
            out.append(f"GRID*   {nid[k]:>16d}{'':16s}{x:16.9E}{y:16.9E}*       ")
            out.append(f"*       {z:16.9E}")
        elif fmt == "small":
            out.append(f"GRID    {nid[k]:>8d}{'':8s}{_real8(x)}{_real8(y)}{_real8(z)}")
        else:
            out.append(f"GRID    {nid[k]},,{x:.10E},{y:.10E},{z:.10E}")
    for e in range(len(conn)):
        n = nid[conn[e][[0, 1, 3, 2]]]
        if fmt == "large":
            out.append(f"{keyword + '*':<8s}{eid[e]:>16d}{comp[e] + 1:>16d}{n[0]:>16d}{n[1]:>16d}")
            out.append(f"*       {n[2]:>16d}{n[3]:>16d}")
        else:
            out.append(f"{keyword:<8s}{eid[e]:>8d}{comp[e] + 1:>8d}{n[0]:>8d}{n[1]:>8d}{n[2]:>8d}{n[3]:>8d}")
    vals = np.zeros(len(bc_nodes)) if bc_vals is None else np.broadcast_to(bc_vals, (len(bc_nodes),))
    for b in range(len(bc_nodes)):
        d = bc_dofs[b] if isinstance(bc_dofs[b], str) else "".join(str(int(v) + 1) for v in bc_dofs[b])
        node = nid[bc_nodes[b]]
        if fmt == "large":
            v = "0." if vals[b] == 0 else f"{vals[b]:.9E}"
            out.append(f"SPC*    {1:>16d}{node:>16d}{d:>16s}{v:>16s}")
        else:
            out.append(f"SPC     {1:>8d}{node:>8d}{d:>8s}{_real8(vals[b])}")
    out.append("ENDDATA")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
