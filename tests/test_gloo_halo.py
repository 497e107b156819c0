"""world_size-2 gloo test (CPU) of the host-side multi-rank logic: every rank builds its slab
with plate_slab, the halo plans are exercised with real inter-process sends, and the ghost
values / reverse-added contributions are checked by global node id."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    import importlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    a2ds = importlib.import_module("a2d-shells_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny = 6, 4
    s = a2ds.meshes.plate_slab(rank, world, nx, ny)
    n_owned, n = s["n_owned"], s["n_nodes"]
    # forward: owners -> ghosts.  value = global id (so the check is by global id)
    vec = np.full((n, 6), -1.0)
    vec[:n_owned] = s["glob"][:n_owned, None] + 0.1 * np.arange(6)[None, :]
    reqs = []; bufs = []
    for p, sl, rl in zip(s["peers"], s["send_lists"], s["recv_lists"]):
        if len(sl):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(vec[sl])), int(p)))
        if len(rl):
            b = torch.empty((len(rl), 6), dtype=torch.float64); bufs.append((rl, b))
            reqs.append(dist.irecv(b, int(p)))
    for r in reqs:
        r.wait()
    for rl, b in bufs:
        vec[rl] = b.numpy()
    ok_fwd = np.array_equal(vec, s["glob"][:, None] + 0.1 * np.arange(6)[None, :])
    # reverse: ghost contributions added to owners.  every rank contributes 1 per element
    # corner touching a node; summed over ranks the owners must hold the node's valence
    contrib = np.zeros((n, 6))
    np.add.at(contrib, s["conn"].ravel(), 1.0)
    reqs = []; bufs = []
    for p, sl, rl in zip(s["peers"], s["send_lists"], s["recv_lists"]):
        if len(rl):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(contrib[rl])), int(p)))
        if len(sl):
            b = torch.empty((len(sl), 6), dtype=torch.float64); bufs.append((sl, b))
            reqs.append(dist.irecv(b, int(p)))
    for r in reqs:
        r.wait()
    for sl, b in bufs:
        contrib[sl] += b.numpy()
    gi = s["glob"][:n_owned] % (nx + 1); gj = s["glob"][:n_owned] // (nx + 1)
    val = (1 + ((gi > 0) & (gi < nx))) * (1 + ((gj > 0) & (gj < ny * world)))
    ok_rev = np.array_equal(contrib[:n_owned, 0], val.astype(float))
    ret[rank] = (bool(ok_fwd), bool(ok_rev), int(n_owned))
    dist.destroy_process_group()


def test_halo_plan_two_ranks_gloo():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29400 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r][0] and ret[r][1], ret[r]
    # every node of the 6 x 8 plate is owned exactly once
    assert sum(ret[r][2] for r in range(world)) == 7 * 9


def _exchange(dist, torch, s, vec, reverse):
    """the halo of a2ds_set_halo with gloo sends: forward owners -> ghosts (overwrite),
    reverse ghosts -> owners (add)"""
    reqs = []; bufs = []
    for p, sl, rl in zip(s["peers"], s["send_lists"], s["recv_lists"]):
        out, inn = (rl, sl) if reverse else (sl, rl)
        if len(out):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(vec[out])), int(p)))
        if len(inn):
            b = torch.empty((len(inn), 6), dtype=torch.float64); bufs.append((inn, b))
            reqs.append(dist.irecv(b, int(p)))
    for r in reqs:
        r.wait()
    for idx, b in bufs:
        if reverse:
            vec[idx] += b.numpy()
        else:
            vec[idx] = b.numpy()


def _matvec_worker(rank, world, port, ret):
    """the algorithm of a2ds_mat_mult_dist_dev on the CPU: per-rank matrices assembled by the
    oracle over the local nodes (interface rows unassembled), x forward, local product,
    y reverse-add, BC rows y = x"""
    import importlib
    import torch
    import torch.distributed as dist
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    a2ds = importlib.import_module("a2d-shells_b200")
    import oracle_py as orc
    from helpers import bcsr_matvec
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny = 5, 3
    s = a2ds.meshes.plate_slab(rank, world, nx, ny, bump=3e-2)
    n_owned, n = s["n_owned"], s["n_nodes"]
    Cs, eth = a2ds.iso_shell_tables()
    comp = orc.make_comp(0, Cs, eth)
    rowp, cols = orc.pattern(n, s["conn"])
    bcn = np.asarray(s["bc_nodes"], dtype=np.int32)
    bv = np.full(len(bcn), 63, dtype=np.int32)
    _, K = orc.assemble(2, s["conn"], np.zeros(len(s["conn"]), dtype=np.int32), [comp], s["X"],
                        np.zeros((n, 6)), rowp, cols, bcn, bv, np.zeros((len(bcn), 6)))
    x = np.full((n, 6), np.nan)
    x[:n_owned] = a2ds.meshes.seeded_state(s["glob"][:n_owned] + 99, 1.0)
    _exchange(dist, torch, s, x, reverse=False)
    y = bcsr_matvec(K, rowp, cols, x)
    _exchange(dist, torch, s, y, reverse=True)
    own_bc = bcn[bcn < n_owned]
    y[own_bc] = x[own_bc]
    ret[rank] = (s["glob"][:n_owned].copy(), y[:n_owned].copy())
    dist.destroy_process_group()


def test_distributed_matvec_two_ranks_gloo():
    import importlib
    import torch.multiprocessing as mp
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    a2ds = importlib.import_module("a2d-shells_b200")
    import oracle_py as orc
    from helpers import bcsr_matvec
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29650 + (os.getpid() % 200)
    mp.spawn(_matvec_worker, args=(world, port, ret), nprocs=world, join=True)
    nx, ny = 5, 3
    conn, X, bcn = a2ds.meshes.plate(nx, ny * world, ly=1.0 * world, bump=3e-2)
    n = len(X)
    Xs = np.zeros_like(X)
    for r in range(world):   # the geometry exactly as the slabs see it
        s = a2ds.meshes.plate_slab(r, world, nx, ny, bump=3e-2)
        Xs[s["glob"]] = s["X"]
    Cs, eth = a2ds.iso_shell_tables()
    rowp, cols = orc.pattern(n, conn)
    bv = np.full(len(bcn), 63, dtype=np.int32)
    _, K = orc.assemble(2, conn, np.zeros(len(conn), dtype=np.int32), [orc.make_comp(0, Cs, eth)],
                        Xs, np.zeros((n, 6)), rowp, cols, bcn, bv, np.zeros((len(bcn), 6)))
    y_all = bcsr_matvec(K, rowp, cols, a2ds.meshes.seeded_state(np.arange(n) + 99, 1.0))
    seen = 0
    for r in range(world):
        glob, y = ret[r]
        assert np.abs(y - y_all[glob]).max() < 1e-13 * np.abs(y_all).max()
        seen += len(glob)
    assert seen == n


def _partition_worker(rank, world, port, ret):
    """the native planner (a2ds_partition_build) end to end on the CPU: every rank derives its
    sub-mesh and halo lists from the global wing-box mesh on its own, ghost states arrive by
    real sends, the oracle assembles the local residual, ghost rows are reverse-added"""
    import importlib
    import torch
    import torch.distributed as dist
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    a2ds = importlib.import_module("a2d-shells_b200")
    import oracle_py as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    conn, X, comp, root = a2ds.meshes.wingbox(5, 2, 3, 2)
    n = len(X)
    elem_rank = ((np.arange(len(conn)) * 7919) % 97) % world      # scattered: every pair talks
    P = a2ds.Partition(conn, n, elem_rank, world, rank)
    s = dict(peers=P.peers, send_lists=P.send_lists, recv_lists=P.recv_lists)
    Cs, eth = a2ds.iso_shell_tables()
    comps = [orc.make_comp(0, Cs * (1 + 0.2 * c), eth) for c in range(int(comp.max()) + 1)]
    u = np.full((P.n_nodes, 6), np.nan)
    u[:P.n_owned] = a2ds.meshes.seeded_state(P.glob[:P.n_owned], 1e-4)
    _exchange(dist, torch, s, u, reverse=False)
    assert not np.isnan(u).any()
    rowp, cols = orc.pattern(P.n_nodes, P.conn_local)
    none = np.zeros(0, dtype=np.int32)
    r, _ = orc.assemble(1, P.conn_local, comp[P.elems], comps, X[P.glob], u, rowp, cols, none, none,
                        np.zeros((0, 6)))
    _exchange(dist, torch, s, r, reverse=True)
    ret[rank] = (P.glob[:P.n_owned].copy(), r[:P.n_owned].copy())
    dist.destroy_process_group()


def test_native_partition_three_ranks_gloo():
    import importlib
    import torch.multiprocessing as mp
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    a2ds = importlib.import_module("a2d-shells_b200")
    import oracle_py as orc
    world = 3
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29850 + (os.getpid() % 100)
    mp.spawn(_partition_worker, args=(world, port, ret), nprocs=world, join=True)
    conn, X, comp, root = a2ds.meshes.wingbox(5, 2, 3, 2)
    n = len(X)
    Cs, eth = a2ds.iso_shell_tables()
    comps = [orc.make_comp(0, Cs * (1 + 0.2 * c), eth) for c in range(int(comp.max()) + 1)]
    rowp, cols = orc.pattern(n, conn)
    none = np.zeros(0, dtype=np.int32)
    r_all, _ = orc.assemble(1, conn, comp, comps, X, a2ds.meshes.seeded_state(np.arange(n), 1e-4),
                            rowp, cols, none, none, np.zeros((0, 6)))
    seen = np.zeros(n, dtype=int)
    for rank in range(world):
        glob, r = ret[rank]
        seen[glob] += 1
        assert np.abs(r - r_all[glob]).max() <= 1e-12 * np.abs(r_all).max()
    assert (seen == 1).all()


def test_halo_plan_from_the_references_distribute_plan(a2ds):
    """a2ds_halo_from_distribute: TACSBVecDistribute's slab lists of GLOBAL node numbers
    (src/bpmat/TACSBVecDistribute.h:156-172), emulated here from an element partition the way
    the reference builds them (ext_vars = sorted ghost globals grouped by owner, req_vars = what
    each other rank's ext list asks of this rank; owners hold contiguous global ranges), must
    give the same peers / send / receive lists as the native planner — for a 4-rank scattered
    partition of an unstructured mesh and for row slabs."""
    for conn, X, er, nr in _distribute_cases(a2ds):
        parts = a2ds.meshes.partition_rows(conn, len(X), er)
        for q in parts:
            q["n_owned"] = len(q["owned"])
        # reference numbering: rank r owns the contiguous global range [lo_r, lo_r + n_owned_r)
        lo = np.concatenate([[0], np.cumsum([p["n_owned"] for p in parts])])
        new_of = np.full(len(X), -1, dtype=np.int64)
        for r, p in enumerate(parts):
            new_of[p["glob"][:p["n_owned"]]] = lo[r] + np.arange(p["n_owned"])
        owner = np.searchsorted(lo, np.arange(lo[-1]), side="right") - 1
        ext = []      # per rank: sorted global (new numbering) ghost ids
        for r, p in enumerate(parts):
            ext.append(np.sort(new_of[p["glob"][p["n_owned"]:]]))
        for r, p in enumerate(parts):
            ev = ext[r]
            eo = owner[ev]
            ext_proc = np.unique(eo)
            ext_ptr = np.array([np.searchsorted(eo, q) for q in ext_proc])
            ext_count = np.array([np.sum(eo == q) for q in ext_proc])
            req_proc, req_vars = [], []
            for q in range(nr):
                if q == r:
                    continue
                mine = ext[q][owner[ext[q]] == r]
                if len(mine):
                    req_proc.append(q); req_vars.append(mine)
            req_count = np.array([len(v) for v in req_vars], dtype=np.int64)
            req_ptr = np.concatenate([[0], np.cumsum(req_count)])[:-1] if len(req_vars) else np.zeros(0)
            peers, sends, recvs = a2ds.halo_from_distribute(
                int(lo[r]), p["n_owned"], ext_proc, ext_ptr, ext_count, req_proc, req_ptr, req_count,
                np.concatenate(req_vars) if req_vars else np.zeros(0))
            # the same exchange in the reference's numbering: map our local ids back to globals
            # of the ORIGINAL mesh and compare as sets per peer with the native plan
            glob_new = np.concatenate([lo[r] + np.arange(p["n_owned"]), ev])   # local -> new global
            old_of_new = np.zeros(lo[-1], dtype=np.int64); old_of_new[new_of[new_of >= 0]] = np.nonzero(new_of >= 0)[0]
            native = {int(q): (set(p["glob"][s].tolist()), set(p["glob"][v].tolist()))
                      for q, s, v in zip(p["peers"], p["send_lists"], p["recv_lists"])}
            got = {int(q): (set(old_of_new[glob_new[s]].tolist()), set(old_of_new[glob_new[v]].tolist()))
                   for q, s, v in zip(peers, sends, recvs)}
            got = {q: sv for q, sv in got.items() if sv[0] or sv[1]}
            native = {q: sv for q, sv in native.items() if sv[0] or sv[1]}
            assert got == native
            for v in recvs:
                assert np.all(v >= p["n_owned"])
            for s in sends:
                assert np.all((s >= 0) & (s < p["n_owned"]))


def _distribute_cases(a2ds):
    conn, X, _ = a2ds.meshes.plate(9, 8)
    yield conn, X, (np.arange(len(conn)) // 9 // 2).astype(np.int32), 4
    conn, X, _ = a2ds.meshes.cubed_sphere(4, shuffle_seed=3)
    rng = np.random.default_rng(5)
    yield conn, X, rng.integers(0, 4, len(conn)).astype(np.int32), 4
