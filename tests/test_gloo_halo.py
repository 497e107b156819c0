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
