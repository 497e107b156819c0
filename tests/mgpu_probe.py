"""Launched under torchrun by test_multi_gpu.py: N ranks, one GPU each, plate slabs with NCCL
halo exchange; rank 0 also assembles the whole plate on its own GPU and compares."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
a2ds = importlib.import_module("a2d-shells_b200")


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    nx, ny = 24, 10
    slab = a2ds.meshes.plate_slab(rank, world, nx, ny, bump=2e-2)
    Cs, eth = a2ds.iso_shell_tables()
    asm = a2ds.Assembler(local)
    asm.set_mesh(slab["conn"], slab["n_nodes"], slab["n_owned"])
    asm.set_nodes(slab["X"])
    asm.set_components(Cs[None], eth[None])
    asm.set_bcs(slab["bc_nodes"], 63)
    uid = [asm.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    asm.comm_init(world, rank, uid[0])
    asm.set_halo(slab["peers"], slab["send_lists"], slab["recv_lists"])
    # owned part of the state only: ghosts must arrive through the halo
    u_owned = a2ds.meshes.seeded_state(slab["glob"][:slab["n_owned"]], 1e-5)
    asm.set_state(u_owned)
    asm.halo_forward()
    k = asm.create_mat(); g = asm.create_mat()
    res = asm.assembleAll(k, g)
    rowp, cols = asm.mat_pattern(k)
    K = asm.mat_values(k); G = asm.mat_values(g)
    # distributed mat-vec y = K x: x known on the owned nodes only, halo inside the call
    nl, no = slab["n_nodes"], slab["n_owned"]
    x = torch.zeros((nl, 6), dtype=torch.float64, device="cuda")
    x[:no] = torch.from_numpy(a2ds.meshes.seeded_state(slab["glob"][:no] + 31337, 1.0)).cuda()
    x[no:] = float("nan")        # ghosts must come through the halo
    y = torch.zeros_like(x)
    asm.mat_mult_dist_dev(k, x.data_ptr(), y.data_ptr())
    asm.synchronize()
    out = dict(glob=slab["glob"], n_owned=slab["n_owned"], res=res, rowp=rowp, cols=cols, K=K, G=G,
               y=y[:no].cpu().numpy())
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(out, gathered, dst=0)
    asm.close()
    if rank == 0:
        conn, X, bcn = a2ds.meshes.plate(nx, ny * world, ly=1.0 * world, bump=2e-2)
        # same geometry as the slabs: plate_slab uses a per-slab bump period of ly
        Xs = np.zeros_like(X)
        for o in gathered:
            pass
        n = len(X)
        ref = a2ds.Assembler(local)
        ref.set_mesh(conn, n)
        # rebuild X exactly as the slabs see it
        for r in range(world):
            s = a2ds.meshes.plate_slab(r, world, nx, ny, bump=2e-2)
            Xs[s["glob"]] = s["X"]
        ref.set_nodes(Xs); ref.set_components(Cs[None], eth[None]); ref.set_bcs(bcn, 63)
        ref.set_state(a2ds.meshes.seeded_state(np.arange(n), 1e-5))
        kk = ref.create_mat(); gg = ref.create_mat()
        r_all = ref.assembleAll(kk, gg)
        rp, cl = ref.mat_pattern(kk)
        K_all = ref.mat_values(kk); G_all = ref.mat_values(gg)
        ref.close()
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import bcsr_matvec
        y_all = bcsr_matvec(K_all, rp, cl, a2ds.meshes.seeded_state(np.arange(n) + 31337, 1.0))
        worst_y = max(np.abs(o["y"] - y_all[o["glob"][:o["n_owned"]]]).max() for o in gathered)
        worst_y /= np.abs(y_all).max()
        print("MGPU_MATVEC", worst_y)
        assert worst_y < 1e-13, worst_y
        worst = [0.0, 0.0, 0.0]
        shared_rows = set()
        for r in range(1, world):
            shared_rows.update((r * ny * (nx + 1) + np.arange(nx + 1)).tolist())
        for o in gathered:
            glob, no = o["glob"], o["n_owned"]
            worst[0] = max(worst[0], np.abs(o["res"] - r_all[glob[:no]]).max() / np.abs(r_all).max())
            for lr in range(no):
                gr = int(glob[lr])
                if gr in shared_rows:
                    continue  # interface rows stay unassembled per rank (TACSSchurMat convention)
                for kb in range(o["rowp"][lr], o["rowp"][lr + 1]):
                    gc = int(glob[o["cols"][kb]])
                    j = rp[gr] + int(np.searchsorted(cl[rp[gr]:rp[gr + 1]], gc))
                    assert cl[j] == gc
                    worst[1] = max(worst[1], np.abs(o["K"][kb] - K_all[j]).max() / np.abs(K_all).max())
                    worst[2] = max(worst[2], np.abs(o["G"][kb] - G_all[j]).max() / np.abs(G_all).max())
        print("MGPU_RESULT", world, *worst)
        assert worst[0] < 1e-12 and worst[1] < 1e-10 and worst[2] < 1e-10, worst
        print("MGPU_OK")
    parallel_mat_flavour(rank, world, local)
    native_partition(rank, world, local)
    dist.destroy_process_group()


def native_partition(rank, world, local):
    """a2ds_partition_build / a2ds_partition_apply (host C++): every rank derives its sub-mesh
    and halo plan from the global mesh on its own; unstructured mesh, scattered element ->
    rank map (every rank talks to every other); the residual of the owned nodes (ghost
    contributions reverse-added over NCCL) must equal the one-GPU residual"""
    conn, X = a2ds.meshes.cubed_sphere(6, shuffle_seed=11)[:2]
    n = len(X)
    elem_rank = ((np.arange(len(conn)) * 7919) % 97) % world
    elem_comp = (np.arange(len(conn)) % 3).astype(np.int32)
    Cs, eth = a2ds.iso_shell_tables()
    Cs3 = np.stack([Cs * (1.0 + 0.5 * k) for k in range(3)]); eth3 = np.stack([eth] * 3)
    P = a2ds.Partition(conn, n, elem_rank, world, rank)
    asm = a2ds.Assembler(local)
    uid = [asm.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    asm.comm_init(world, rank, uid[0])
    P.apply(asm, elem_comp)
    asm.set_nodes(X[P.glob]); asm.set_components(Cs3, eth3)
    asm.set_state(a2ds.meshes.seeded_state(P.glob[:P.n_owned], 1e-5)); asm.halo_forward()
    res = asm.assembleRes()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(dict(glob=P.glob, no=P.n_owned, res=res), gathered, dst=0)
    asm.close()
    if rank != 0:
        return
    ref = a2ds.Assembler(local)
    ref.set_mesh(conn, n, elem_comp=elem_comp); ref.set_nodes(X); ref.set_components(Cs3, eth3)
    ref.set_state(a2ds.meshes.seeded_state(np.arange(n), 1e-5))
    r_all = ref.assembleRes()
    ref.close()
    seen = np.zeros(n, dtype=int)
    worst = 0.0
    for o in gathered:
        own = o["glob"][:o["no"]]
        seen[own] += 1
        worst = max(worst, np.abs(o["res"] - r_all[own]).max() / np.abs(r_all).max())
    print("MGPU_NATIVE_PARTITION", world, worst)
    assert np.all(seen[np.unique(conn)] == 1) and worst < 1e-12, worst
    print("MGPU_NATIVE_PARTITION_OK")


def parallel_mat_flavour(rank, world, local):
    """TACSParallelMat flavour: generic element partition (column strips, so every rank has
    two neighbours at world > 2), matrix halo -> the OWNED rows of every rank must equal the
    rows of the matrix assembled on one GPU, for K (Jacobian), G and M"""
    nx, ny = 4 * world, 9
    conn, X, bcn = a2ds.meshes.plate(nx, ny, bump=3e-2)
    n = len(X)
    elem_rank = (np.arange(nx * ny) % nx) // 4
    P = a2ds.meshes.partition_rows(conn, n, elem_rank, matrix_halo=True)[rank]
    Cs, eth = a2ds.iso_shell_tables(t_offset=0.1)
    mom = a2ds.iso_mass_moments(2718.0, 0.01, 0.1)
    glob = P["glob"]; no = len(P["owned"]); nl = len(glob)
    asm = a2ds.Assembler(local)
    asm.set_mesh(P["conn_local"], nl, no); asm.set_nodes(X[glob])
    asm.set_components(Cs[None], eth[None]); asm.set_mass_moments(mom[None])
    bc_local = P["local_of"][bcn]; bc_local = bc_local[bc_local >= 0].astype(np.int32)
    asm.set_bcs(bc_local, 63)
    uid = [asm.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    asm.comm_init(world, rank, uid[0])
    asm.set_halo(P["peers"], P["send_lists"], P["recv_lists"])
    # the native planner (a2ds_partition_build_matrix) must give the very same plan; the first
    # matrix is created from it, the others from the Python plan
    PN = a2ds.Partition(conn, n, elem_rank, world, rank, matrix_halo=True)
    assert np.array_equal(PN.glob, glob) and np.array_equal(PN.rowp, P["rowp"]) and np.array_equal(PN.cols, P["cols"])
    for a, b in zip(PN.mat_send_lists + PN.mat_recv_lists, P["mat_send_lists"] + P["mat_recv_lists"]):
        assert np.array_equal(a, b)
    ident = np.arange(nl, dtype=np.int32)
    mats = [PN.create_mat(asm)]
    for _ in range(2):
        m = asm.create_mat_from_pattern([dict(nrows=nl, rowp=P["rowp"], cols=P["cols"], row_map=ident,
                                              col_map=ident, ident=1)])
        asm.mat_set_halo(m, P["peers"], P["mat_send_lists"], P["mat_recv_lists"])
        mats.append(m)
    asm.set_state(a2ds.meshes.seeded_state(glob[:no], 1e-5)); asm.halo_forward()
    res = asm.assembleJacobian(1.0, 0.0, 0.0, mats[0])
    asm.assembleMatType(a2ds.GEOMETRIC_STIFFNESS_MATRIX, mats[1])
    asm.assembleMatType(a2ds.MASS_MATRIX, mats[2])
    vals = [asm.mat_values(m) for m in mats]
    # ParallelMat::mult with the assembled owned rows: only the forward exchange of x
    x = torch.full((nl, 6), float("nan"), dtype=torch.float64, device="cuda")
    x[:no] = torch.from_numpy(a2ds.meshes.seeded_state(glob[:no] + 4242, 1.0)).cuda()
    asm.set_state_dev(nl, x.data_ptr()); asm.halo_forward()       # reuse the state halo for x
    y = torch.zeros_like(x)
    asm.mat_mult_dev(mats[0], asm.state_dev(), y.data_ptr())
    asm.synchronize()
    out = dict(glob=glob, no=no, rowp=P["rowp"], cols=P["cols"], res=res, vals=vals,
               y=y[:no].cpu().numpy())
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(out, gathered, dst=0)
    asm.close()
    if rank != 0:
        return
    ref = a2ds.Assembler(local)
    ref.set_mesh(conn, n); ref.set_nodes(X); ref.set_components(Cs[None], eth[None])
    ref.set_mass_moments(mom[None]); ref.set_bcs(bcn, 63)
    ref.set_state(a2ds.meshes.seeded_state(np.arange(n), 1e-5))
    k = ref.create_mat(); g = ref.create_mat(); mm = ref.create_mat()
    r_all = ref.assembleJacobian(1.0, 0.0, 0.0, k)
    ref.assembleMatType(a2ds.GEOMETRIC_STIFFNESS_MATRIX, g)
    ref.assembleMatType(a2ds.MASS_MATRIX, mm)
    rp, cl = ref.mat_pattern(k)
    full = [ref.mat_values(m) for m in (k, g, mm)]
    ref.close()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import bcsr_matvec
    y_all = bcsr_matvec(full[0], rp, cl, a2ds.meshes.seeded_state(np.arange(n) + 4242, 1.0))
    worst = [0.0, 0.0, 0.0, 0.0, 0.0]
    rows = 0
    for o in gathered:
        gl, no = o["glob"], o["no"]
        worst[3] = max(worst[3], np.abs(o["res"] - r_all[gl[:no]]).max() / np.abs(r_all).max())
        worst[4] = max(worst[4], np.abs(o["y"] - y_all[gl[:no]]).max() / np.abs(y_all).max())
        for lr in range(no):
            gr = int(gl[lr]); rows += 1
            assert o["rowp"][lr + 1] - o["rowp"][lr] == rp[gr + 1] - rp[gr]
            for kb in range(o["rowp"][lr], o["rowp"][lr + 1]):
                gc = int(gl[o["cols"][kb]])
                j = rp[gr] + int(np.searchsorted(cl[rp[gr]:rp[gr + 1]], gc))
                assert cl[j] == gc
                for t in range(3):
                    worst[t] = max(worst[t], np.abs(o["vals"][t][kb] - full[t][j]).max() /
                                   np.abs(full[t]).max())
    print("MGPU_PARALLELMAT", world, rows, *worst)
    assert rows == n and max(worst[:3]) < 1e-10 and worst[3] < 1e-12 and worst[4] < 1e-12, worst
    print("MGPU_PARALLELMAT_OK")


if __name__ == "__main__":
    main()
