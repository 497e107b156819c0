"""Host-side logic and the C-ABI surface.  CPU only: no compute calls."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "a2ds.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(a2ds_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(a2ds):
    L = a2ds.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(a2ds.capi.SYMBOLS) == declared
    assert b"sm_100a" in L.a2ds_version()


def test_no_gpu_means_loud_failure(a2ds):
    from conftest import has_gpu
    if has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(a2ds.A2dsError, match="no CUDA device"):
        a2ds.Assembler(0)


def test_product_does_not_reference_the_oracle():
    """the product tree must never import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "a2d-shells_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), (dirpath, f)
    hdr = open(os.path.join(ROOT, "include", "a2ds.h")).read()
    assert "oracle" not in hdr.lower()


def test_host_pattern_matches_oracle_pattern(a2ds, orc):
    for conn, X, _ in (a2ds.meshes.plate(9, 6), a2ds.meshes.cylinder(10, 5)):
        rp, cl = a2ds.host_pattern(len(X), conn)
        rpo, clo = orc.pattern(len(X), conn)
        assert rp.tobytes() == rpo.tobytes() and cl.tobytes() == clo.tobytes()
    rp, cl = a2ds.host_pattern(4, np.zeros((0, 4), dtype=np.int32))
    assert rp.tolist() == [0] * 5 and len(cl) == 0


def test_host_pattern_threaded_path(a2ds, orc):
    """above 65 536 nodes the sort / unique sweep of the pattern runs on several host threads"""
    for conn, X in (a2ds.meshes.plate(300, 280)[:2], a2ds.meshes.cubed_sphere(120, shuffle_seed=5)[:2]):
        assert len(X) > (1 << 16)
        rp, cl = a2ds.host_pattern(len(X), conn)
        rpo, clo = orc.pattern(len(X), conn)
        assert rp.tobytes() == rpo.tobytes() and cl.tobytes() == clo.tobytes()


def test_host_pattern_matches_golden(a2ds):
    for name in ("plate", "cylinder"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        rp, cl = a2ds.host_pattern(len(g["X"]), g["conn"])
        assert rp.tobytes() == g["rowp"].tobytes() and cl.tobytes() == g["cols"].tobytes()


def test_coloring_is_valid(a2ds):
    for conn, X, _ in (a2ds.meshes.plate(11, 7), a2ds.meshes.cylinder(9, 4),
                       a2ds.meshes.cylinder(8, 3)):
        color, nc = a2ds.host_color_elements(len(X), conn)
        assert color.min() == 0 and color.max() == nc - 1 and nc <= 9
        for c in range(nc):
            nodes = conn[color == c].ravel()
            assert len(nodes) == len(np.unique(nodes))  # no node shared inside a colour


def test_structured_plate_needs_four_colors(a2ds):
    conn, X, _ = a2ds.meshes.plate(20, 20)
    _, nc = a2ds.host_color_elements(len(X), conn)
    assert nc == 4


def _fan(m):
    """m quads around one hub node (a valence-m junction): every pair of them shares the hub"""
    conn = np.array([[0, 1 + 2 * k, 2 + 2 * k, 1 + (2 * k + 2) % (2 * m)] for k in range(m)], dtype=np.int32)
    return conn, 2 * m + 1


def test_hashed_coloring_is_valid(a2ds):
    """the device's colouring rule stepped on the host: valid on structured and unstructured
    meshes, and on a fan that needs more colours than a machine word holds"""
    cases = [(c, len(X)) for c, X, _ in (a2ds.meshes.plate(11, 7), a2ds.meshes.cylinder(9, 4),
                                         a2ds.meshes.cubed_sphere(6, shuffle_seed=3))]
    cases.append(_fan(70))
    for conn, n in cases:
        color, nc = a2ds.host_color_elements_hashed(n, conn)
        assert color.min() == 0 and color.max() == nc - 1
        for c in range(nc):
            nodes = conn[color == c].ravel()
            assert len(nodes) == len(np.unique(nodes))
    assert nc == 70   # the fan: all elements pairwise adjacent
    conn, X, _ = a2ds.meshes.plate(40, 40)
    assert a2ds.host_color_elements_hashed(len(X), conn)[1] <= 9


def test_seeded_state_is_partition_independent(a2ds):
    ids = np.arange(1000)
    u = a2ds.meshes.seeded_state(ids, 1e-5)
    perm = np.random.default_rng(0).permutation(1000)
    assert np.array_equal(a2ds.meshes.seeded_state(ids[perm], 1e-5), u[perm])
    assert np.abs(u).max() <= 1e-5 and abs(u.mean()) < 1e-6
    assert not np.array_equal(u, a2ds.meshes.seeded_state(ids, 1e-5, seed=1))


def test_partition_ownership_follows_first_touch(a2ds):
    conn, X, _ = a2ds.meshes.cylinder(12, 8)
    ne = len(conn)
    elem_rank = (np.arange(ne) * 4) // ne
    parts = a2ds.meshes.partition_rows(conn, len(X), elem_rank)
    owned_all = np.concatenate([p["owned"] for p in parts])
    assert len(owned_all) == len(X) and len(np.unique(owned_all)) == len(X)
    for p in parts:
        glob = p["glob"]
        assert np.array_equal(glob[p["conn_local"]], conn[p["elems"]])
        # halo plans pair up: what r sends to q is what q receives from r
        for k, q in enumerate(p["peers"]):
            other = parts[q]
            kk = list(other["peers"]).index(p["rank"])
            assert np.array_equal(glob[p["send_lists"][k]], other["glob"][other["recv_lists"][kk]])
            assert np.all(p["send_lists"][k] < len(p["owned"]))
            assert np.all(p["recv_lists"][k] >= len(p["owned"]))


def _build_probe():
    import subprocess
    exe = os.path.join(ROOT, "tests", "_device_assembler_probe")
    src = os.path.join(ROOT, "tests", "device_assembler_probe.cpp")
    lib = os.path.join(ROOT, "a2d-shells_b200", "lib")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-std=c++11", "-O1", "-I" + os.path.join(ROOT, "include"), src,
                               "-o", exe, "-L" + lib, "-la2ds_b200", "-Wl,-rpath," + lib])
    return exe


def test_cpp_sidecar_compiles_and_fails_loudly_without_gpu(a2ds):
    """a2ds::DeviceAssembler (host/DeviceAssembler.h): header-only C++ mirror of the three
    TACSAssembler entry points over the C ABI"""
    import subprocess
    from conftest import has_gpu
    exe = _build_probe()
    out = subprocess.run([exe], capture_output=True, text=True)
    if has_gpu():
        assert "DEVICE_ASSEMBLER_OK" in out.stdout, out.stdout + out.stderr
    else:
        assert out.returncode == 1 and "no CUDA device" in out.stdout


def test_structured_slabs_match_generic_partition(a2ds):
    """plate_slab / cylinder_slab (built per rank, no global mesh) against partition_rows"""
    nx, ny, N = 5, 3, 3
    conn, X, _ = a2ds.meshes.plate(nx, ny * N, ly=float(N))
    parts = a2ds.meshes.partition_rows(conn, len(X), np.arange(len(conn)) // (nx * ny))
    for r in range(N):
        s, p = a2ds.meshes.plate_slab(r, N, nx, ny), parts[r]
        assert np.array_equal(s["glob"], p["glob"]) and np.array_equal(s["conn"], p["conn_local"])
        assert s["n_owned"] == len(p["owned"]) and list(s["peers"]) == list(p["peers"])
        for a, b in zip(s["send_lists"] + s["recv_lists"], p["send_lists"] + p["recv_lists"]):
            assert np.array_equal(a, b)
    nt, nxp = 6, 2
    conn, X, _ = a2ds.meshes.cylinder(nt, nxp * N)
    parts = a2ds.meshes.partition_rows(conn, len(X), np.arange(len(conn)) // (nt * nxp))
    for r in range(N):
        s, p = a2ds.meshes.cylinder_slab(r, N, nt, nxp), parts[r]
        assert np.array_equal(s["glob"], p["glob"]) and np.array_equal(s["conn"], p["conn_local"])
        assert np.allclose(s["X"], X[s["glob"]])
        for a, b in zip(s["send_lists"] + s["recv_lists"], p["send_lists"] + p["recv_lists"]):
            assert np.array_equal(a, b)


def test_pattern_and_colouring_on_unstructured_mesh(a2ds, orc):
    """host set-up on a mesh with irregular valence and shuffled numbering: pattern
    bit-identical to the oracle's, colouring valid (no two elements of a colour share a node)"""
    for seed in (None, 3):
        conn, X, _ = a2ds.meshes.cubed_sphere(6, shuffle_seed=seed)
        n = len(X)
        assert n == 6 * 36 + 2 and np.bincount(np.bincount(conn.ravel()))[3] == 8
        rowp, cols = a2ds.host_pattern(n, conn)
        rp, cl = orc.pattern(n, conn)
        assert rowp.tobytes() == rp.tobytes() and cols.tobytes() == cl.tobytes()
        color, nc = a2ds.host_color_elements(n, conn)
        assert color.min() == 0 and color.max() == nc - 1
        for c in range(nc):
            nodes = conn[color == c].ravel()
            assert len(np.unique(nodes)) == len(nodes)


def test_matrix_halo_plan_assembles_owned_rows(a2ds, orc):
    """ParallelMat flavour on a 2 x 2 partition (the centre node is shared by four ranks):
    per-rank matrices from the oracle on the local meshes and patterns of
    partition_rows(matrix_halo=True), the block exchange of a2ds_mat_set_halo emulated in
    numpy -> every rank's owned rows equal the rows of the globally assembled matrix."""
    nx = 6
    conn, X, _ = a2ds.meshes.plate(nx, nx, bump=4e-2)
    n = len(X)
    ei = np.arange(nx * nx) % nx; ej = np.arange(nx * nx) // nx
    elem_rank = (ei >= nx // 2).astype(int) + 2 * (ej >= nx // 2).astype(int)
    parts = a2ds.meshes.partition_rows(conn, n, elem_rank, matrix_halo=True)
    Cs, eth = a2ds.iso_shell_tables()
    comp = orc.make_comp(0, Cs, eth)
    u = a2ds.meshes.seeded_state(np.arange(n), 1e-5)
    rowp_g, cols_g = orc.pattern(n, conn)
    _, K_g = orc.assemble(2, conn, np.zeros(len(conn), dtype=np.int32), [comp], X, u, rowp_g, cols_g)
    K_loc = []
    for P in parts:
        _, K = orc.assemble(2, P["conn_local"], np.zeros(len(P["conn_local"]), dtype=np.int32),
                            [comp], X[P["glob"]], u[P["glob"]], P["rowp"], P["cols"])
        K_loc.append(K)
    # the exchange: sender packs its listed blocks, receiver adds them at its listed blocks
    recv = [K.copy() for K in K_loc]
    for P in parts:
        for ip, p in enumerate(P["peers"]):
            Q = parts[int(p)]
            iq = list(Q["peers"]).index(P["rank"])
            src = K_loc[P["rank"]][P["mat_send_lists"][ip]]
            dst = Q["mat_recv_lists"][iq]
            assert len(src) == len(dst)
            np.add.at(recv[int(p)], dst, src)
    seen = 0
    for P, K in zip(parts, recv):
        for l, g in enumerate(P["owned"]):
            row_l = {int(P["glob"][c]): K[k] for k, c in
                     zip(range(P["rowp"][l], P["rowp"][l + 1]), P["cols"][P["rowp"][l]:P["rowp"][l + 1]])}
            for k in range(rowp_g[g], rowp_g[g + 1]):
                assert np.abs(row_l[int(cols_g[k])] - K_g[k]).max() <= 1e-13 * np.abs(K_g).max()
            assert len(row_l) == rowp_g[g + 1] - rowp_g[g]
            seen += 1
    assert seen == n


def test_native_partition_matches_reference_ownership_rule(a2ds):
    """a2ds_partition_build (host C++): one rank's sub-mesh and halo plan from the global mesh;
    against meshes.partition_rows (first-touch ownership of TACSCreator.cpp:1156-1205, checked
    above) on structured and unstructured meshes, slab and random element -> rank maps"""
    rng = np.random.default_rng(5)
    meshes = dict(plate=a2ds.meshes.plate(9, 7)[:2], cyl=a2ds.meshes.cylinder(10, 6)[:2],
                  sphere=a2ds.meshes.cubed_sphere(4, shuffle_seed=3)[:2])
    for name, (conn, X) in meshes.items():
        for N in (1, 2, 3, 5):
            for mode in ("slab", "random"):
                er = (np.arange(len(conn)) * N // len(conn)) if mode == "slab" \
                    else rng.integers(0, N, len(conn))
                if er.max() + 1 < N:
                    continue
                parts = a2ds.meshes.partition_rows(conn, len(X), er)
                for r in range(N):
                    P, Q = a2ds.Partition(conn, len(X), er, N, r), parts[r]
                    assert np.array_equal(P.glob, Q["glob"]) and P.n_owned == len(Q["owned"])
                    assert np.array_equal(P.elems, Q["elems"])
                    assert np.array_equal(P.conn_local, Q["conn_local"])
                    assert np.array_equal(P.ghost_owner, Q["ghost_owner"])
                    assert np.array_equal(P.peers, Q["peers"])
                    assert len(P.send_lists) == len(Q["send_lists"])
                    for a, b in zip(P.send_lists + P.recv_lists, Q["send_lists"] + Q["recv_lists"]):
                        assert np.array_equal(a, b)
    # the two sides of every pair name the same global nodes in the same order
    conn, X = meshes["sphere"]
    er = rng.integers(0, 4, len(conn))
    P = [a2ds.Partition(conn, len(X), er, 4, r) for r in range(4)]
    owners = np.full(len(X), -1)
    for a in P:
        assert np.all(owners[a.glob[:a.n_owned]] == -1)       # every node has one owner
        owners[a.glob[:a.n_owned]] = 1
        for k, q in enumerate(a.peers):
            b = P[int(q)]
            kb = list(b.peers).index(P.index(a))
            assert np.array_equal(a.glob[a.send_lists[k]], b.glob[b.recv_lists[kb]])
    assert np.all(owners[np.unique(conn)] == 1)
    # a rank without elements, and bad input
    er0 = np.zeros(len(conn), dtype=np.int32)
    E = a2ds.Partition(conn, len(X), er0, 2, 1)
    assert E.n_nodes == 0 and len(E.elems) == 0 and len(E.peers) == 0
    with pytest.raises(a2ds.A2dsError, match="elem_rank"):
        a2ds.Partition(conn, len(X), er0 + 7, 2, 0)
    with pytest.raises(a2ds.A2dsError, match="outside"):
        a2ds.Partition(conn + len(X), len(X), er0, 1, 0)


def test_native_matrix_halo_plan_matches_python_planner(a2ds):
    """a2ds_partition_build_matrix (TACSParallelMat flavour): extended ghost set, local pattern
    and per-peer block lists identical to meshes.partition_rows(matrix_halo=True), whose plan
    the 2- and 4-GPU tests assemble with"""
    rng = np.random.default_rng(1)
    meshes = dict(plate=a2ds.meshes.plate(8, 9)[:2], cyl=a2ds.meshes.cylinder(8, 5)[:2],
                  sphere=a2ds.meshes.cubed_sphere(3, shuffle_seed=2)[:2],
                  wing=a2ds.meshes.wingbox(3, 2, 3, 2)[:2])
    for name, (conn, X) in meshes.items():
        n = len(X)
        for N in (2, 3, 4):
            for mode in ("slab", "strip", "random"):
                er = {"slab": np.arange(len(conn)) * N // len(conn), "strip": np.arange(len(conn)) % N,
                      "random": rng.integers(0, N, len(conn))}[mode]
                if er.max() + 1 < N:
                    continue
                parts = a2ds.meshes.partition_rows(conn, n, er, matrix_halo=True)
                for r in range(N):
                    P, Q = a2ds.Partition(conn, n, er, N, r, matrix_halo=True), parts[r]
                    assert np.array_equal(P.glob, Q["glob"]) and P.n_owned == len(Q["owned"])
                    assert np.array_equal(P.conn_local, Q["conn_local"])
                    assert np.array_equal(P.peers, Q["peers"])
                    assert np.array_equal(P.rowp, Q["rowp"]) and np.array_equal(P.cols, Q["cols"])
                    pairs = zip(P.send_lists + P.recv_lists + P.mat_send_lists + P.mat_recv_lists,
                                Q["send_lists"] + Q["recv_lists"] + Q["mat_send_lists"] + Q["mat_recv_lists"])
                    for a, b in pairs:
                        assert np.array_equal(a, b), (name, N, mode, r)
    # the vector-only partition does not carry a matrix plan
    conn, X = meshes["plate"]
    P = a2ds.Partition(conn, len(X), np.zeros(len(conn), dtype=np.int32), 1, 0)
    assert not hasattr(P, "rowp")
