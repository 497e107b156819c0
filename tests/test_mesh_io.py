"""Mesh input: NASTRAN bulk-data reader and binary container (a2ds_mesh_*, host only).

Parity is against the reference's own loader (TACSMeshLoader::scanBDFFile,
src/io/TACSMeshLoader.cpp:570): array for array, bit-exact, live where oracle/_ref is built
and against tests/golden/bdf.npz (generated from it by tests/golden/make_golden.py bdf)
everywhere.  One GPU test at the end runs the shipped-example flow deck -> assembly."""
import importlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
DECKS = ["mixed", "cyl_large", "cyl_small", "cyl_free"]
INT_KEYS = ("elem_ptr", "elem_conn", "elem_comp", "bc_nodes", "bc_ptr", "bc_vars")


def _same(m, ref, prefix=""):
    g = (lambda k: ref[prefix + k])
    for k in INT_KEYS:
        assert np.array_equal(getattr(m, k), g(k)), k
    # coordinates and prescribed values: the same decimal strings through strtod -> identical bits
    assert m.X.tobytes() == np.ascontiguousarray(g("X"), dtype=np.float64).tobytes()
    assert m.bc_vals.tobytes() == np.ascontiguousarray(g("bc_vals"), dtype=np.float64).tobytes()
    assert m.elem_descript == [str(s) for s in g("elem_descript")]
    assert m.comp_descript == [str(s) for s in g("comp_descript")]


@pytest.mark.parametrize("deck", DECKS)
def test_bdf_reader_matches_golden(a2ds, deck):
    gold = np.load(os.path.join(GOLD, "bdf.npz"))
    m = a2ds.Mesh.read_bdf(os.path.join(GOLD, deck + ".bdf"))
    _same(m, gold, deck + "_")
    assert m.n_comp == int(gold[deck + "_n_comp"])


@pytest.mark.parametrize("deck", DECKS)
def test_bdf_reader_matches_reference_loader(a2ds, ref, deck, capfd):
    path = os.path.join(GOLD, deck + ".bdf")
    fail, r, _ = ref.bdf_scan(path)
    assert fail == 0
    _same(a2ds.Mesh.read_bdf(path), r)
    if deck == "mixed":   # unknown cards are reported, not fatal — by both readers
        err = capfd.readouterr().err
        assert err.count("PSHELL") >= 2 and "a2ds_mesh_read_bdf: card not recognized" in err


def test_shipped_example_decks(a2ds, ref):
    """examples/cylinder-buckling/{mech,therm}-cylinder.bdf (BASELINE configs[0]): live against
    the reference loader in the build container, by digest elsewhere (the decks do not travel)"""
    ex = "/root/reference/examples/cylinder-buckling"
    if not os.path.isdir(ex):
        pytest.skip("reference examples not present")
    gold = np.load(os.path.join(GOLD, "bdf.npz"))
    for name in ("mech-cylinder", "therm-cylinder"):
        path = os.path.join(ex, name + ".bdf")
        fail, r, _ = ref.bdf_scan(path)
        assert fail == 0
        m = a2ds.Mesh.read_bdf(path)
        _same(m, r)
        assert [m.n_nodes, m.n_elems, m.n_bcs] == gold[name + "_sizes"].tolist() == [3300, 3200, 600]
        digest = [float(m.X.sum()), float(np.abs(m.X).sum()),
                  float((m.elem_conn.astype(np.int64) * (1 + np.arange(len(m.elem_conn)) % 7)).sum()),
                  float((m.bc_nodes.astype(np.int64) * (1 + m.bc_vars)).sum()),
                  float(m.bc_vals.sum())]
        assert digest == gold[name + "_digest"].tolist()
        conn, masks, vals = m.quad4()
        assert conn.shape == (3200, 4) and set(np.unique(masks)) == {1, 2, 4}
        # mech deck: end shortening -1e-5 on one ring; therm deck: all zero
        assert vals.min() == (-1e-5 if name.startswith("mech") else 0.0)


def test_writer_reader_round_trip_all_formats(a2ds, tmp_path):
    """meshes.write_bdf -> read_bdf: numbering by file number, tensor node order, BC cards"""
    conn, X, bcn = a2ds.meshes.plate(6, 4, bump=3e-2)
    rng = np.random.default_rng(3)
    nid = rng.permutation(len(X)) * 2 + 3
    eid = rng.permutation(len(conn)) + 100
    comp = rng.integers(0, 3, len(conn))
    dofs = ["123456" if i % 2 else "246" for i in range(len(bcn))]
    for fmt in ("large", "small", "free"):
        p = str(tmp_path / f"p_{fmt}.bdf")
        a2ds.meshes.write_bdf(p, conn, X, bcn, dofs, -2.5e-4, comp, fmt, nid, eid)
        m = a2ds.Mesh.read_bdf(p)
        nrank = np.argsort(np.argsort(nid))          # node k of ours -> position by file number
        order = np.argsort(eid)
        c4, masks, vals = m.quad4()
        assert np.array_equal(c4, nrank[conn][order])
        assert np.array_equal(m.elem_comp, comp[order])
        assert np.array_equal(m.node_nums, np.sort(nid) - 1) and np.array_equal(m.elem_nums, np.sort(eid) - 1)
        # significant digits the writer keeps: 10 (large, as the shipped decks), 4-5 (small), 11
        tol = {"large": 1e-9, "small": 2e-4, "free": 1e-10}[fmt]
        assert np.abs(m.X[nrank] - X).max() <= tol * max(1.0, np.abs(X).max())
        assert np.array_equal(m.bc_nodes, nrank[bcn])
        want = [sum(1 << (int(ch) - 1) for ch in d) for d in dofs]
        assert masks.tolist() == want
        assert np.allclose(vals[masks.astype(bool)][:, 1], -2.5e-4, rtol=1e-3)


def test_thread_count_does_not_change_the_result(a2ds, tmp_path):
    """a deck above the 1 MB threshold is cut into chunks parsed concurrently"""
    conn, X, bcn = a2ds.meshes.cylinder(120, 90)
    p = str(tmp_path / "big.bdf")
    a2ds.meshes.write_bdf(p, conn, X, bcn, ["12346"] * len(bcn), 0.0,
                          np.arange(len(conn)) % 5, "large", comp_names=["A", "B"])
    assert os.path.getsize(p) > (1 << 20)
    one = a2ds.Mesh.read_bdf(p, 1)
    for nt in (2, 3, 7, 16):
        many = a2ds.Mesh.read_bdf(p, nt)
        for k in INT_KEYS + ("node_nums", "elem_nums"):
            assert np.array_equal(getattr(one, k), getattr(many, k)), (nt, k)
        assert one.X.tobytes() == many.X.tobytes() and one.bc_vals.tobytes() == many.bc_vals.tobytes()
        assert one.comp_descript == many.comp_descript == ["A", "B", "", "", ""]
    c4, _, _ = one.quad4()
    assert np.array_equal(c4, conn)


def test_binary_container_round_trip(a2ds, tmp_path):
    for deck in ("mixed", "cyl_small"):
        m = a2ds.Mesh.read_bdf(os.path.join(GOLD, deck + ".bdf"))
        p = str(tmp_path / (deck + ".a2dm"))
        m.write_bin(p)
        b = a2ds.Mesh.read_bin(p)
        for k in INT_KEYS + ("node_nums", "elem_nums"):
            assert np.array_equal(getattr(m, k), getattr(b, k)), k
        assert m.X.tobytes() == b.X.tobytes() and m.bc_vals.tobytes() == b.bc_vals.tobytes()
        assert m.elem_descript == b.elem_descript and m.comp_descript == b.comp_descript
    # generated mesh -> container -> arrays
    conn, X, bcn = a2ds.meshes.plate(5, 3)
    g = a2ds.Mesh.from_arrays(conn, X, None, bcn, [[0, 1, 2]] * len(bcn), [1e-3] * len(bcn))
    p = str(tmp_path / "gen.a2dm")
    g.write_bin(p)
    b = a2ds.Mesh.read_bin(p)
    c4, masks, vals = b.quad4()
    assert np.array_equal(c4, conn) and masks.tolist() == [7] * len(bcn)
    assert np.array_equal(vals[:, :3], np.full((len(bcn), 3), 1e-3)) and not vals[:, 3:].any()
    assert b.X.tobytes() == np.ascontiguousarray(X).tobytes()


def test_empty_deck_and_deck_without_bulk_marker(a2ds, tmp_path):
    p = str(tmp_path / "empty.bdf")
    open(p, "w").write("SOL 103\nCEND\nBEGIN BULK\nENDDATA\n")
    m = a2ds.Mesh.read_bdf(p)
    assert (m.n_nodes, m.n_elems, m.n_bcs, m.n_comp) == (0, 0, 0, 0)
    assert m.elem_ptr.tolist() == [0] and m.bc_ptr.tolist() == [0]
    # no BEGIN BULK: the whole file is bulk data (src/io/TACSMeshLoader.cpp:752-757)
    p = str(tmp_path / "nobulk.bdf")
    open(p, "w").write("GRID           2              1.      2.      3.\n"
                       "GRID           1              0.      0.      0.\n")
    m = a2ds.Mesh.read_bdf(p)
    assert m.node_nums.tolist() == [0, 1] and m.X.tolist() == [[0, 0, 0], [1, 2, 3]]


def test_reader_failures_are_loud(a2ds, tmp_path):
    with pytest.raises(a2ds.A2dsError, match="unable to open"):
        a2ds.Mesh.read_bdf(str(tmp_path / "missing.bdf"))
    with pytest.raises(a2ds.A2dsError, match="unable to open"):
        a2ds.Mesh.read_bin(str(tmp_path / "missing.a2dm"))
    grid = "".join(f"GRID    {k:>8d}{'':8s}{float(k):8.1f}{0.0:8.1f}{0.0:8.1f}\n" for k in range(1, 5))
    cases = {
        # the reference stops at an empty line too (returns fail, :772-776)
        "empty line": grid + "\nCQUAD4         1       1       1       2       3       4\n",
        "not within limits": grid + "CQUAD4         1       1       1       2       3\nENDDATA\n",
        "undefined grid": grid + "CQUAD4         1       1       1       2       3       9\n",
        "positive element": grid + "CQUAD4         0       1       1       2       3       4\n",
        "second line": "GRID*                  1                              0.              0.*\n",
    }
    for what, text in cases.items():
        p = str(tmp_path / "bad.bdf")
        open(p, "w").write("BEGIN BULK\n" + text)
        with pytest.raises(a2ds.A2dsError, match=what):
            a2ds.Mesh.read_bdf(p)
    # a deck of triangles is not a 4-node shell mesh
    p = str(tmp_path / "tri.bdf")
    open(p, "w").write("BEGIN BULK\n" + grid + "CTRIA3         1       1       1       2       3\n")
    with pytest.raises(a2ds.A2dsError, match="does not have 4 nodes"):
        a2ds.Mesh.read_bdf(p).quad4()
    # corrupt containers
    p = str(tmp_path / "junk.a2dm")
    open(p, "wb").write(b"not a mesh at all, just bytes" * 4)
    with pytest.raises(a2ds.A2dsError, match="not a mesh container"):
        a2ds.Mesh.read_bin(p)
    good = str(tmp_path / "good.a2dm")
    a2ds.Mesh.read_bdf(os.path.join(GOLD, "cyl_small.bdf")).write_bin(good)
    raw = open(good, "rb").read()
    open(p, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(a2ds.A2dsError, match="truncated or inconsistent"):
        a2ds.Mesh.read_bin(p)
    # a header that claims gigabytes is refused before anything is allocated from it
    import struct
    open(p, "wb").write(b"A2DSMSH1" + struct.pack("<6q", 2**31 - 1, 2**31 - 1, 2**31 - 1, 0, 0, 1) + b"\0" * 64)
    with pytest.raises(a2ds.A2dsError, match="truncated or inconsistent"):
        a2ds.Mesh.read_bin(p)


def test_nastran_real_forms(a2ds, tmp_path):
    """E / D exponents and the compact form without a letter, in 8 and 16 columns"""
    forms = {"1.5E-3": 1.5e-3, "1.5D-3": 1.5e-3, "1.5-3": 1.5e-3, "-2.5+2": -250.0, ".25": 0.25,
             "7.": 7.0, "-1.-5": -1e-5, "1.5e+1": 15.0, "": 0.0}
    lines = ["BEGIN BULK"]
    for k, s in enumerate(forms):
        lines.append(f"GRID    {k + 1:>8d}{'':8s}{s:>8s}{s:<8s}{'0.':>8s}")
        lines.append(f"GRID*   {k + 101:>16d}{'':16s}{s:>16s}{s:<16s}*")
        lines.append(f"*       {s:>16s}")
    p = str(tmp_path / "reals.bdf")
    open(p, "w").write("\n".join(lines) + "\n")
    m = a2ds.Mesh.read_bdf(p)
    want = np.array(list(forms.values()))
    n = len(forms)
    assert np.array_equal(m.X[:n, 0], want) and np.array_equal(m.X[:n, 1], want)
    assert np.array_equal(m.X[n:], np.repeat(want[:, None], 3, axis=1))


def test_cpp_mesh_loader_mirror(a2ds, tmp_path):
    """a2ds::MeshLoader (host/MeshLoader.h): TACSMeshLoader's method names over the C ABI"""
    import subprocess
    exe = os.path.join(ROOT, "tests", "_mesh_loader_probe")
    src = os.path.join(ROOT, "tests", "mesh_loader_probe.cpp")
    hdr = os.path.join(ROOT, "a2d-shells_b200", "host", "MeshLoader.h")
    lib = os.path.join(ROOT, "a2d-shells_b200", "lib")
    a2ds.load_library()
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++11", "-O1", "-I" + os.path.join(ROOT, "include"), src,
                               "-o", exe, "-L" + lib, "-la2ds_b200", "-Wl,-rpath," + lib])
    deck = os.path.join(GOLD, "cyl_small.bdf")
    out = subprocess.run([exe, deck, str(tmp_path / "m.a2dm")], capture_output=True, text=True)
    assert "MESH_LOADER_OK" in out.stdout, out.stdout + out.stderr
    assert "unable to open file" in out.stdout
    m = a2ds.Mesh.read_bdf(deck)
    csum = int((m.elem_conn.astype(np.int64) * (1 + np.arange(len(m.elem_conn)) % 7)).sum())
    xsum = 0.0
    for v in m.X.ravel():
        xsum += v
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("MESH ")][0].split()
    assert [int(x) for x in line[1:6]] == [m.n_nodes, m.n_elems, m.n_bcs, m.n_comp, csum]
    assert float(line[6]) == xsum and int(line[7]) == m.bc_size
    assert f"FIND 0 {m.n_nodes - 1} -1" in out.stdout
    assert "COMP 0 [CQUAD4] [SKIN]" in out.stdout and "COMP 1 [CQUAD4] [RIB.001]" in out.stdout
    assert f"BINARY {m.n_nodes} {m.n_elems}" in out.stdout


def test_real_conversion_is_correctly_rounded(a2ds, tmp_path):
    """the reader converts short decimals itself (one exact multiply / divide) and leaves the
    rest to strtod: every value must carry the bits of the correctly rounded conversion —
    what the reference's atof gives and what Python's float() gives"""
    rng = np.random.default_rng(11)
    texts = []
    for _ in range(6000):
        nd = int(rng.integers(1, 14))
        digs = "".join(str(d) for d in rng.integers(0, 10, nd))
        point = int(rng.integers(0, nd + 1))
        mant = (digs[:point] or "") + "." + digs[point:]
        sign = "-" if rng.random() < 0.5 else ""
        style = rng.integers(0, 4)
        ex = int(rng.integers(-30, 31))
        body = sign + mant + ("", f"E{ex:+d}", f"D{ex:+d}", f"{ex:+d}")[style]
        if len(body) <= 16 and any(ch != "0" for ch in digs):
            texts.append(body)
    texts = texts[:3 * (len(texts) // 3)]
    lines = ["BEGIN BULK"]
    for k in range(0, len(texts), 3):
        a, b, c = texts[k:k + 3]
        lines.append(f"GRID*   {k // 3 + 1:>16d}{'':16s}{a:>16s}{b:>16s}*")
        lines.append(f"*       {c:>16s}")
    p = str(tmp_path / "rounding.bdf")
    open(p, "w").write("\n".join(lines) + "\n")
    m = a2ds.Mesh.read_bdf(p)

    def py(t):   # the compact form spelled out, then Python's correctly rounded conversion
        t = t.replace("D", "E")
        if "E" not in t:
            body = t.lstrip("-")
            for i, ch in enumerate(body):
                if ch in "+-":
                    t = ("-" if t.startswith("-") else "") + body[:i] + "E" + body[i:]
                    break
        return float(t)

    want = np.array([py(t) for t in texts]).reshape(-1, 3)
    assert m.X.tobytes() == want.tobytes()


@pytest.mark.gpu
def test_deck_to_assembly_on_the_device(a2ds, orc):
    """the shipped-example flow without the reference: deck -> a2ds_mesh_quad4 -> set_mesh /
    set_bcs -> residual + K + G, checked against the oracle on the same arrays"""
    from helpers import relmax
    m = a2ds.Mesh.read_bdf(os.path.join(GOLD, "cyl_large.bdf"))
    conn, masks, vals = m.quad4()
    u = a2ds.meshes.seeded_state(m.node_nums, 1e-5)
    Cs, eth = a2ds.iso_shell_tables()
    comps_Cs = np.repeat(Cs[None], m.n_comp, axis=0) * (1.0 + 0.25 * np.arange(m.n_comp))[:, None]
    asm = a2ds.Assembler(0)
    asm.set_mesh(conn, m.n_nodes, elem_comp=m.elem_comp)
    asm.set_nodes(m.X)
    asm.set_components(comps_Cs, np.repeat(eth[None], m.n_comp, axis=0))
    asm.set_bcs(m.bc_nodes, masks, vals)
    asm.set_state(u)
    kmat, gmat = asm.create_mat(), asm.create_mat()
    res = asm.assembleAll(kmat, gmat)
    rowp, cols = asm.mat_pattern(kmat)
    K, G = asm.mat_values(kmat), asm.mat_values(gmat)
    asm.close()
    comps = [orc.make_comp(0, comps_Cs[k], eth) for k in range(m.n_comp)]
    r_o, k_o = orc.assemble(1, conn, m.elem_comp, comps, m.X, u, rowp, cols, m.bc_nodes, masks, vals)
    _, g_o = orc.assemble(3, conn, m.elem_comp, comps, m.X, u, rowp, cols, m.bc_nodes, masks, vals)
    assert relmax(res, r_o) < 1e-12 and relmax(K, k_o) < 1e-10 and relmax(G, g_o) < 1e-10


def test_shipped_cylinder_deck_assembly_oracle_vs_reference(a2ds, orc, ref):
    """BASELINE configs[0] as shipped (examples/cylinder-buckling/mech-cylinder.bdf, 3200 MITC4
    elements): deck read by OUR reader, then residual, Kmat and Gmat of the oracle against the
    unmodified reference assembling the same arrays with the example's section and BCs —
    pins the oracle at the size and numbering of the shipped example (build container only)"""
    from helpers import relmax
    deck = "/root/reference/examples/cylinder-buckling/mech-cylinder.bdf"
    if not os.path.exists(deck):
        pytest.skip("reference examples not present")
    m = a2ds.Mesh.read_bdf(deck)
    conn, masks, vals = m.quad4()
    n = m.n_nodes
    bc_vars = [[k for k in range(6) if mk >> k & 1] for mk in masks]
    bc_vals = [[vals[b, k] for k in v] for b, v in enumerate(bc_vars)]
    props = ref.iso_props()                       # the section of mechBuckling.cpp:41-58
    ra = ref.RefAssembler(conn, m.X, m.elem_comp, props[None], m.bc_nodes, bc_vars, bc_vals)
    try:
        conn_r, X_r = ra.conn(), ra.nodes()
        nodes_b, vars_b, vals_b = ra.bcs()
        u = np.zeros((n, 6)); u[ra.new_nodes] = a2ds.meshes.seeded_state(m.node_nums, 1e-5)
        ra.set_state(u)
        mat = ra.mat_create(0)
        r_ref = ra.assemble_jacobian(mat)
        blk = ra.mat_block(mat, 0)
        rowp, cols, K_ref = blk["rowp"], blk["cols"], blk["A"]
        ra.assemble_mat_type(1, mat)
        G_ref = ra.mat_block(mat, 0)["A"]
    finally:
        ra.close()
    assert len(conn_r) == 3200 and X_r.shape == (3300, 3) and masks.sum() > 0
    Cs, eth, mom = ref.con_tables(props)
    comp = orc.make_comp(0, Cs, eth, mom)
    ec = np.zeros(len(conn_r), dtype=np.int32)
    r, K = orc.assemble(1, conn_r, ec, [comp], X_r, u, rowp, cols, nodes_b, vars_b, vals_b)
    _, G = orc.assemble(3, conn_r, ec, [comp], X_r, u, rowp, cols, nodes_b, vars_b, vals_b)
    assert relmax(r, r_ref) < 1e-12 and relmax(K, K_ref) < 1e-13 and relmax(G, G_ref) < 1e-9
    # the prescribed end shortening of the deck arrives in the residual rows: r = u - ubar
    b = int(np.argmin(vals_b.min(axis=1)))
    nd, k = int(nodes_b[b]), int(np.argmin(vals_b[b]))
    assert vals_b[b, k] == -1e-5 and r_ref[nd, k] == u[nd, k] + 1e-5
