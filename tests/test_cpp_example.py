"""The C++ face for drivers without TACS: a2ds::IsoShell (host tables), and the example
driver examples/cylinder_buckling_assembly.cpp (deck -> device Kmat / Gmat / residual)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "a2d-shells_b200", "lib")


def _build(src, exe, deps=()):
    srcs = [src] + [os.path.join(ROOT, "a2d-shells_b200", "host", d) for d in deps]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-std=c++11", "-O1", "-I" + os.path.join(ROOT, "include"), src,
                               "-o", exe, "-L" + LIB, "-la2ds_b200", "-Wl,-rpath," + LIB])
    return exe


def test_iso_shell_header_matches_python_tables(a2ds):
    a2ds.load_library()
    exe = _build(os.path.join(ROOT, "tests", "iso_shell_probe.cpp"),
                 os.path.join(ROOT, "tests", "_iso_shell_probe"), ["IsoShell.h"])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    rows = [np.array(ln.split()[1:], dtype=float) for ln in out.splitlines() if ln.startswith("SECTION")]
    assert len(rows) == 3
    for k, (row, off) in enumerate(zip(rows, (0.0, 0.3, -0.45))):
        t, E = 0.010 + 0.004 * k, 72e9 * (1 + k)
        Cs, eth = a2ds.iso_shell_tables(E=E, t=t, t_offset=off)
        mom = a2ds.iso_mass_moments(2718.0, t, off)
        assert np.array_equal(row[:22], Cs) and np.array_equal(row[22:31], eth)
        assert np.array_equal(row[31:], mom)


def test_iso_shell_header_matches_reference_constitutive(a2ds, ref):
    """the same tables from the reference's own TACSIsoShellConstitutive object"""
    exe = os.path.join(ROOT, "tests", "_iso_shell_probe")
    if not os.path.exists(exe):
        pytest.skip("probe not built")
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    rows = [np.array(ln.split()[1:], dtype=float) for ln in out.splitlines() if ln.startswith("SECTION")]
    for k, (row, off) in enumerate(zip(rows, (0.0, 0.3, -0.45))):
        p = ref.iso_props(E=72e9 * (1 + k), t=0.010 + 0.004 * k, t_offset=off)
        Cs, eth, mom = ref.con_tables(p)
        assert np.allclose(row[:22], Cs, rtol=1e-15, atol=0) and np.allclose(row[22:31], eth, rtol=1e-15)
        assert np.allclose(row[31:], mom, rtol=1e-15, atol=1e-300)


def _example():
    return _build(os.path.join(ROOT, "examples", "cylinder_buckling_assembly.cpp"),
                  os.path.join(ROOT, "tests", "_cylinder_assembly"),
                  ["IsoShell.h", "MeshLoader.h", "DeviceAssembler.h"])


def test_example_driver_compiles_and_fails_loudly_without_gpu(a2ds):
    a2ds.load_library()
    exe = _example()
    if has_gpu():
        pytest.skip("a GPU is present")
    out = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "cyl_large.bdf")],
                         capture_output=True, text=True)
    assert out.returncode == 1 and "80 elements" in out.stdout and "no CUDA device" in out.stdout


@pytest.mark.gpu
def test_example_driver_matches_the_python_path(a2ds):
    """deck -> Kmat, Gmat, residual through the C++ headers alone; checksums against the same
    assembly driven through the ctypes face"""
    exe = _example()
    deck = os.path.join(ROOT, "tests", "golden", "cyl_large.bdf")
    out = subprocess.run([exe, deck], capture_output=True, text=True)
    assert "CYLINDER_ASSEMBLY_OK" in out.stdout, out.stdout + out.stderr
    got = {}
    for ln in out.stdout.splitlines():
        if ln.startswith("CHECK "):
            _, name, n, mx, ws = ln.split()
            got[name] = (int(n[2:]), float(mx[4:]), float(ws[5:]))
    m = a2ds.Mesh.read_bdf(deck)
    conn, masks, vals = m.quad4()
    Cs, eth = a2ds.iso_shell_tables()
    asm = a2ds.Assembler(0)
    asm.set_mesh(conn, m.n_nodes, elem_comp=m.elem_comp); asm.set_nodes(m.X)
    asm.set_components(np.repeat(Cs[None], m.n_comp, 0), np.repeat(eth[None], m.n_comp, 0))
    asm.set_bcs(m.bc_nodes, masks, vals)
    asm.set_state(a2ds.meshes.seeded_state(m.node_nums, 1e-5))
    k, g = asm.create_mat(), asm.create_mat()
    asm.assembleMatType(a2ds.STIFFNESS_MATRIX, k)
    asm.assembleMatType(a2ds.GEOMETRIC_STIFFNESS_MATRIX, g)
    res = asm.assembleRes()
    want = dict(K=asm.mat_values(k).ravel(), G=asm.mat_values(g).ravel(), res=res.ravel())
    asm.close()
    for name, v in want.items():
        n, mx, ws = got[name]
        assert n == v.size
        # atomic accumulation order differs run to run: compare to rounding, not bits
        assert abs(mx - np.abs(v).max()) <= 1e-12 * np.abs(v).max()
        w = np.cos((np.arange(v.size) % 1000003).astype(float))
        assert abs(ws - float((v * w).sum())) <= 1e-9 * np.abs(v).sum()
