"""The kernel's per-lane arithmetic (a2d-shells_b200/csrc/mitc4_math.h) stepped on the host
(tests/host_emul.cpp, compiled with FMA contraction like nvcc) against the golden
fixtures and the oracle.  CPU only — the GPU tests check the same code on the device."""
import os

import numpy as np
import pytest

from helpers import emul_element, random_elements, relmax

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("tr", [0, 1])
@pytest.mark.parametrize("ci", [0, 1])
@pytest.mark.parametrize("tying", [False, True])
def test_kernel_math_against_golden(emul, kind, tr, ci, tying):
    """both formulations of the element kernels stepped on the host: the first one (strain rows
    per Gauss point, k_assemble) and the tying-level one (mitc4_tying.h, k_assemble_t)"""
    g = np.load(os.path.join(GOLD, "elements.npz"))
    key = f"k{kind}_t{tr}_c{ci}"
    T = float(g[key + "_T"])
    if tying and np.any(g[key + "_Cs"][6:12] != 0.0):
        pytest.skip("coupled section: the tying-level kernel is not used for it")
    for e in range(g["X"].shape[0]):
        r, K, G = emul_element(emul, g[key + "_Cs"], g[key + "_eth"], T, kind, tr, g["axis"],
                               g["X"][e], g["q"][e], 1 if kind == 0 else 0, tying=tying)
        assert relmax(r, g[key + "_res"][e]) < 1e-12      # north_star residual tolerance
        assert relmax(K, g[key + "_K"][e]) < 1e-10        # north_star matrix tolerance
        if kind == 0:
            assert relmax(G, g[key + "_G"][e]) < (1e-10 if T == 0.0 else 1e-6)
            assert np.abs(G - G.T).max() <= 1e-13 * np.abs(G).max()


def test_tying_level_formulation_against_oracle_many(emul, orc, a2ds):
    """mitc4_tying.h (K = Bt^T H Bt over the tying points): 300 random distorted elements, both
    classes, both transforms, with and without temperature — residual, tangent and geometric
    stiffness against the oracle and against the first formulation"""
    X, q = random_elements(300, seed=17)
    axis = np.array([0.3, 1.0, 0.2])
    Cs, eth = a2ds.iso_shell_tables()
    for kind in (0, 1):
        for tr in (0, 1):
            for T in (0.0, 10.0):
                comp = orc.make_comp(kind, Cs, eth, (0, 0, 0), T, tr, axis)
                worst = [0.0, 0.0, 0.0]
                for e in range(X.shape[0]):
                    wg = 1 if kind == 0 else 0
                    r, K, G = emul_element(emul, Cs, eth, T, kind, tr, axis, X[e], q[e], wg, tying=True)
                    ro, ko = orc.jacobian(comp, X[e].ravel(), q[e].ravel())
                    worst[0] = max(worst[0], relmax(r, ro)); worst[1] = max(worst[1], relmax(K, ko))
                    if wg:
                        _, _, G0 = emul_element(emul, Cs, eth, T, kind, tr, axis, X[e], q[e], 1)
                        worst[2] = max(worst[2], relmax(G, G0))
                        assert np.abs(G - G.T).max() <= 1e-13 * np.abs(G).max()
                assert worst[0] < 1e-12 and worst[1] < 1e-11 and worst[2] < 1e-13, worst


def test_kernel_math_against_oracle_many(emul, orc, a2ds):
    """400 random elements: tight agreement with the oracle (same exact-derivative
    definition), including the residual whose rounding follows the reference's order"""
    X, q = random_elements(400, seed=11)
    axis = np.array([0.3, 1.0, 0.2])
    Cs, eth = a2ds.iso_shell_tables(t_offset=0.25)
    for kind in (0, 1):
        for tr in (0, 1):
            comp = orc.make_comp(kind, Cs, eth, (0, 0, 0), 0.0, tr, axis)
            worst = [0.0, 0.0]
            for e in range(X.shape[0]):
                r, K, _ = emul_element(emul, Cs, eth, 0.0, kind, tr, axis, X[e], q[e], 0)
                ro, ko = orc.jacobian(comp, X[e].ravel(), q[e].ravel())
                worst[0] = max(worst[0], relmax(r, ro)); worst[1] = max(worst[1], relmax(K, ko))
            assert worst[0] < 1e-12 and worst[1] < 1e-11, worst  # badly conditioned quads included


def test_geometric_stiffness_is_exact_linear_part(emul, orc, a2ds):
    """our G is the linear-in-state part of the nonlinear tangent:
    G(u) = 1/2 (K_nl(u) - K_nl(-u)) exactly, where the reference takes a noisy difference"""
    X, q = random_elements(20, seed=5, state_range=(-4, -3))
    Cs, eth = a2ds.iso_shell_tables()
    comp_nl = orc.make_comp(1, Cs, eth)
    for e in range(X.shape[0]):
        _, _, G = emul_element(emul, Cs, eth, 0.0, 0, 0, (1, 0, 0), X[e], q[e], 1)
        kp = orc.mat_type(comp_nl, 0, X[e].ravel(), q[e].ravel())
        km = orc.mat_type(comp_nl, 0, X[e].ravel(), -q[e].ravel())
        assert relmax(G, 0.5 * (kp - km)) < 1e-10


def test_iso_tables_match_reference(a2ds, ref):
    for off in (0.0, 0.3):
        Cs, eth = a2ds.iso_shell_tables(t_offset=off)
        Cr, er, _ = ref.con_tables(ref.iso_props(t_offset=off))
        assert relmax(Cs, Cr) < 1e-15 and np.array_equal(eth, er)


def test_mass_block_against_oracle(emul, orc):
    """the mass kernel's per-pair block (mass_block, mitc4_math.h) stepped on the host
    against the oracle's restatement of TACS_MASS_MATRIX"""
    import ctypes as C
    X, q = random_elements(30, seed=21)
    mom = np.array([27.18, -0.8154, 0.0246])   # m1 != 0: offset reference surface
    axis = np.array([0.3, 1.0, 0.2]); axis /= np.linalg.norm(axis)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for tr in (0, 1):
        comp = orc.make_comp(0, np.ones(22), np.zeros(9), mom, 0.0, tr, axis)
        for e in range(X.shape[0]):
            Xe = np.ascontiguousarray(X[e].ravel())
            M = np.zeros(576)
            emul.emul_mass(p(mom), C.c_int(tr), p(axis), p(Xe), p(M))
            M_or = orc.mat_type(comp, 2, Xe, np.zeros(24))
            assert relmax(M.reshape(24, 24), M_or) < 1e-13


def test_tying_rows_have_rank_nine_across_the_gauss_points(emul, a2ds):
    """Structure behind a cheaper contraction (profiles/README.md, experiment 1): the membrane /
    transverse-shear rows of B0 and of B1(q) at the four Gauss points are interpolations of the
    same 9 tying-point rows, so the 20 x 24 stack has rank 9 — K and Z could contract over 9
    tying points (3 k-steps of the m8n8k4 MMA) instead of 5 strains x 4 points (5 k-steps)"""
    import ctypes as C
    X, q = random_elements(6, seed=3)
    Cs, _ = a2ds.iso_shell_tables()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for tr, axis in ((0, np.array([1.0, 0.0, 0.0])), (1, np.array([0.6, 0.8, 0.0]))):
        for e in range(X.shape[0]):
            B0 = np.zeros((4, 9, 24)); B1 = np.zeros((4, 9, 24))
            emul.emul_strain_matrices(p(np.ascontiguousarray(Cs)), C.c_int(tr), p(axis),
                                      p(np.ascontiguousarray(X[e]).ravel()),
                                      p(np.ascontiguousarray(q[e]).ravel()), p(B0), p(B1))
            assert np.abs(B1[:, 8]).max() == 0.0          # the drilling strain is linear in the state
            for B in (B0, B1):
                sv = np.linalg.svd(B[:, [0, 1, 2, 6, 7], :].reshape(20, 24), compute_uv=False)
                assert sv[8] > 1e-6 * sv[0] and sv[9] < 1e-13 * sv[0]


@pytest.mark.parametrize("tr", [0, 1])
@pytest.mark.parametrize("ci", [0, 1])
def test_quad9_kernel_math_against_golden(emul, tr, ci):
    """the work items of k_assemble9 (a2d-shells_b200/csrc/mitc9_math.h: node / tying point /
    Gauss point frames, the tying and drill derivative tables, the columns of B) stepped on the
    host against the reference's TACSQuad9Shell (tests/golden/quad9.npz)"""
    import ctypes as C
    g = np.load(os.path.join(GOLD, "quad9.npz"))
    key = f"k0_t{tr}_c{ci}"
    T = float(g[key + "_T"])
    p = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.c_void_p)
    ax = g["axis"] / np.linalg.norm(g["axis"])
    for e in range(g["X"].shape[0]):
        res = np.zeros(54); K = np.zeros(54 * 54)
        G = np.zeros(54 * 54)
        emul.emul_element9(p(g[key + "_Cs"]), p(g[key + "_eth"]), C.c_double(T), C.c_int(tr), p(ax),
                           p(g["X"][e]), p(g["q"][e]), p(res), p(K), C.c_int(0), p(G))
        assert relmax(res, g[key + "_res"][e]) < 1e-12
        assert relmax(K.reshape(54, 54), g[key + "_K"][e]) < 1e-10
        # geometric stiffness: analytic here, a central difference in the reference (noise 1e-12
        # without, ~1e-7 with a temperature)
        assert relmax(G.reshape(54, 54), g[key + "_G"][e]) < (1e-10 if T == 0.0 else 1e-6)
        # nonlinear strain model (TACSQuad9NonlinearShell): residual and tangent about the state
        keyn = f"k1_t{tr}_c{ci}"
        emul.emul_element9(p(g[keyn + "_Cs"]), p(g[keyn + "_eth"]), C.c_double(T), C.c_int(tr), p(ax),
                           p(g["X"][e]), p(g["q"][e]), p(res), p(K), C.c_int(1), None)
        assert relmax(res, g[keyn + "_res"][e]) < 1e-12
        assert relmax(K.reshape(54, 54), g[keyn + "_K"][e]) < 1e-10


def test_quad9_kernel_math_against_oracle_many(emul, orc, a2ds):
    from helpers import random_elements9
    import ctypes as C
    X, q = random_elements9(60, seed=77)
    Cs, eth = a2ds.iso_shell_tables(t_offset=0.2)
    p = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.c_void_p)
    axis = np.array([0.3, 1.0, 0.2]); ax = axis / np.linalg.norm(axis)
    worst = [0.0, 0.0]
    for tr in (0, 1):
        for T in (0.0, 25.0):
            comp = orc.make_comp(0, Cs, eth, (0, 0, 0), T, tr, axis)
            for e in range(len(X)):
                res = np.zeros(54); K = np.zeros(54 * 54)
                emul.emul_element9(p(Cs), p(eth), C.c_double(T), C.c_int(tr), p(ax), p(X[e]), p(q[e]),
                                   p(res), p(K), C.c_int(0), None)
                r_o, k_o = orc.jacobian(comp, X[e].ravel(), q[e].ravel(), order=3)
                worst[0] = max(worst[0], relmax(res, r_o))
                worst[1] = max(worst[1], relmax(K.reshape(54, 54), k_o))
    assert worst[0] < 1e-12 and worst[1] < 1e-10, worst


def test_quad9_mass_math_against_oracle(emul, orc):
    """q9_mass_pair (the work items of k_mass9) against the order-3 oracle's TACS_MASS_MATRIX,
    with a non-zero first mass moment"""
    from helpers import random_elements9
    import ctypes as C
    X, _ = random_elements9(20, seed=3)
    mom = np.array([27.18, 0.011, 2.3e-4])
    p = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.c_void_p)
    comp = orc.make_comp(0, np.ones(22), np.zeros(9), mom)
    for e in range(len(X)):
        M = np.zeros(54 * 54)
        emul.emul_mass9(p(mom), p(X[e]), p(M))
        m_o = orc.mat_type(comp, 2, X[e].ravel(), np.zeros(54), order=3)
        assert relmax(M.reshape(54, 54), m_o) < 1e-13
