// mitc4_tying.h — the MITC4 element contraction at the TYING-POINT level (per-lane arithmetic
// of k_assemble_t; host-steppable like mitc4_math.h, see tests/host_emul.cpp).
//
// mitc4_math.h's first formulation contracts over (strain row, Gauss point): 9 x 4 = 36 rows,
// K = sum_qp B_qp^T (w C) B_qp.  The five membrane / transverse-shear rows of every Gauss point
// are however linear images of the SAME nine tying-point strains g_t (the MITC interpolation,
// TACSShellElementQuadBasis.h:530-672, and e0ty = S^T gty S, TACSShellElement.h:520-534):
//        e_ms(qp) = M_qp I_qp g,      g = Gt q  (Gt: 9 x 24, no Gauss point in it)
// so with the GENERALISED strain rows
//        Bt = [ bending rows of the 4 Gauss points (12) | tying rows Gt (9) | drill rows (4) ]
// the element tangent is  K = Bt^T H Bt  with a block matrix H that holds the constitutive
// data and all the Gauss-point geometry of the membrane/shear part:
//        H_tt = sum_qp I_qp^T (w M_qp^T C_ms M_qp) I_qp   (9 x 9, built once per element)
//        H_bb = w D per Gauss point,  H_dd = w k_drill per Gauss point.
// The contraction runs over 25 (padded 28 = 7 DMMA k-steps) instead of 36 rows (9 k-steps),
// G = Bt1^T (H Bt0) + (H Bt0)^T Bt1 over 21 (24 = 6 k-steps; the drill strain is linear in
// the state) instead of 32, and the tying rows of a lane are plain table entries: the
// 5 x 5 map M never touches the per-lane code (it is folded into H by one lane per Gauss
// point in the batched phase).  Per element the FP64 work drops from ~1200 scalar warp
// instructions + 126 DMMA to ~650 + 96.
//
// This path covers components without membrane-bending coupling (B block of the ABD matrix
// zero: isotropic / symmetric sections); coupled components stay on the first formulation
// (k_assemble), selected per element list on the host.
//
// Generalised row order (DMMA contraction index k = 4 s + kk, kk = lane & 3):
//   s = 0,1,2  bending strains e3, e4, e5 at Gauss point kk
//   s = 3      tying strains g11(eta=-1), g11(eta=+1), g22(xi=-1), g22(xi=+1)      TI 0,2,4,6
//   s = 4      tying strains g13(eta=-1), g13(eta=+1), g23(xi=-1), g23(xi=+1)      TI 1,3,5,7
//   s = 5      g12 at the centre (kk = 0; the other three slots are zero rows)     TI 8
//   s = 6      drilling strain at Gauss point kk
// Tying index TI (rows / columns of H): pairs (g11, g13) on the two eta edges, (g22, g23) on
// the two xi edges, g12 — a node's five non-zero tying rows then sit in two adjacent column
// pairs (2x, 2x+1), (4+2y, 5+2y) and column 8 (x = m / 2, y = m % 2).
#ifndef A2DS_MITC4_TYING_H
#define A2DS_MITC4_TYING_H

#include "mitc4_math.h"

namespace a2ds {

static const int TY_LD = 10;  // leading dimension of H; row 9 is a zero row (padding slots)
static const int TY_ROWS = 7; // k-steps of the tangent contraction (6 for B1: no drill row)

// Tying rows of one node, before the MITC interpolation: derivative of the node's tying
// strains w.r.t. its displacement (h = 0) / rotation (h = 1) — the columns of Gt.
struct NodeTab {
  double gm[10];    // h = 0 only: g11 row (3), g22 row (3), g12 row (3), pad
  double gs[2][6];  // [h]: g13 row (3), g23 row (3)
  double pad_[2];   // 24 doubles: the (node, h) pairs a half-warp reads fall into distinct banks
};

// Per Gauss point record (written by phase_qp_t, one lane per Gauss point)
struct QpRec {
  double t0[3], t1[3];  // T columns 0, 1
  double S[6], Sz[6];   // (Xd^-1 T)[i][j], its thickness derivative; i = 0..2, j = 0..1
  double wD[6];         // w * D   (bending block, symmetric storage)
  double wdrill;        // w * drilling stiffness
  double P0[6], P1[6];  // T u0x[:,j], T u1x[:,j] of the state
  double Pq[6];         // T T^T
  double Q[15];         // w M^T C_ms M, upper triangle, g5 order (g11, g12, g13, g22, g23)
  double sig[9];        // this point's share of the tying-point stresses (TI order)
  double sg[3];         // w s3, w s4, w s5
  double sd;            // w s8
};  // 71 doubles = 7 (mod 16)

struct ElemRec {
  double X[12], q[24];
  double fn[12], dr[12], wn[12], cdr[36], etn[4];  // as ElemGeom (phase_node)
  NodeTab t0[4];  // tying rows of B0 (field X, normals fn)
  NodeTab t1[4];  // tying rows of B1(q) (field u, directors)
  QpRec qp[4];
  double pad_[2];
};

// per element working set of the column phase (one per warp)
struct alignas(16) TyWork {
  double sigt[TY_LD];               // tying-point stresses summed over the Gauss points
  double H[TY_LD * TY_LD];          // H_tt (9 x 9) + zero row / column
  double ca[4][8][2], cb[4][8][2];  // geometric phase: coefficient pairs per Gauss point, gen. node
  double pad_[28];                  // H .. pad_ = 256 doubles: scratch of the geometric phase (mq of
                                    // the 64 node pairs x 4 Gauss points), once H, ca, cb are consumed
};

// g5 component (g11=0, g12=1, g13=2, g22=3, g23=4) of tying index t, 3 bits each
A2DS_HD int ty_comp(int t) { return (int)((0143432020u >> (3 * t)) & 7u); }
// interpolation weight of tying point t at Gauss point qp (evalTyingInterp, QuadBasis.h:569-616)
A2DS_HD double ty_iota(int qp, int t) {
  if (t >= 8) return 1.0;
  const int side = (t >> 1) & 1;
  const int bit = (t < 4) ? ((qp >> 1) & 1) : (qp & 1);  // g11/g13: linear in eta; g22/g23: in xi
  return (bit == side) ? 0.5 * (1.0 + A2DS_GAUSS_PT) : 0.5 * (1.0 - A2DS_GAUSS_PT);
}
A2DS_HD int ty_qidx(int a, int b) {  // upper triangle of the 5 x 5, row major
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return (lo * (11 - lo)) / 2 + hi - lo;
}

// One of the 45 entries (t <= tp) of H_tt: its four Gauss-point weights and the Q slot
struct TyPlan {
  double w[4];
  int qidx, t, tp;
};
A2DS_HD void ty_plan(int entry, TyPlan &pl) {
  // entry -> (t, tp) of the upper triangle, row major
  int t = 0, rem = entry;
  while (rem >= 9 - t) { rem -= 9 - t; t++; }
  const int tp = t + rem;
  pl.t = t; pl.tp = tp;
  pl.qidx = ty_qidx(ty_comp(t), ty_comp(tp));
  for (int qp = 0; qp < 4; qp++) pl.w[qp] = ty_iota(qp, t) * ty_iota(qp, tp);
}
template <class Rec>
A2DS_HD void ty_H_entry(const Rec &s, const TyPlan &pl, double *H) {
  const double h = pl.w[0] * s.qp[0].Q[pl.qidx] + pl.w[1] * s.qp[1].Q[pl.qidx] +
                   pl.w[2] * s.qp[2].Q[pl.qidx] + pl.w[3] * s.qp[3].Q[pl.qidx];
  H[TY_LD * pl.t + pl.tp] = h;
  H[TY_LD * pl.tp + pl.t] = h;
}

// ---- batched node phase, second half: the node's tying rows ---------------------------
// (after phase_node of all four nodes: needs the neighbours' normals / directors)
//   g11(eta edge of m) = u,xi . X,xi          -> d/du_m = N_m,xi X,xi(edge)
//   g22(xi edge)       = u,eta . X,eta
//   g12(centre)        = 1/2 (u,xi . X,eta + u,eta . X,xi)
//   g13(eta edge)      = 1/2 (X,xi . d0 + n0 . u,xi),   g23(xi edge) likewise with eta
// (TACSShellElementModel.h:50-74; nonlinear additions :661-694 give the same rows with
// (X, fn) replaced by (u, d): B1(q).)  Rotation columns: d_m = theta_m x fn_m.
A2DS_HD void node_tab(NodeTab &t, const double *field, int ld, const double *nrm,
                      const double *fnm, int m) {
  const int sx = m % 2, sy = m / 2, mx = m ^ 1, my = m ^ 2;
  const double dN = sx ? 0.5 : -0.5, dM = sy ? 0.5 : -0.5;
  double fxi[3], feta[3];
  edge_xi(field, ld, sy, fxi);
  edge_eta(field, ld, sx, feta);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double cxi = 0.25 * (field[ld + k] - field[k] + field[3 * ld + k] - field[2 * ld + k]);
    const double ceta = 0.25 * (field[2 * ld + k] - field[k] + field[3 * ld + k] - field[ld + k]);
    t.gm[k] = dN * fxi[k];
    t.gm[3 + k] = dM * feta[k];
    t.gm[6 + k] = 0.25 * (dN * ceta + dM * cxi);
    const double n_eta_edge = 0.5 * (nrm[3 * m + k] + nrm[3 * mx + k]);  // nodes (m, m^1)
    const double n_xi_edge = 0.5 * (nrm[3 * m + k] + nrm[3 * my + k]);   // nodes (m, m^2)
    t.gs[0][k] = (0.5 * dN) * n_eta_edge;
    t.gs[0][3 + k] = (0.5 * dM) * n_xi_edge;
  }
  t.gm[9] = 0.0;
  double x13[3], x23[3];
  cross(fnm, fxi, x13);
  cross(fnm, feta, x23);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    t.gs[1][k] = 0.25 * x13[k];
    t.gs[1][3 + k] = 0.25 * x23[k];
  }
}
A2DS_HD void phase_node_tab(ElemRec &s, int m, bool want_state) {
  node_tab(s.t0[m], s.X, 3, s.fn, &s.fn[3 * m], m);
  if (want_state) node_tab(s.t1[m], s.q, 6, s.dr, &s.fn[3 * m], m);
}

// ---- batched Gauss point phase -------------------------------------------------------
template <class Rec>
A2DS_HD void phase_qp_t(const CompData &c, const Rec &s, QpRec &d, int qp, const Want &w) {
  const bool need_state = w.gmat || w.nonlinear;
  const bool want_e = w.res || need_state;
  QpGeom g;
  qp_geometry(c, s, qp, want_e, need_state, w.nonlinear, g);
#pragma unroll
  for (int k = 0; k < 3; k++) { d.t0[k] = g.t0[k]; d.t1[k] = g.t1[k]; }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) { d.S[2 * i + j] = g.S[3 * i + j]; d.Sz[2 * i + j] = g.Sz[3 * i + j]; }
  const double *A = &c.Cs[0], *D = &c.Cs[12], *As = &c.Cs[18];
#pragma unroll
  for (int i = 0; i < 6; i++) d.wD[i] = g.w * D[i];
  d.wdrill = g.w * c.Cs[21];
  if (need_state) {
#pragma unroll
    for (int i = 0; i < 6; i++) { d.P0[i] = g.P0[i]; d.P1[i] = g.P1[i]; }
    const double *t0 = g.t0, *t1 = g.t1, *tn = g.tn;
    d.Pq[0] = t0[0] * t0[0] + t1[0] * t1[0] + tn[0] * tn[0];
    d.Pq[1] = t0[0] * t0[1] + t1[0] * t1[1] + tn[0] * tn[1];
    d.Pq[2] = t0[0] * t0[2] + t1[0] * t1[2] + tn[0] * tn[2];
    d.Pq[3] = t0[1] * t0[1] + t1[1] * t1[1] + tn[1] * tn[1];
    d.Pq[4] = t0[1] * t0[2] + t1[1] * t1[2] + tn[1] * tn[2];
    d.Pq[5] = t0[2] * t0[2] + t1[2] * t1[2] + tn[2] * tn[2];
  }
  if (w.kmat || w.gmat) {
    // Q = w M^T C_ms M, C_ms = diag(A (3x3), As (2x2)) on the rows (e0, e1, e2, e6, e7)
    double CM[5][5];
#pragma unroll
    for (int cc = 0; cc < 5; cc++) {
      const double m0 = g.M[cc], m1 = g.M[5 + cc], m2 = g.M[10 + cc], m3 = g.M[15 + cc],
                   m4 = g.M[20 + cc];
      CM[0][cc] = g.w * (A[0] * m0 + A[1] * m1 + A[2] * m2);
      CM[1][cc] = g.w * (A[1] * m0 + A[3] * m1 + A[4] * m2);
      CM[2][cc] = g.w * (A[2] * m0 + A[4] * m1 + A[5] * m2);
      CM[3][cc] = g.w * (As[0] * m3 + As[1] * m4);
      CM[4][cc] = g.w * (As[1] * m3 + As[2] * m4);
    }
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 5; a++)
#pragma unroll
      for (int b = a; b < 5; b++, idx++)
        d.Q[idx] = g.M[a] * CM[0][b] + g.M[5 + a] * CM[1][b] + g.M[10 + a] * CM[2][b] +
                   g.M[15 + a] * CM[3][b] + g.M[20 + a] * CM[4][b];
  }
  if (want_e) {
    // stresses of the state (TACSShellConstitutive::computeStress) pulled back to the
    // generalised rows: dU/dg5 = M^T (w s_ms), then the tying interpolation transposed
    const double Tth = w.thermal * c.temperature;
    double e[9], st[9];
#pragma unroll
    for (int r = 0; r < 9; r++) e[r] = g.e[r] - Tth * c.eth[r];
    apply_C(c.Cs, e, st, false);
    const double sm[5] = {g.w * st[0], g.w * st[1], g.w * st[2], g.w * st[6], g.w * st[7]};
    double dg[5];
#pragma unroll
    for (int cc = 0; cc < 5; cc++)
      dg[cc] = g.M[cc] * sm[0] + g.M[5 + cc] * sm[1] + g.M[10 + cc] * sm[2] +
               g.M[15 + cc] * sm[3] + g.M[20 + cc] * sm[4];
    d.sig[0] = g.nb[0] * dg[0]; d.sig[1] = g.nb[0] * dg[2];
    d.sig[2] = g.nb[1] * dg[0]; d.sig[3] = g.nb[1] * dg[2];
    d.sig[4] = g.na[0] * dg[3]; d.sig[5] = g.na[0] * dg[4];
    d.sig[6] = g.na[1] * dg[3]; d.sig[7] = g.na[1] * dg[4];
    d.sig[8] = dg[1];
    d.sg[0] = g.w * st[3]; d.sg[1] = g.w * st[4]; d.sg[2] = g.w * st[5];
    d.sd = g.w * st[8];
  }
}

// tying-point stresses summed over the Gauss points (lanes 0..8 of the element prologue)
template <class Rec>
A2DS_HD void ty_sum_stress(const Rec &s, TyWork &wk, int t) {
  wk.sigt[t] = s.qp[0].sig[t] + s.qp[1].sig[t] + s.qp[2].sig[t] + s.qp[3].sig[t];
}

// ---- column phase: lane = (kk = lane & 3, m, h) -------------------------------------
// The lane's three columns (DOFs 6 m + 3 h + 0..2) of the generalised rows k = 4 s + kk.

// Lane constants of the column phase (lane = (kk, m, h)): kept in registers across the
// element loop (made opaque to the compiler in the kernel so that they are not recomputed
// per element)
struct LaneConst {
  double Nxi, Neta, N;  // N_m,xi  N_m,eta  N_m at Gauss point kk
  double Nq[4];         // N_n(kk), n = 0..3; 0 for the node diagonally opposite to m
};
A2DS_HD void lane_const(int lane, LaneConst &lc) {
  const int kk = lane_qp(lane), m = lane_m(lane);
  double na[2], nb[2];
  qp_shape(kk, na, nb);
  const double dN = (m % 2) ? 0.5 : -0.5, dM = (m / 2) ? 0.5 : -0.5;
  const double nam = (m % 2) ? na[1] : na[0], nbm = (m / 2) ? nb[1] : nb[0];
  lc.Nxi = dN * nbm; lc.Neta = nam * dM; lc.N = nam * nbm;
  for (int n = 0; n < 4; n++) lc.Nq[n] = ((m ^ n) == 3) ? 0.0 : na[n % 2] * nb[n / 2];
}

// non-zero tying rows of the lane's columns (g11, g13, g22, g23, g12) and its own slots s = 3,4,5
// zero10: ten zeros (the rotation columns have no g11 / g22 / g12 rows)
A2DS_HD void lane_tying(const NodeTab &t, const double *zero10, int m, int h, int kk,
                        double Gnz[5][3], double R[3][3]) {
  const int x = m / 2, y = m % 2;
  const double *gm = h ? zero10 : t.gm;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    Gnz[0][k] = gm[k];
    Gnz[1][k] = t.gs[h][k];
    Gnz[2][k] = gm[3 + k];
    Gnz[3][k] = t.gs[h][3 + k];
    Gnz[4][k] = gm[6 + k];
  }
  // slot kk of s = 3, 4 is a tying point on the eta edge kk (kk < 2) / the xi edge kk - 2
  const bool on = kk < 2 ? (x == kk) : (y == kk - 2);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    R[0][k] = on ? (kk < 2 ? Gnz[0][k] : Gnz[2][k]) : 0.0;
    R[1][k] = on ? (kk < 2 ? Gnz[1][k] : Gnz[3][k]) : 0.0;
    R[2][k] = kk == 0 ? Gnz[4][k] : 0.0;
  }
}

// W rows of the lane's tying slots: W[s] = sum_t' H[row(s, kk)][t'] Gt[t'][cols]
A2DS_HD void lane_w_tying(const double *H, int m, int kk, const double Gnz[5][3],
                          double W[3][3]) {
  const int x = m / 2, y = m % 2;
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const int row = (s < 2) ? 2 * kk + s : (kk == 0 ? 8 : 9);
    const double *h = H + TY_LD * row;
    const double h0 = h[2 * x], h1 = h[2 * x + 1], h2 = h[4 + 2 * y], h3 = h[5 + 2 * y], h4 = h[8];
#pragma unroll
    for (int k = 0; k < 3; k++)
      W[s][k] = h0 * Gnz[0][k] + h1 * Gnz[1][k] + h2 * Gnz[2][k] + h3 * Gnz[3][k] + h4 * Gnz[4][k];
  }
}

A2DS_HD void node_coef_t(const QpRec &g, const LaneConst &lc, NodeCoef &n) {
  const double Nxi = lc.Nxi, Neta = lc.Neta, N = lc.N;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    n.a[j] = Nxi * g.S[j] + Neta * g.S[2 + j];
    n.az[j] = Nxi * g.Sz[j] + Neta * g.Sz[2 + j];
    n.b[j] = N * g.S[4 + j];
    n.cc[j] = n.a[j] + N * g.Sz[4 + j];
  }
}

// bending rows (e3, e4, e5 at Gauss point kk) of B0 and the drilling row
template <class Rec>
A2DS_HD void lane_bend0(const Rec &s, int m, int h, int qp, const LaneConst &lc,
                        const NodeCoef &nc, double R[3][3], double Rd[3]) {
  const QpRec &g = s.qp[qp];
  const double *fn = &s.fn[3 * m];
  const double cz0 = h ? nc.cc[0] : nc.az[0], cz1 = h ? nc.cc[1] : nc.az[1];
  double X0[3], X1[3];
  cross(fn, g.t0, X0);
  cross(fn, g.t1, X1);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double z0 = h ? X0[k] : g.t0[k], z1 = h ? X1[k] : g.t1[k];
    R[0][k] = cz0 * z0;
    R[1][k] = cz1 * z1;
    R[2][k] = cz1 * z0 + cz0 * z1;
  }
  // drilling strain row: et(qp) = sum_n N_n etn_n,
  // etn_n = 1/2 (u0x[1][0] - u0x[0][1]) - theta_n . (t0n x t1n)   (TACSDirector.h:560-564)
  double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int n = 0; n < 4; n++) {
    const int slot = m ^ n;  // 3: diagonally opposite node, no contribution (Nq = 0)
    const double *cd = &s.cdr[9 * n + 3 * (slot & ~(slot >> 1))];
#pragma unroll
    for (int k = 0; k < 3; k++) acc[k] += lc.Nq[n] * cd[k];
  }
#pragma unroll
  for (int k = 0; k < 3; k++) Rd[k] = h ? -lc.N * s.wn[3 * m + k] : acc[k];
}

// bending rows of B1(q); publishes the coefficient pairs of the geometric phase
template <class Rec>
A2DS_HD void lane_bend1(const Rec &s, TyWork &wk, int m, int h, int qp, const NodeCoef &nc,
                        double R[3][3]) {
  const QpRec &g = s.qp[qp];
  const double *fn = &s.fn[3 * m];
  const double ca0 = h ? nc.b[0] : nc.a[0], ca1 = h ? nc.b[1] : nc.a[1];
  const double cz0 = h ? nc.cc[0] : nc.az[0], cz1 = h ? nc.cc[1] : nc.az[1];
  const int p = 4 * h + m;
  wk.ca[qp][p][0] = ca0; wk.ca[qp][p][1] = ca1;
  wk.cb[qp][p][0] = cz0; wk.cb[qp][p][1] = cz1;
  double r[3][3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    r[0][k] = ca0 * g.P1[k] + cz0 * g.P0[k];
    r[1][k] = ca1 * g.P1[3 + k] + cz1 * g.P0[3 + k];
    r[2][k] = ca0 * g.P1[3 + k] + cz1 * g.P0[k] + ca1 * g.P1[k] + cz0 * g.P0[3 + k];
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    double x[3];
    cross(fn, r[j], x);  // rotation columns: d = theta x fn
#pragma unroll
    for (int k = 0; k < 3; k++) R[j][k] = h ? x[k] : r[j][k];
  }
}

// The lane's fragments.  B[s][k], W[s][k] (s = 0..6), B1[s][k] (s = 0..5).
//   nonlinear model: B = B0 + B1(q), W = H B           (B1 is then not returned separately)
//   geometric stiffness: B1 returned, W = H B0
struct LaneFrag {
  double B[TY_ROWS][3], W[TY_ROWS][3];
};

template <class Rec>
A2DS_HD void lane_fragments(const Rec &s, TyWork &wk, int lane, const LaneConst &lc, const Want &w,
                            LaneFrag &f, double B1[6][3]) {
  const int kk = lane_qp(lane), m = lane_m(lane), h = lane_h(lane);
  const double *zero10 = &wk.H[TY_LD * 9];
  NodeCoef nc;
  node_coef_t(s.qp[kk], lc, nc);
  double Gnz[5][3];
  lane_bend0(s, m, h, kk, lc, nc, &f.B[0], f.B[6]);
  lane_tying(s.t0[m], zero10, m, h, kk, Gnz, &f.B[3]);
  if (w.gmat || w.nonlinear) {
    double G1[5][3];
    lane_bend1(s, wk, m, h, kk, nc, &B1[0]);
    lane_tying(s.t1[m], zero10, m, h, kk, G1, &B1[3]);
    if (w.nonlinear) {
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) f.B[r][k] += B1[r][k];
#pragma unroll
      for (int r = 0; r < 5; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) Gnz[r][k] += G1[r][k];
    }
  }
  if (w.kmat || w.gmat) {
    const QpRec &g = s.qp[kk];
    const double *D = g.wD;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double b0 = f.B[0][k], b1 = f.B[1][k], b2 = f.B[2][k];
      f.W[0][k] = D[0] * b0 + D[1] * b1 + D[2] * b2;
      f.W[1][k] = D[1] * b0 + D[3] * b1 + D[4] * b2;
      f.W[2][k] = D[2] * b0 + D[4] * b1 + D[5] * b2;
      f.W[6][k] = g.wdrill * f.B[6][k];
    }
    lane_w_tying(wk.H, m, kk, Gnz, &f.W[3]);
  }
}

// residual partial of the lane (to be summed over the four kk lanes of a column group):
// r = Bt^T sigma with the generalised stresses (w s_b per point | tying stresses | w s8)
template <class Rec>
A2DS_HD void lane_residual(const Rec &s, const TyWork &wk, int lane, const LaneFrag &f,
                           double r3[3]) {
  const int kk = lane_qp(lane);
  const QpRec &g = s.qp[kk];
  const double s3 = wk.sigt[2 * kk], s4 = wk.sigt[2 * kk + 1], s5 = wk.sigt[8];
#pragma unroll
  for (int k = 0; k < 3; k++)
    r3[k] = f.B[0][k] * g.sg[0] + f.B[1][k] * g.sg[1] + f.B[2][k] * g.sg[2] + f.B[3][k] * s3 +
            f.B[4][k] * s4 + f.B[5][k] * s5 + f.B[6][k] * g.sd;
}

// ---- geometric stiffness: one 3x3 block for the generalised node pair (p, pp) -----------
// p, pp in 0..7: 0..3 displacement of node p, 4..7 director of node p - 4 (as geo_block in
// mitc4_math.h, on the records of this formulation).  Three pieces:
//   geo_tying_scalar: the scalar on the identity from the tying-point stresses
//   geo_bending_mq:   per Gauss point, the scalar multiplying T T^T (bending stresses)
//   geo_fold:         rows / columns of director nodes folded onto the rotations
A2DS_HD double geo_tying_scalar(const double *sig, int p, int pp) {
  const int m = p & 3, mm = pp & 3;
  const bool pd = p >= 4, ppd = pp >= 4;
  const double dN = (m % 2) ? 0.5 : -0.5, dM = (m / 2) ? 0.5 : -0.5;
  const double dNN = (mm % 2) ? 0.5 : -0.5, dMM = (mm / 2) ? 0.5 : -0.5;
  double sc = 0.0;
  if (!pd && !ppd) {
    if (m / 2 == mm / 2) sc += sig[2 * (m / 2)] * dN * dNN;          // g11 on the shared eta edge
    if (m % 2 == mm % 2) sc += sig[4 + 2 * (m % 2)] * dM * dMM;      // g22 on the shared xi edge
    sc += sig[8] * 0.5 * ((0.5 * dN) * (0.5 * dMM) + (0.5 * dM) * (0.5 * dNN));
  } else if (pd != ppd) {
    const int md = pd ? m : mm, mu = pd ? mm : m;
    const double dNu = (mu % 2) ? 0.5 : -0.5, dMu = (mu / 2) ? 0.5 : -0.5;
    if (md % 2 == mu % 2) sc += sig[5 + 2 * (md % 2)] * 0.5 * 0.5 * dMu;  // g23
    if (md / 2 == mu / 2) sc += sig[1 + 2 * (md / 2)] * 0.5 * 0.5 * dNu;  // g13
  }
  return sc;
}
// mq = alpha_p^T Sigma beta_pp + beta_p^T Sigma alpha_pp, Sigma = [[s3, s5], [s5, s4]]
A2DS_HD double geo_bending_mq(const double *ap, const double *bp, const double *app,
                              const double *bpp, const double *sg) {
  const double s3 = sg[0], s4 = sg[1], s5 = sg[2];
  return ap[0] * (s3 * bpp[0] + s5 * bpp[1]) + ap[1] * (s5 * bpp[0] + s4 * bpp[1]) +
         bp[0] * (s3 * app[0] + s5 * app[1]) + bp[1] * (s5 * app[0] + s4 * app[1]);
}
A2DS_HD void geo_fold(const double *fn12, int p, int pp, double blk[9]) {
  const int m = p & 3, mm = pp & 3;
  if (p >= 4) {  // rows: skew(fn_m) * blk
    const double *f = &fn12[3 * m];
    double t[9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      t[j] = f[1] * blk[6 + j] - f[2] * blk[3 + j];
      t[3 + j] = f[2] * blk[j] - f[0] * blk[6 + j];
      t[6 + j] = f[0] * blk[3 + j] - f[1] * blk[j];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) blk[i] = t[i];
  }
  if (pp >= 4) {  // columns: blk * skew(fn_mm)^T
    const double *f = &fn12[3 * mm];
    double t[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double *r = &blk[3 * i];
      t[3 * i] = f[1] * r[2] - f[2] * r[1];
      t[3 * i + 1] = f[2] * r[0] - f[0] * r[2];
      t[3 * i + 2] = f[0] * r[1] - f[1] * r[0];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) blk[i] = t[i];
  }
}
// block from the four bending scalars mq[qp] (however they were formed)
template <class Rec>
A2DS_HD void geo_block_from_mq(const Rec &gm, const TyWork &s, int p, int pp, const double mq[4],
                               double out[9]) {
  const double sc = geo_tying_scalar(s.sigt, p, pp);
  double blk[9] = {sc, 0.0, 0.0, 0.0, sc, 0.0, 0.0, 0.0, sc};
#pragma unroll
  for (int qp = 0; qp < 4; qp++) {
    const double *P = gm.qp[qp].Pq;
    blk[0] += mq[qp] * P[0]; blk[1] += mq[qp] * P[1]; blk[2] += mq[qp] * P[2];
    blk[3] += mq[qp] * P[1]; blk[4] += mq[qp] * P[3]; blk[5] += mq[qp] * P[4];
    blk[6] += mq[qp] * P[2]; blk[7] += mq[qp] * P[4]; blk[8] += mq[qp] * P[5];
  }
  geo_fold(gm.fn, p, pp, blk);
#pragma unroll
  for (int i = 0; i < 9; i++) out[i] = blk[i];
}
template <class Rec>
A2DS_HD void geo_block_t(const Rec &gm, const TyWork &s, int p, int pp, double out[9]) {
  double mq[4];
#pragma unroll
  for (int qp = 0; qp < 4; qp++)
    mq[qp] = geo_bending_mq(s.ca[qp][p], s.cb[qp][p], s.ca[qp][pp], s.cb[qp][pp], gm.qp[qp].sg);
  geo_block_from_mq(gm, s, p, pp, mq, out);
}

}  // namespace a2ds
#endif
