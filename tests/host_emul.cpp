// host_emul.cpp — steps the kernel's per-lane phase functions (mitc4_math.h) lane by
// lane on the CPU so the arithmetic can be checked without a GPU.  Test-only: the
// product never builds or loads this.  The DMMA contraction of the real kernel is
// replaced by plain loops over the SAME per-lane fragments, using the same
// (lane -> Gauss point, node, u/theta) mapping; the scatter is replaced by dense output.
#include <string.h>

#include "../a2d-shells_b200/csrc/mitc4_math.h"

using namespace a2ds;

extern "C" int emul_element(const double *Cs, const double *eth, double temperature, int model,
                            int transform, const double *axis, const double *X, const double *q,
                            int want_gmat, double *res, double *K, double *G) {
  CompData c;
  memcpy(c.Cs, Cs, sizeof(c.Cs));
  memcpy(c.eth, eth, sizeof(c.eth));
  c.temperature = temperature;
  c.model = model;
  c.transform = transform;
  c.coupled = 0;
  for (int k = 6; k < 12; k++)
    if (c.Cs[k] != 0.0) c.coupled = 1;
  memcpy(c.axis, axis, sizeof(c.axis));
  static ElemGeom s;
  static ElemWork wk;
  memset(&s, 0, sizeof(s));
  memset(&wk, 0, sizeof(wk));
  memcpy(s.X, X, sizeof(s.X));
  memcpy(s.q, q, sizeof(s.q));
  Want w;
  w.res = true; w.kmat = true; w.gmat = want_gmat != 0; w.nonlinear = model == 1; w.thermal = 1.0;
  for (int m = 0; m < 4; m++) phase_node(c, s, m);
  static double Pq[4][6];
  for (int qp = 0; qp < 4; qp++)
    phase_qp(c, s, qp, true, w.gmat || w.nonlinear, w.nonlinear, Pq[qp]);
  static double Bc[32][9][3], Wc[32][9][3], Bq[32][9][3];
  memset(Bq, 0, sizeof(Bq));
  for (int lane = 0; lane < 32; lane++) {
    if (w.gmat || w.nonlinear) lane_b1(s, wk, lane, Bq[lane]);
    lane_b0w(c, s, lane, w, Bq[lane], Bc[lane], Wc[lane]);
  }
  memset(res, 0, 24 * sizeof(double));
  for (int lane = 0; lane < 32; lane++) {
    double r3[3];
    lane_stress(c, s, wk, lane, w, Wc[lane], r3);
    const int col = 6 * lane_m(lane) + 3 * lane_h(lane);
    for (int k = 0; k < 3; k++) res[col + k] += r3[k];
  }
  for (int t = 0; t < 9; t++) sum_tying_stress(wk, t);
  // dense operands [dof][k = (strain, qp)] assembled from the per-lane fragments
  static double BA[24][36], W[24][36], B1[24][36];
  for (int lane = 0; lane < 32; lane++) {
    const int qp = lane_qp(lane), col = 6 * lane_m(lane) + 3 * lane_h(lane);
    for (int r = 0; r < 9; r++)
      for (int k = 0; k < 3; k++) {
        BA[col + k][4 * r + qp] = Bc[lane][r][k];
        W[col + k][4 * r + qp] = Wc[lane][r][k];
        B1[col + k][4 * r + qp] = Bq[lane][r][k];
      }
  }
  double geo[576];
  memset(geo, 0, sizeof(geo));
  if (w.gmat || w.nonlinear)
    for (int p = 0; p < 8; p++)
      for (int pp = 0; pp < 8; pp++) {
        double blk[9];
        geo_block(s, wk, &Pq[0][0], p, pp, blk);
        int r0 = 6 * (p & 3) + (p >= 4 ? 3 : 0), c0 = 6 * (pp & 3) + (pp >= 4 ? 3 : 0);
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) geo[24 * (r0 + i) + c0 + j] += blk[3 * i + j];
      }
  for (int r = 0; r < 24; r++)
    for (int cc = 0; cc < 24; cc++) {
      double k = 0.0, gm = 0.0;
      for (int t = 0; t < 36; t++) {
        k += BA[r][t] * W[cc][t];
        gm += B1[r][t] * W[cc][t] + W[r][t] * B1[cc][t];
      }
      if (w.nonlinear) k += geo[24 * r + cc];
      K[24 * r + cc] = k;
      if (G) G[24 * r + cc] = gm + geo[24 * r + cc];
    }
  return 0;
}

// mass matrix of one element through the kernel's mass_block (k_mass in a2ds.cu)
extern "C" int emul_mass(const double *mom, int transform, const double *axis, const double *X,
                         double *M) {
  CompData c;
  memset(&c, 0, sizeof(c));
  memcpy(c.mom, mom, sizeof(c.mom));
  c.transform = transform;
  memcpy(c.axis, axis, sizeof(c.axis));
  static ElemGeom s;
  memset(&s, 0, sizeof(s));
  memcpy(s.X, X, sizeof(s.X));
  for (int m = 0; m < 4; m++) mass_node(s, m);
  for (int qp = 0; qp < 4; qp++) mass_qp(s, qp);
  for (int p = 0; p < 8; p++)
    for (int pp = 0; pp < 8; pp++) {
      double blk[9];
      mass_block(c, s, p, pp, blk);
      int r0 = 6 * (p & 3) + (p >= 4 ? 3 : 0), c0 = 6 * (pp & 3) + (pp >= 4 ? 3 : 0);
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[24 * (r0 + i) + c0 + j] = blk[3 * i + j];
    }
  return 0;
}

// strain matrices B0[qp][strain][dof] (and B1(q) for the state) of one element, assembled from
// the per-lane fragments: lets the tests look at the structure the contraction exploits
extern "C" int emul_strain_matrices(const double *Cs, int transform, const double *axis,
                                    const double *X, const double *q, double *B0, double *B1) {
  CompData c;
  memset(&c, 0, sizeof(c));
  memcpy(c.Cs, Cs, sizeof(c.Cs));
  c.transform = transform;
  memcpy(c.axis, axis, sizeof(c.axis));
  static ElemGeom s;
  static ElemWork wk;
  memset(&s, 0, sizeof(s));
  memset(&wk, 0, sizeof(wk));
  memcpy(s.X, X, sizeof(s.X));
  memcpy(s.q, q, sizeof(s.q));
  Want w;
  w.res = true; w.kmat = true; w.gmat = true; w.nonlinear = false; w.thermal = 1.0;
  for (int m = 0; m < 4; m++) phase_node(c, s, m);
  static double Pq[4][6];
  for (int qp = 0; qp < 4; qp++) phase_qp(c, s, qp, true, true, false, Pq[qp]);
  for (int lane = 0; lane < 32; lane++) {
    double Bc[9][3], Wc[9][3], Bq[9][3];
    lane_b1(s, wk, lane, Bq);
    lane_b0w(c, s, lane, w, Bq, Bc, Wc);
    const int qp = lane_qp(lane), col = 6 * lane_m(lane) + 3 * lane_h(lane);
    for (int r = 0; r < 9; r++)
      for (int k = 0; k < 3; k++) {
        B0[(qp * 9 + r) * 24 + col + k] = Bc[r][k];
        B1[(qp * 9 + r) * 24 + col + k] = Bq[r][k];
      }
  }
  return 0;
}

// ---- tying-level formulation (mitc4_tying.h, kernel k_assemble_t) ----------------------
#include "../a2d-shells_b200/csrc/mitc4_tying.h"

extern "C" int emul_element_t(const double *Cs, const double *eth, double temperature, int model,
                              int transform, const double *axis, const double *X, const double *q,
                              int want_gmat, double *res, double *K, double *G) {
  CompData c;
  memcpy(c.Cs, Cs, sizeof(c.Cs));
  memcpy(c.eth, eth, sizeof(c.eth));
  c.temperature = temperature;
  c.model = model;
  c.transform = transform;
  c.coupled = 0;
  for (int k = 6; k < 12; k++)
    if (c.Cs[k] != 0.0) return 1;  // this formulation is for uncoupled sections
  memcpy(c.axis, axis, sizeof(c.axis));
  static ElemRec s;
  static TyWork wk;
  memset(&s, 0, sizeof(s));
  memset(&wk, 0, sizeof(wk));
  memcpy(s.X, X, sizeof(s.X));
  memcpy(s.q, q, sizeof(s.q));
  Want w;
  w.res = true; w.kmat = true; w.gmat = want_gmat != 0; w.nonlinear = model == 1; w.thermal = 1.0;
  const bool need_state = w.gmat || w.nonlinear;
  for (int m = 0; m < 4; m++) phase_node(c, s, m);
  for (int m = 0; m < 4; m++) phase_node_tab(s, m, need_state);
  for (int qp = 0; qp < 4; qp++) phase_qp_t(c, s, s.qp[qp], qp, w);
  for (int e = 0; e < 45; e++) {
    TyPlan pl;
    ty_plan(e, pl);
    ty_H_entry(s, pl, wk.H);
  }
  for (int t = 0; t < 9; t++) ty_sum_stress(s, wk, t);
  static LaneFrag f[32];
  static double B1[32][6][3];
  memset(B1, 0, sizeof(B1));
  memset(res, 0, 24 * sizeof(double));
  for (int lane = 0; lane < 32; lane++) {
    LaneConst lc;
    lane_const(lane, lc);
    lane_fragments(s, wk, lane, lc, w, f[lane], B1[lane]);
    double r3[3];
    lane_residual(s, wk, lane, f[lane], r3);
    const int col = 6 * lane_m(lane) + 3 * lane_h(lane);
    for (int k = 0; k < 3; k++) res[col + k] += r3[k];
  }
  // dense operands [dof][k = 4 s + kk]
  static double BA[24][28], W[24][28], BB[24][28];
  memset(BB, 0, sizeof(BB));
  for (int lane = 0; lane < 32; lane++) {
    const int kk = lane_qp(lane), col = 6 * lane_m(lane) + 3 * lane_h(lane);
    for (int r = 0; r < 7; r++)
      for (int k = 0; k < 3; k++) {
        BA[col + k][4 * r + kk] = f[lane].B[r][k];
        W[col + k][4 * r + kk] = f[lane].W[r][k];
        if (r < 6) BB[col + k][4 * r + kk] = B1[lane][r][k];
      }
  }
  double geo[576];
  memset(geo, 0, sizeof(geo));
  if (need_state)
    for (int p = 0; p < 8; p++)
      for (int pp = 0; pp < 8; pp++) {
        double blk[9];
        geo_block_t(s, wk, p, pp, blk);
        int r0 = 6 * (p & 3) + (p >= 4 ? 3 : 0), c0 = 6 * (pp & 3) + (pp >= 4 ? 3 : 0);
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) geo[24 * (r0 + i) + c0 + j] += blk[3 * i + j];
      }
  for (int r = 0; r < 24; r++)
    for (int cc = 0; cc < 24; cc++) {
      double k = 0.0, gm = 0.0;
      for (int t = 0; t < 28; t++) {
        k += BA[r][t] * W[cc][t];
        gm += BB[r][t] * W[cc][t] + W[r][t] * BB[cc][t];
      }
      if (w.nonlinear) k += geo[24 * r + cc];
      K[24 * r + cc] = k;
      if (G) G[24 * r + cc] = gm + geo[24 * r + cc];
    }
  return 0;
}

// ---- 9-node element (mitc9_math.h): the work items of k_assemble9 stepped in the kernel's
// phase order; the DMMA contraction replaced by plain loops over the same B / CB tables ----
#include "../a2d-shells_b200/csrc/mitc9_math.h"

extern "C" int emul_element9(const double *Cs, const double *eth, double temperature, int transform,
                             const double *axis, const double *X, const double *q, double *res,
                             double *K, int model, double *G) {
  CompData c;
  memset(&c, 0, sizeof(c));
  memcpy(c.Cs, Cs, sizeof(c.Cs));
  memcpy(c.eth, eth, sizeof(c.eth));
  c.temperature = temperature;
  c.transform = transform;
  c.model = model;
  memcpy(c.axis, axis, sizeof(c.axis));
  const bool nl = model == 1;
  static Elem9 E;
  static Tab9 Tb;
  memset(&E, 0, sizeof(E));
  memset(&Tb, 0, sizeof(Tb));
  memcpy(E.X, X, sizeof(E.X));
  memcpy(E.q, q, sizeof(E.q));
  static Shape9 H;
  for (int pnt = 0; pnt < 46; pnt++) q9_shape_tables(H, pnt);
  for (int n = 0; n < 9; n++) q9_node(c, E, n);
  for (int t = 0; t < 28; t++) q9_tying(E, H, t, nl);
  for (int qp = 0; qp < 9; qp++) q9_qp(c, E, H, qp);
  for (int qp = 0; qp < 9; qp++) q9_qp_state(c, E, H, qp, 1.0, nl);
  for (int t = 0; t < 28; t++) q9_sigt(E, H, t);
  for (int t = 0; t < 28; t++)
    for (int d = 0; d < 54; d++) { Tb.Gt[t][d] = q9_gt(E, H, t, d); Tb.Gt1[t][d] = q9_gt1(E, H, t, d); }
  for (int n = 0; n < 9; n++)
    for (int d = 0; d < 54; d++) Tb.Dn[n][d] = q9_dn(E, H, n, d);
  memset(res, 0, 54 * sizeof(double));
  memset(K, 0, 54 * 54 * sizeof(double));
  if (G) memset(G, 0, 54 * 54 * sizeof(double));
  static double B1[12][56], CB1[12][56];
  for (int qp = 0; qp < 9; qp++) {
    for (int d = 0; d < 54; d++) {
      double Bk[9], B1k[9], Sk[9];
      q9_bcol(E, Tb, H, qp, d, Bk, B1k);
      if (nl)
        for (int k = 0; k < 9; k++) Bk[k] += B1k[k];
      q9_stress(c.Cs, Bk, Sk);
      for (int k = 0; k < 9; k++) { Tb.B[k][d] = Bk[k]; Tb.CB[k][d] = E.qw[qp] * Sk[k]; }
      q9_stress(c.Cs, B1k, Sk);
      for (int k = 0; k < 9; k++) { B1[k][d] = B1k[k]; CB1[k][d] = E.qw[qp] * Sk[k]; }
    }
    for (int a = 0; a < 54; a++) {
      double r = 0.0;
      for (int k = 0; k < 9; k++) r += Tb.B[k][a] * E.sq[qp][k];
      res[a] += r;
      for (int b = 0; b < 54; b++) {
        double s = 0.0, z = 0.0;
        for (int k = 0; k < 9; k++) {
          s += Tb.B[k][a] * Tb.CB[k][b];
          z += Tb.B[k][a] * CB1[k][b] + CB1[k][a] * Tb.B[k][b];
        }
        K[54 * a + b] += s;
        if (G) G[54 * a + b] += z;
      }
    }
  }
  // geometric term of the node pairs: the whole of it for G (stresses of the linear model),
  // added to the tangent of the nonlinear model
  if (G || nl)
    for (int ma = 0; ma < 9; ma++)
      for (int mb = 0; mb < 9; mb++) {
        double blk[36];
        q9_geo_pair(E, H, ma, mb, blk);
        for (int i = 0; i < 6; i++)
          for (int j = 0; j < 6; j++) {
            if (G) G[54 * (6 * ma + i) + 6 * mb + j] += blk[6 * i + j];
            if (nl) K[54 * (6 * ma + i) + 6 * mb + j] += blk[6 * i + j];
          }
      }
  return 0;
}

// mass matrix of the 9-node element through q9_mass_pair (the work items of k_mass9)
extern "C" int emul_mass9(const double *mom, const double *X, double *M) {
  CompData c;
  memset(&c, 0, sizeof(c));
  memcpy(c.mom, mom, sizeof(c.mom));
  static Elem9 E;
  static Shape9 H;
  memset(&E, 0, sizeof(E));
  memcpy(E.X, X, sizeof(E.X));
  for (int pnt = 0; pnt < 46; pnt++) q9_shape_tables(H, pnt);
  for (int n = 0; n < 9; n++) q9_node_normal(E, n);
  for (int q = 0; q < 9; q++) q9_qp_det(E, H, q);
  for (int ma = 0; ma < 9; ma++)
    for (int mb = 0; mb < 9; mb++) {
      double blk[36];
      q9_mass_pair(c, E, H, ma, mb, blk);
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) M[54 * (6 * ma + i) + 6 * mb + j] = blk[6 * i + j];
    }
  return 0;
}
