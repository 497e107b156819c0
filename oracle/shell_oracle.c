/*
 * shell_oracle.c — plain-C restatement of the reference's MITC4 shell element
 * and its single-rank assembly.  TEST INFRASTRUCTURE ONLY (see shell_oracle.h).
 *
 * The forward strain evaluation follows the reference step by step (each
 * function cites the reference lines it restates).  Residual and tangent are
 * then obtained by exact differentiation of the strain energy
 *     U(q) = 1/2 * sum_qp detXd * (e - T*eth)^T C (e - T*eth),
 * which is what the reference's hand-derived forward/reverse code computes
 * (TACSShellElement.h:303-672):  every strain is a polynomial of degree <= 2 in
 * the element state, e(q) = Lin(q) + 1/2 Bil(q,q), so
 *     de/dq_a      = Lin(1_a) + Bil(q, 1_a)
 *     d2e/dq_a dq_b = Bil(1_a, 1_b)
 * are evaluated from the reference's own directional-derivative recipes
 * (computeTyingStrainDeriv, TacsShellComputeDispGradDeriv, evalStrainDeriv,
 * TacsShellComputeDrillStrainDeriv).  The geometric stiffness is the
 * reference's central difference of the nonlinear tangent, restated literally
 * (TACSShellElement.h:705-751).
 *
 * Mass/inertial terms (beta, gamma) are outside the static hot path and are not
 * restated.
 */
#include "shell_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* 2-point Gauss abscissa, 15-digit literal: basis/TACSGaussQuadrature.h:26 */
static const double GAUSS_PT = 0.577350269189626;

/* ---- small algebra, TACSElementAlgebra.h ------------------------------- */
static void cross3(const double x[3], const double y[3], double o[3]) { /* :39 */
  o[0] = x[1] * y[2] - x[2] * y[1];
  o[1] = x[2] * y[0] - x[0] * y[2];
  o[2] = x[0] * y[1] - x[1] * y[0];
}
static double dot3(const double x[3], const double y[3]) { /* :139 */
  return x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
}
static void matmul(const double A[9], const double B[9], double C[9]) { /* :825 */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void matmul_add(const double A[9], const double B[9], double C[9]) { /* :957 */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[3 * i + j] += A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void trans_matmul(const double A[9], const double B[9], double C[9]) { /* :913 */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
static double inv3(const double A[9], double Ai[9]) { /* inv3x3 :1980 */
  double det = (A[8] * (A[0] * A[4] - A[3] * A[1]) - A[7] * (A[0] * A[5] - A[3] * A[2]) +
                A[6] * (A[1] * A[5] - A[2] * A[4]));
  double di = 1.0 / det;
  Ai[0] = (A[4] * A[8] - A[5] * A[7]) * di;
  Ai[1] = -(A[1] * A[8] - A[2] * A[7]) * di;
  Ai[2] = (A[1] * A[5] - A[2] * A[4]) * di;
  Ai[3] = -(A[3] * A[8] - A[5] * A[6]) * di;
  Ai[4] = (A[0] * A[8] - A[2] * A[6]) * di;
  Ai[5] = -(A[0] * A[5] - A[2] * A[3]) * di;
  Ai[6] = (A[3] * A[7] - A[4] * A[6]) * di;
  Ai[7] = -(A[0] * A[7] - A[1] * A[6]) * di;
  Ai[8] = (A[0] * A[4] - A[1] * A[3]) * di;
  return det;
}
/* A = T^T S T for symmetric S[6], mat3x3SymmTransformTranspose :1094 */
static void symm_transform_t(const double T[9], const double S[6], double A[6]) {
  double W[9];
  W[0] = S[0] * T[0] + S[1] * T[3] + S[2] * T[6];
  W[1] = S[0] * T[1] + S[1] * T[4] + S[2] * T[7];
  W[2] = S[0] * T[2] + S[1] * T[5] + S[2] * T[8];
  W[3] = S[1] * T[0] + S[3] * T[3] + S[4] * T[6];
  W[4] = S[1] * T[1] + S[3] * T[4] + S[4] * T[7];
  W[5] = S[1] * T[2] + S[3] * T[5] + S[4] * T[8];
  W[6] = S[2] * T[0] + S[4] * T[3] + S[5] * T[6];
  W[7] = S[2] * T[1] + S[4] * T[4] + S[5] * T[7];
  W[8] = S[2] * T[2] + S[4] * T[5] + S[5] * T[8];
  A[0] = T[0] * W[0] + T[3] * W[3] + T[6] * W[6];
  A[1] = T[0] * W[1] + T[3] * W[4] + T[6] * W[7];
  A[2] = T[0] * W[2] + T[3] * W[5] + T[6] * W[8];
  A[3] = T[1] * W[1] + T[4] * W[4] + T[7] * W[7];
  A[4] = T[1] * W[2] + T[4] * W[5] + T[7] * W[8];
  A[5] = T[2] * W[2] + T[5] * W[5] + T[8] * W[8];
}

/* ---- basis, TACSShellElementQuadBasis.h (order 2) ----------------------- */
typedef struct { double N[4], Nxi[4], Neta[4]; } shape_t;

static void shape_eval(const double pt[2], shape_t *s) { /* :62-114, :171-232 */
  double na[2] = {0.5 * (1.0 - pt[0]), 0.5 * (1.0 + pt[0])};
  double nb[2] = {0.5 * (1.0 - pt[1]), 0.5 * (1.0 + pt[1])};
  const double dna[2] = {-0.5, 0.5}, dnb[2] = {-0.5, 0.5};
  for (int j = 0; j < 2; j++)
    for (int i = 0; i < 2; i++) {
      s->N[2 * j + i] = na[i] * nb[j];
      s->Nxi[2 * j + i] = dna[i] * nb[j];
      s->Neta[2 * j + i] = na[i] * dnb[j];
    }
}
/* interpFields<stride,3> :171 */
static void interp3(const shape_t *s, const double *v, int stride, double f[3]) {
  f[0] = f[1] = f[2] = 0.0;
  for (int n = 0; n < 4; n++)
    for (int k = 0; k < 3; k++) f[k] += s->N[n] * v[stride * n + k];
}
/* interpFieldsGrad<stride,3> :210 — grad[2k] = d/dxi, grad[2k+1] = d/deta */
static void interp3_grad(const shape_t *s, const double *v, int stride, double g[6]) {
  for (int k = 0; k < 6; k++) g[k] = 0.0;
  for (int n = 0; n < 4; n++)
    for (int k = 0; k < 3; k++) {
      g[2 * k] += s->Nxi[n] * v[stride * n + k];
      g[2 * k + 1] += s->Neta[n] * v[stride * n + k];
    }
}
static void node_point(int n, double pt[2]) { /* getNodePoint :147 */
  pt[0] = -1.0 + 2.0 * (n % 2);
  pt[1] = -1.0 + 2.0 * (n / 2);
}
/* tying points, getTyingPoint :530-564 with knots {-1,1} / {0}; field per index
   getTyingField :487: 0,1 g11; 2,3 g22; 4 g12; 5,6 g23; 7,8 g13 */
static const double TY_PT[9][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}, {0, 0},
                                   {-1, 0}, {1, 0}, {0, -1}, {0, 1}};
static const int TY_FIELD[9] = {0, 0, 1, 1, 2, 3, 3, 4, 4};
/* interpTyingStrain :651-672 (evalTyingInterp :569-616): gty = [g11 g12 g13 g22 g23 0] */
static void interp_tying(const double pt[2], const double ety[9], double gty[6]) {
  double na[2] = {0.5 * (1.0 - pt[0]), 0.5 * (1.0 + pt[0])};
  double nb[2] = {0.5 * (1.0 - pt[1]), 0.5 * (1.0 + pt[1])};
  gty[0] = 0.0; gty[0] += (1.0 * nb[0]) * ety[0]; gty[0] += (1.0 * nb[1]) * ety[1];
  gty[3] = 0.0; gty[3] += (na[0] * 1.0) * ety[2]; gty[3] += (na[1] * 1.0) * ety[3];
  gty[1] = 0.0; gty[1] += (1.0 * 1.0) * ety[4];
  gty[4] = 0.0; gty[4] += (na[0] * 1.0) * ety[5]; gty[4] += (na[1] * 1.0) * ety[6];
  gty[2] = 0.0; gty[2] += (1.0 * nb[0]) * ety[7]; gty[2] += (1.0 * nb[1]) * ety[8];
  gty[5] = 0.0;
}

/* ---- transforms, TACSShellElementTransform.h ---------------------------- */
static void compute_transform(const oracle_comp_t *c, const double Xxi[6], const double n0[3],
                              double T[9]) {
  double n[3] = {n0[0], n0[1], n0[2]};
  double inv = 1.0 / sqrt(dot3(n, n));
  n[0] *= inv; n[1] *= inv; n[2] *= inv;
  double t1[3], t2[3];
  if (c->transform == 0) {
    /* natural transform :25-92, including the t1[0]-only projection at :42-44 */
    t1[0] = Xxi[0]; t1[1] = Xxi[2]; t1[2] = Xxi[4];
    double d = dot3(n, t1);
    t1[0] = t1[0] - d * n[0];
    t1[0] = t1[0] - d * n[0];
    t1[0] = t1[0] - d * n[0];
  } else {
    /* reference-axis transform :116-213 */
    double an = dot3(c->axis, n);
    t1[0] = c->axis[0] - an * n[0];
    t1[1] = c->axis[1] - an * n[1];
    t1[2] = c->axis[2] - an * n[2];
  }
  inv = 1.0 / sqrt(dot3(t1, t1));
  t1[0] *= inv; t1[1] *= inv; t1[2] *= inv;
  cross3(n, t1, t2);
  T[0] = t1[0]; T[3] = t1[1]; T[6] = t1[2];
  T[1] = t2[0]; T[4] = t2[1]; T[7] = t2[2];
  T[2] = n[0]; T[5] = n[1]; T[8] = n[2];
}

/* ---- element geometry (state independent) ------------------------------- */
typedef struct {
  double fn[12], Xdn[36], Tn[36], XdinvTn[36];
  shape_t sn[4];                     /* shape functions at the nodes */
  shape_t st[9];                     /* ... at the tying points */
  double Xxi_t[9][6], n0_t[9][3];
  shape_t sq[4];                     /* ... at the Gauss points */
  double pt[4][2], T[4][9], XdinvT[4][9], XdinvzT[4][9], detXd[4];
} geo_t;

/* frame [a|b|c] with the vectors in columns, TacsShellAssembleFrame (TACSShellUtilities.h:8-35) */
static void frame_xn(const double Xxi[6], const double n[3], double Xd[9]) {
  Xd[0] = Xxi[0]; Xd[1] = Xxi[1]; Xd[2] = n[0];
  Xd[3] = Xxi[2]; Xd[4] = Xxi[3]; Xd[5] = n[1];
  Xd[6] = Xxi[4]; Xd[7] = Xxi[5]; Xd[8] = n[2];
}
static void frame_x0(const double nxi[6], double Xdz[9]) {
  Xdz[0] = nxi[0]; Xdz[1] = nxi[1]; Xdz[2] = 0.0;
  Xdz[3] = nxi[2]; Xdz[4] = nxi[3]; Xdz[5] = 0.0;
  Xdz[6] = nxi[4]; Xdz[7] = nxi[5]; Xdz[8] = 0.0;
}

static void geometry(const oracle_comp_t *c, const double X[12], geo_t *g) {
  /* TacsShellComputeNodeNormals, TACSShellUtilities.h:301-342 */
  for (int i = 0; i < 4; i++) {
    double pt[2];
    node_point(i, pt);
    shape_eval(pt, &g->sn[i]);
    double Xxi[6];
    interp3_grad(&g->sn[i], X, 3, Xxi);
    double a[3] = {Xxi[0], Xxi[2], Xxi[4]}, b[3] = {Xxi[1], Xxi[3], Xxi[5]};
    cross3(a, b, &g->fn[3 * i]);
    double norm = sqrt(dot3(&g->fn[3 * i], &g->fn[3 * i]));
    if (norm != 0.0) {
      double s = 1.0 / norm;
      g->fn[3 * i] *= s; g->fn[3 * i + 1] *= s; g->fn[3 * i + 2] *= s;
    }
    frame_xn(Xxi, &g->fn[3 * i], &g->Xdn[9 * i]);
    /* node transform and Xdinv*T, TacsShellComputeDrillStrain :656-674 */
    compute_transform(c, Xxi, &g->fn[3 * i], &g->Tn[9 * i]);
    double Xdinv[9];
    inv3(&g->Xdn[9 * i], Xdinv);
    matmul(Xdinv, &g->Tn[9 * i], &g->XdinvTn[9 * i]);
  }
  /* tying-point frames, TACSShellElementModel.h:36-47,62-64 */
  for (int t = 0; t < 9; t++) {
    shape_eval(TY_PT[t], &g->st[t]);
    interp3_grad(&g->st[t], X, 3, g->Xxi_t[t]);
    interp3(&g->st[t], g->fn, 3, g->n0_t[t]);
  }
  /* Gauss points, TACSShellElement.h:514-534 and TacsShellComputeDispGrad
     (TACSShellUtilities.h:369-393); quadrature order xi fastest
     (TACSShellElementQuadrature.h:22-27), weight 1 */
  for (int q = 0; q < 4; q++) {
    g->pt[q][0] = (q % 2 == 0) ? -GAUSS_PT : GAUSS_PT;
    g->pt[q][1] = (q / 2 == 0) ? -GAUSS_PT : GAUSS_PT;
    shape_eval(g->pt[q], &g->sq[q]);
    double Xxi[6], n0[3], nxi[6];
    interp3_grad(&g->sq[q], X, 3, Xxi);
    interp3(&g->sq[q], g->fn, 3, n0);
    compute_transform(c, Xxi, n0, g->T[q]);
    interp3_grad(&g->sq[q], g->fn, 3, nxi);
    double Xd[9], Xdz[9], Xdinv[9], neg[9];
    frame_xn(Xxi, n0, Xd);
    frame_x0(nxi, Xdz);
    g->detXd[q] = inv3(Xd, Xdinv) * 1.0;
    matmul(Xdinv, Xdz, neg);
    for (int k = 0; k < 9; k++) neg[k] *= -1.0;
    matmul(Xdinv, g->T[q], g->XdinvT[q]);
    matmul(neg, g->XdinvT[q], g->XdinvzT[q]);
  }
}

/* ---- quantities that are LINEAR in the element state ---------------------- */
typedef struct {
  double d[12];                       /* director d = q x fn, TACSDirector.h:244-267 */
  double Uxi_t[9][6], d0_t[9][3];     /* at the tying points */
  double u0x[4][9], u1x[4][9];        /* at the Gauss points */
  double u0xn[4][9], Ctn_lin[4][9];   /* at the nodes: u0x and T^T(-q^x)T */
} lin_t;

static void linear_maps(const geo_t *g, const double v[24], lin_t *l) {
  for (int n = 0; n < 4; n++) cross3(&v[6 * n + 3], &g->fn[3 * n], &l->d[3 * n]);
  for (int t = 0; t < 9; t++) {
    interp3_grad(&g->st[t], v, 6, l->Uxi_t[t]);
    interp3(&g->st[t], l->d, 3, l->d0_t[t]);
  }
  /* TacsShellComputeDispGrad, TACSShellUtilities.h:395-418 */
  for (int q = 0; q < 4; q++) {
    double d0[3], d0xi[6], u0xi[6], u0d[9], u1d[9], tmp[9];
    interp3(&g->sq[q], l->d, 3, d0);
    interp3_grad(&g->sq[q], l->d, 3, d0xi);
    interp3_grad(&g->sq[q], v, 6, u0xi);
    frame_xn(u0xi, d0, u0d);
    frame_x0(d0xi, u1d);
    matmul(u1d, g->XdinvT[q], tmp);
    matmul_add(u0d, g->XdinvzT[q], tmp);
    trans_matmul(g->T[q], tmp, l->u1x[q]);
    matmul(u0d, g->XdinvT[q], tmp);
    trans_matmul(g->T[q], tmp, l->u0x[q]);
  }
  /* TacsShellComputeDrillStrain(Deriv), TACSShellUtilities.h:665-691, 740-759 */
  for (int n = 0; n < 4; n++) {
    double u0xi[6], u0d[9], tmp[9], Cd[9];
    interp3_grad(&g->sn[n], v, 6, u0xi);
    frame_x0(u0xi, u0d);
    matmul(u0d, &g->XdinvTn[9 * n], tmp);
    trans_matmul(&g->Tn[9 * n], tmp, l->u0xn[n]);
    /* Cd = -qd^x, setMatSkew(-1, q, C): TACSElementAlgebra.h:1518 */
    const double *qd = &v[6 * n + 3];
    Cd[0] = 0.0; Cd[1] = qd[2]; Cd[2] = -qd[1];
    Cd[3] = -qd[2]; Cd[4] = 0.0; Cd[5] = qd[0];
    Cd[6] = qd[1]; Cd[7] = -qd[0]; Cd[8] = 0.0;
    trans_matmul(&g->Tn[9 * n], Cd, tmp);
    matmul(tmp, &g->Tn[9 * n], l->Ctn_lin[n]);
  }
}

/* Linear part of the strains at Gauss point q for the linear maps l:
   TACSShellLinearModel::computeTyingStrain(Deriv) (TACSShellElementModel.h:33-77),
   interpTyingStrain, mat3x3SymmTransformTranspose, evalStrain (:440-455),
   drill strain evalDrillStrainDeriv (TACSDirector.h:590-599). */
static void tying_linear(const geo_t *g, const lin_t *l, double ety[9]) {
  for (int t = 0; t < 9; t++) {
    const double *Uxi = l->Uxi_t[t], *Xxi = g->Xxi_t[t], *d0 = l->d0_t[t], *n0 = g->n0_t[t];
    switch (TY_FIELD[t]) {
      case 0: ety[t] = (Uxi[0] * Xxi[0] + Uxi[2] * Xxi[2] + Uxi[4] * Xxi[4]); break;
      case 1: ety[t] = (Uxi[1] * Xxi[1] + Uxi[3] * Xxi[3] + Uxi[5] * Xxi[5]); break;
      case 2:
        ety[t] = 0.5 * (Uxi[0] * Xxi[1] + Uxi[2] * Xxi[3] + Uxi[4] * Xxi[5] + Uxi[1] * Xxi[0] +
                        Uxi[3] * Xxi[2] + Uxi[5] * Xxi[4]);
        break;
      case 3:
        ety[t] = 0.5 * (Xxi[1] * d0[0] + Xxi[3] * d0[1] + Xxi[5] * d0[2] + n0[0] * Uxi[1] +
                        n0[1] * Uxi[3] + n0[2] * Uxi[5]);
        break;
      default:
        ety[t] = 0.5 * (Xxi[0] * d0[0] + Xxi[2] * d0[1] + Xxi[4] * d0[2] + n0[0] * Uxi[0] +
                        n0[1] * Uxi[2] + n0[2] * Uxi[4]);
    }
  }
}
/* Bilinear (polar) part of the nonlinear tying strains,
   TACSShellNonlinearModel::computeTyingStrainDeriv (TACSShellElementModel.h:1033-1109) */
static void tying_bilinear(const lin_t *a, const lin_t *b, double ety[9]) {
  for (int t = 0; t < 9; t++) {
    const double *Ua = a->Uxi_t[t], *Ub = b->Uxi_t[t], *da = a->d0_t[t], *db = b->d0_t[t];
    switch (TY_FIELD[t]) {
      case 0: ety[t] = Ua[0] * Ub[0] + Ua[2] * Ub[2] + Ua[4] * Ub[4]; break;
      case 1: ety[t] = Ua[1] * Ub[1] + Ua[3] * Ub[3] + Ua[5] * Ub[5]; break;
      case 2:
        ety[t] = 0.5 * (Ub[0] * Ua[1] + Ub[2] * Ua[3] + Ub[4] * Ua[5] + Ua[0] * Ub[1] +
                        Ua[2] * Ub[3] + Ua[4] * Ub[5]);
        break;
      case 3:
        ety[t] = 0.5 * (da[0] * Ub[1] + db[0] * Ua[1] + da[1] * Ub[3] + db[1] * Ua[3] +
                        da[2] * Ub[5] + db[2] * Ua[5]);
        break;
      default:
        ety[t] = 0.5 * (da[0] * Ub[0] + db[0] * Ua[0] + da[1] * Ub[2] + db[1] * Ua[2] +
                        da[2] * Ub[4] + db[2] * Ua[4]);
    }
  }
}
/* tying strains -> e[0,1,2,6,7] at Gauss point q */
static void membrane_shear(const geo_t *g, int q, const double ety[9], double e[9]) {
  double gty[6], e0ty[6];
  interp_tying(g->pt[q], ety, gty);
  symm_transform_t(g->XdinvT[q], gty, e0ty);
  e[0] = e0ty[0];
  e[1] = e0ty[3];
  e[2] = 2.0 * e0ty[1];
  e[6] = 2.0 * e0ty[4];
  e[7] = 2.0 * e0ty[2];
}
static void strain_linear(const geo_t *g, const lin_t *l, const double ety[9], int q,
                          double e[9]) {
  membrane_shear(g, q, ety, e);
  const double *u1x = l->u1x[q];
  e[3] = u1x[0];
  e[4] = u1x[4];
  e[5] = u1x[1] + u1x[3];
  double et = 0.0;
  for (int n = 0; n < 4; n++) {
    double etn = 0.5 * (l->Ctn_lin[n][3] + l->u0xn[n][3] - l->Ctn_lin[n][1] - l->u0xn[n][1]);
    et += g->sq[q].N[n] * etn;
  }
  e[8] = et;
}
/* bilinear part: TACSShellNonlinearModel::evalStrainDeriv (TACSShellElementModel.h:1172-1211) */
static void strain_bilinear(const geo_t *g, const lin_t *a, const lin_t *b,
                            const double ety_ab[9], int q, double e[9]) {
  membrane_shear(g, q, ety_ab, e);
  const double *u0a = a->u0x[q], *u1a = a->u1x[q], *u0b = b->u0x[q], *u1b = b->u1x[q];
  e[3] = (u0b[0] * u1a[0] + u0b[3] * u1a[3] + u0b[6] * u1a[6] + u0a[0] * u1b[0] +
          u0a[3] * u1b[3] + u0a[6] * u1b[6]);
  e[4] = (u0b[1] * u1a[1] + u0b[4] * u1a[4] + u0b[7] * u1a[7] + u0a[1] * u1b[1] +
          u0a[4] * u1b[4] + u0a[7] * u1b[7]);
  e[5] = (u0b[0] * u1a[1] + u0b[3] * u1a[4] + u0b[6] * u1a[7] + u1b[0] * u0a[1] +
          u1b[3] * u0a[4] + u1b[6] * u0a[7] + u0a[0] * u1b[1] + u0a[3] * u1b[4] +
          u0a[6] * u1b[7] + u1a[0] * u0b[1] + u1a[3] * u0b[4] + u1a[6] * u0b[7]);
  e[8] = 0.0;
}

/* ---- forward strain evaluation exactly as the reference orders it ----------
   TACSShellElement::addResidual, TACSShellElement.h:314-373 */
static void forward_strain(const oracle_comp_t *c, const geo_t *g, const double q[24],
                           double e_out[36]) {
  /* drill strain at the nodes, TacsShellComputeDrillStrain (TACSShellUtilities.h:651-693) */
  double etn[4];
  for (int i = 0; i < 4; i++) {
    double u0xi[6], u0d[9], C[9], tmp[9], Ct[9], u0x[9];
    interp3_grad(&g->sn[i], q, 6, u0xi);
    frame_x0(u0xi, u0d);
    const double *th = &q[6 * i + 3];
    /* C = I - q^x, TACSDirector.h:25-35 */
    C[0] = 1.0; C[1] = th[2]; C[2] = -th[1];
    C[3] = -th[2]; C[4] = 1.0; C[5] = th[0];
    C[6] = th[1]; C[7] = -th[0]; C[8] = 1.0;
    trans_matmul(&g->Tn[9 * i], C, tmp);
    matmul(tmp, &g->Tn[9 * i], Ct);
    matmul(u0d, &g->XdinvTn[9 * i], tmp);
    trans_matmul(&g->Tn[9 * i], tmp, u0x);
    etn[i] = 0.5 * (Ct[3] + u0x[3] - Ct[1] - u0x[1]); /* TACSDirector.h:560-564 */
  }
  lin_t l;
  linear_maps(g, q, &l);
  /* tying strain, computeTyingStrain (TACSShellElementModel.h:33 / :644) */
  double ety[9];
  tying_linear(g, &l, ety);
  if (c->model == 1) {
    double eb[9];
    tying_bilinear(&l, &l, eb);
    for (int t = 0; t < 9; t++) ety[t] += 0.5 * eb[t];
  }
  for (int qp = 0; qp < 4; qp++) {
    double *e = &e_out[9 * qp];
    membrane_shear(g, qp, ety, e);
    const double *u0x = l.u0x[qp], *u1x = l.u1x[qp];
    e[3] = u1x[0];
    e[4] = u1x[4];
    e[5] = u1x[1] + u1x[3];
    if (c->model == 1) { /* TACSShellElementModel.h:1123-1127 */
      e[3] = u1x[0] + (u0x[0] * u1x[0] + u0x[3] * u1x[3] + u0x[6] * u1x[6]);
      e[4] = u1x[4] + (u0x[1] * u1x[1] + u0x[4] * u1x[4] + u0x[7] * u1x[7]);
      e[5] = u1x[1] + u1x[3] +
             (u0x[0] * u1x[1] + u0x[3] * u1x[4] + u0x[6] * u1x[7] + u1x[0] * u0x[1] +
              u1x[3] * u0x[4] + u1x[6] * u0x[7]);
    }
    double et = 0.0;
    for (int n = 0; n < 4; n++) et += g->sq[qp].N[n] * etn[n];
    e[8] = et;
  }
}

/* TACSShellConstitutive::computeStress, TACSShellConstitutive.h:125-147 */
static void stress(const double Cs[22], const double e[9], double s[9]) {
  const double *A = &Cs[0], *B = &Cs[6], *D = &Cs[12], *As = &Cs[18];
  s[0] = A[0] * e[0] + A[1] * e[1] + A[2] * e[2] + B[0] * e[3] + B[1] * e[4] + B[2] * e[5];
  s[1] = A[1] * e[0] + A[3] * e[1] + A[4] * e[2] + B[1] * e[3] + B[3] * e[4] + B[4] * e[5];
  s[2] = A[2] * e[0] + A[4] * e[1] + A[5] * e[2] + B[2] * e[3] + B[4] * e[4] + B[5] * e[5];
  s[3] = B[0] * e[0] + B[1] * e[1] + B[2] * e[2] + D[0] * e[3] + D[1] * e[4] + D[2] * e[5];
  s[4] = B[1] * e[0] + B[3] * e[1] + B[4] * e[2] + D[1] * e[3] + D[3] * e[4] + D[4] * e[5];
  s[5] = B[2] * e[0] + B[4] * e[1] + B[5] * e[2] + D[2] * e[3] + D[4] * e[4] + D[5] * e[5];
  s[6] = As[0] * e[6] + As[1] * e[7];
  s[7] = As[1] * e[6] + As[2] * e[7];
  s[8] = Cs[21] * e[8];
}

void oracle_strain(const oracle_comp_t *c, const double X[12], const double q[24],
                   double e[36], double detXd[4]) {
  geo_t g;
  geometry(c, X, &g);
  forward_strain(c, &g, q, e);
  for (int i = 0; i < 4; i++) detXd[i] = g.detXd[i];
}

/* residual and (optionally) tangent by exact differentiation of the energy */
static void res_and_tangent(const oracle_comp_t *c, double alpha, double temperature,
                            const double X[12], const double q[24], double res[24],
                            double mat[576]) {
  geo_t g;
  geometry(c, X, &g);
  double e[36];
  forward_strain(c, &g, q, e);

  /* stresses of the mechanical strain, TACSShellElement.h:549-574 */
  double s[4][9];
  for (int qp = 0; qp < 4; qp++) {
    double em[9];
    for (int i = 0; i < 9; i++) em[i] = e[9 * qp + i] - c->eth[i] * temperature;
    stress(c->Cs, em, s[qp]);
  }

  /* linear maps of the 24 unit directions and of the state */
  static const int NDOF = 24;
  lin_t *lu = (lin_t *)malloc(sizeof(lin_t) * 25);
  lin_t *lq = &lu[24];
  double ety_lin[24][9];
  for (int a = 0; a < NDOF; a++) {
    double v[24];
    memset(v, 0, sizeof(v));
    v[a] = 1.0;
    linear_maps(&g, v, &lu[a]);
    tying_linear(&g, &lu[a], ety_lin[a]);
  }
  linear_maps(&g, q, lq);

  /* B[qp][a][i] = de_i/dq_a at the state */
  double B[4][24][9];
  for (int a = 0; a < NDOF; a++) {
    double etb[9];
    if (c->model == 1) tying_bilinear(lq, &lu[a], etb);
    for (int qp = 0; qp < 4; qp++) {
      strain_linear(&g, &lu[a], ety_lin[a], qp, B[qp][a]);
      if (c->model == 1) {
        double eb[9];
        strain_bilinear(&g, lq, &lu[a], etb, qp, eb);
        for (int i = 0; i < 8; i++) B[qp][a][i] += eb[i];
      }
    }
  }

  if (res) {
    for (int a = 0; a < NDOF; a++) {
      double r = 0.0;
      for (int qp = 0; qp < 4; qp++) {
        double t = 0.0;
        for (int i = 0; i < 9; i++) t += B[qp][a][i] * s[qp][i];
        r += g.detXd[qp] * t;
      }
      res[a] = r;
    }
  }

  if (mat) {
    double CB[4][24][9];
    for (int qp = 0; qp < 4; qp++)
      for (int a = 0; a < NDOF; a++) stress(c->Cs, B[qp][a], CB[qp][a]);
    for (int a = 0; a < NDOF; a++) {
      for (int b = a; b < NDOF; b++) {
        double etb[9];
        if (c->model == 1) tying_bilinear(&lu[a], &lu[b], etb);
        double k = 0.0;
        for (int qp = 0; qp < 4; qp++) {
          double t = 0.0;
          for (int i = 0; i < 9; i++) t += B[qp][a][i] * CB[qp][b][i];
          if (c->model == 1) {
            double eb[9];
            strain_bilinear(&g, &lu[a], &lu[b], etb, qp, eb);
            for (int i = 0; i < 8; i++) t += s[qp][i] * eb[i];
          }
          k += g.detXd[qp] * t;
        }
        mat[24 * a + b] = alpha * k;
        mat[24 * b + a] = alpha * k;
      }
    }
  }
  free(lu);
}

void oracle_residual(const oracle_comp_t *c, const double X[12], const double q[24],
                     double res[24]) {
  res_and_tangent(c, 1.0, c->temperature, X, q, res, NULL);
}

void oracle_jacobian(const oracle_comp_t *c, double alpha, const double X[12],
                     const double q[24], double res[24], double mat[576]) {
  res_and_tangent(c, alpha, c->temperature, X, q, res, mat);
}

/* TACSShellElement::getMatType, TACSShellElement.h:675-771 */
/* ---- inertial terms: TACSShellElement::addResidual :410-447, addJacobian :614-648,
   TACSLinearizedRotation::computeDirectorRates (TACSDirector.h:232-262),
   addDirectorResidual (:349-367), addDirectorJacobian (:369-486).
   res (may be NULL) += M qdd;  mat (may be NULL) += gamma * M. --------------------------- */
static void skew_mat(const double a[3], const double B[9], double D[9]) { /* a^x B, TACSElementAlgebra.h:1331 */
  for (int j = 0; j < 3; j++) {
    D[j] = a[1] * B[6 + j] - a[2] * B[3 + j];
    D[3 + j] = a[2] * B[j] - a[0] * B[6 + j];
    D[6 + j] = a[0] * B[3 + j] - a[1] * B[j];
  }
}
static void skew_mat_skew(const double a[3], const double B[9], const double c[3], double D[9]) {
  /* a^x B c^x, TACSElementAlgebra.h:1362 */
  double t[9];
  skew_mat(a, B, t);
  for (int i = 0; i < 3; i++) {
    D[3 * i] = c[2] * t[3 * i + 1] - c[1] * t[3 * i + 2];
    D[3 * i + 1] = c[0] * t[3 * i + 2] - c[2] * t[3 * i];
    D[3 * i + 2] = c[1] * t[3 * i] - c[0] * t[3 * i + 1];
  }
}

static void inertia(const oracle_comp_t *c, double gamma, const double X[12],
                    const double qdd[24], double res[24], double mat[576]) {
  geo_t g;
  geometry(c, X, &g);
  double dddot[12], dd[12], d2Tdotd[144], d2Tdotu[144];
  static const double zero24[24] = {0};
  if (!qdd) qdd = zero24;
  for (int i = 0; i < 4; i++) cross3(&qdd[6 * i + 3], &g.fn[3 * i], &dddot[3 * i]);
  memset(dd, 0, sizeof(dd));
  memset(d2Tdotd, 0, sizeof(d2Tdotd));
  memset(d2Tdotu, 0, sizeof(d2Tdotu));
  for (int q = 0; q < 4; q++) {
    const shape_t *sh = &g.sq[q];
    const double det = g.detXd[q];
    double u0dd[3], d0dd[3];
    interp3(sh, qdd, 6, u0dd);
    interp3(sh, dddot, 3, d0dd);
    for (int i = 0; i < 4; i++)
      for (int k = 0; k < 3; k++) {
        if (res) res[6 * i + k] += sh->N[i] * det * (c->mom[0] * u0dd[k] + c->mom[1] * d0dd[k]);
        dd[3 * i + k] += sh->N[i] * det * (c->mom[1] * u0dd[k] + c->mom[2] * d0dd[k]);
      }
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        const double nn = sh->N[i] * sh->N[j];
        for (int k = 0; k < 3; k++) {
          if (mat) mat[24 * (6 * i + k) + 6 * j + k] += gamma * det * c->mom[0] * nn;
          d2Tdotd[12 * (3 * i + k) + 3 * j + k] += det * c->mom[2] * nn;
          d2Tdotu[12 * (3 * i + k) + 3 * j + k] += det * c->mom[1] * nn;
        }
      }
  }
  for (int i = 0; i < 4; i++) {
    if (res) { /* crossProductAdd(1.0, t, dd, r) */
      double r[3];
      cross3(&g.fn[3 * i], &dd[3 * i], r);
      for (int k = 0; k < 3; k++) res[6 * i + 3 + k] += r[k];
    }
    if (!mat) continue;
    for (int j = 0; j < 4; j++) {
      double d[9], tmp[9];
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) d[3 * a + b] = gamma * d2Tdotd[12 * (3 * i + a) + 3 * j + b];
      skew_mat_skew(&g.fn[3 * i], d, &g.fn[3 * j], tmp);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) mat[24 * (6 * i + 3 + a) + 6 * j + 3 + b] -= tmp[3 * a + b];
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) d[3 * a + b] = gamma * d2Tdotu[12 * (3 * i + a) + 3 * j + b];
      skew_mat(&g.fn[3 * i], d, tmp);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          mat[24 * (6 * i + 3 + a) + 6 * j + b] += tmp[3 * a + b];
          mat[24 * (6 * j + b) + 6 * i + 3 + a] += tmp[3 * a + b];
        }
    }
  }
}

void oracle_jacobian_dyn(const oracle_comp_t *c, double alpha, double gamma, const double X[12],
                         const double q[24], const double qdd[24], double res[24],
                         double mat[576]) {
  oracle_jacobian(c, alpha, X, q, res, mat);
  inertia(c, gamma, X, qdd, res, mat);
}

void oracle_mat_type(const oracle_comp_t *c, int type, const double X[12],
                     const double q[24], double mat[576]) {
  if (type == 0) {
    res_and_tangent(c, 1.0, c->temperature, X, q, NULL, mat);
    return;
  }
  if (type == 2) { /* TACS_MASS_MATRIX: alpha = beta = 0, gamma = 1, :700-705, :769 */
    memset(mat, 0, 576 * sizeof(double));
    inertia(c, 1.0, X, NULL, NULL, mat);
    return;
  }
  /* geometric stiffness: central difference of the nonlinear twin's tangent
     along the element's own state (and temperature), :705-751 */
  oracle_comp_t nl = *c;
  nl.model = 1;
  const double dh = 1e-4;
  double norm = 0.0;
  for (int i = 0; i < 24; i++) norm += q[i] * q[i];
  norm += c->temperature * c->temperature;
  if (norm == 0.0) norm = 1.0; else norm = sqrt(norm);
  double alpha = 0.5 * norm / dh;
  double path[24], mp[576], mm[576];
  for (int i = 0; i < 24; i++) path[i] = dh * q[i] / norm;
  double Tp = c->temperature + dh * c->temperature / norm;
  /* For a linear-model element the perturbation is applied to the hidden nonlinear
     twin about the element's own temperature (:735,:747).  A nonlinear-model element
     is its own twin (nlElem == this, :53), so the second setTemperature call starts
     from the already perturbed value; a single call is restated here (the reference
     additionally leaves that element object's temperature changed for later calls,
     which is not a property of the assembly path and is not reproduced). */
  double Tbase = (c->model == 1) ? Tp : c->temperature;
  double Tm = Tbase - dh * Tbase / norm;
  res_and_tangent(&nl, alpha, Tp, X, path, NULL, mp);
  for (int i = 0; i < 24; i++) path[i] = -dh * q[i] / norm;
  res_and_tangent(&nl, -alpha, Tm, X, path, NULL, mm);
  for (int i = 0; i < 576; i++) mat[i] = mp[i] + mm[i];
}

/* ---- single-rank assembly -------------------------------------------------- */
static int cmp_int(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

/* TACSAssembler::computeLocalNodeToNodeCSR (src/TACSAssembler.cpp:1839) followed by
   TacsSortAndUniquifyCSR (src/utils/TacsUtilities.cpp:280): every node couples to all
   nodes of every element it belongs to; columns sorted ascending, unique. */
int oracle_pattern(int n_nodes, int n_elems, const int *conn, int *rowp, int *cols) {
  int *cnt = (int *)calloc((size_t)n_nodes + 1, sizeof(int));
  for (int e = 0; e < n_elems; e++)
    for (int i = 0; i < 4; i++) cnt[conn[4 * e + i] + 1] += 4;
  for (int i = 0; i < n_nodes; i++) cnt[i + 1] += cnt[i];
  int *tmp = (int *)malloc(sizeof(int) * (size_t)cnt[n_nodes]);
  int *fill = (int *)malloc(sizeof(int) * (size_t)n_nodes);
  for (int i = 0; i < n_nodes; i++) fill[i] = cnt[i];
  for (int e = 0; e < n_elems; e++)
    for (int i = 0; i < 4; i++) {
      int r = conn[4 * e + i];
      for (int j = 0; j < 4; j++) tmp[fill[r]++] = conn[4 * e + j];
    }
  int nnz = 0;
  rowp[0] = 0;
  for (int r = 0; r < n_nodes; r++) {
    int len = cnt[r + 1] - cnt[r];
    int *row = &tmp[cnt[r]];
    qsort(row, len, sizeof(int), cmp_int);
    int last = -1;
    for (int k = 0; k < len; k++) {
      if (k == 0 || row[k] != last) {
        if (cols) cols[nnz] = row[k];
        nnz++;
        last = row[k];
      }
    }
    rowp[r + 1] = nnz;
  }
  free(cnt); free(tmp); free(fill);
  return nnz;
}

/* BCSRMat::addRowValues column search (src/bpmat/BCSRMat.cpp:1778-1827) */
static int find_col(const int *rowp, const int *cols, int r, int c) {
  int lo = rowp[r], hi = rowp[r + 1] - 1;
  while (lo <= hi) {
    int mid = (lo + hi) / 2;
    if (cols[mid] == c) return mid;
    if (cols[mid] < c) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

int oracle_assemble(int op, double alpha, int n_nodes, int n_elems, const int *conn,
                    const int *elem_comp, const oracle_comp_t *comps, const double *X,
                    const double *u, int n_bc, const int *bc_nodes, const int *bc_vars,
                    const double *bc_vals, const int *rowp, const int *cols, double *res,
                    double *A) {
  return oracle_assemble_dyn(op, alpha, 0.0, n_nodes, n_elems, conn, elem_comp, comps, X, u, NULL,
                             n_bc, bc_nodes, bc_vars, bc_vals, rowp, cols, res, A);
}

int oracle_assemble_dyn(int op, double alpha, double gamma, int n_nodes, int n_elems,
                        const int *conn, const int *elem_comp, const oracle_comp_t *comps,
                        const double *X, const double *u, const double *udd, int n_bc,
                        const int *bc_nodes, const int *bc_vars, const double *bc_vals,
                        const int *rowp, const int *cols, double *res, double *A) {
  int missing = 0;
  if (res) memset(res, 0, sizeof(double) * 6 * (size_t)n_nodes);
  if (A) memset(A, 0, sizeof(double) * 36 * (size_t)rowp[n_nodes]);
  /* element loop, src/TACSAssembler.cpp:4038-4054 / :4131-4153 / :4228-4242 */
  for (int e = 0; e < n_elems; e++) {
    const int *nd = &conn[4 * e];
    const oracle_comp_t *c = &comps[elem_comp ? elem_comp[e] : 0];
    double Xe[12], qe[24], qdde[24], re[24], me[576];
    for (int i = 0; i < 4; i++) {
      memcpy(&Xe[3 * i], &X[3 * (size_t)nd[i]], 3 * sizeof(double));
      memcpy(&qe[6 * i], &u[6 * (size_t)nd[i]], 6 * sizeof(double));
      if (udd) memcpy(&qdde[6 * i], &udd[6 * (size_t)nd[i]], 6 * sizeof(double));
      else memset(&qdde[6 * i], 0, 6 * sizeof(double));
    }
    if (op == 0) { oracle_residual(c, Xe, qe, re); if (udd) inertia(c, 0.0, Xe, qdde, re, NULL); }
    else if (op == 1) oracle_jacobian_dyn(c, alpha, gamma, Xe, qe, qdde, re, me);
    else oracle_mat_type(c, op - 2, Xe, qe, me);
    if (res && op <= 1)
      for (int i = 0; i < 4; i++)
        for (int k = 0; k < 6; k++) res[6 * (size_t)nd[i] + k] += re[6 * i + k];
    if (A && op >= 1) {
      /* TACSAssembler::addMatValues -> BCSRMat::addRowValues, block (i,j) row-major */
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
          int k = find_col(rowp, cols, nd[i], nd[j]);
          if (k < 0) { missing++; continue; }
          double *a = &A[36 * (size_t)k];
          for (int r = 0; r < 6; r++)
            for (int cc = 0; cc < 6; cc++) a[6 * r + cc] += me[24 * (6 * i + r) + 6 * j + cc];
        }
    }
  }
  /* residual BCs r[bc] = u[bc] - ubar, TACSBVec::applyBCs (src/bpmat/TACSBVec.cpp:546-585),
     only where a residual is produced (src/TACSAssembler.cpp:4062, :4169) */
  if (res && op <= 1)
    for (int b = 0; b < n_bc; b++)
      for (int k = 0; k < 6; k++)
        if (bc_vars[b] & (1 << k))
          res[6 * (size_t)bc_nodes[b] + k] = u[6 * (size_t)bc_nodes[b] + k] - bc_vals[6 * b + k];
  /* matrix BCs: zero the DOF rows, 1 on the diagonal entry, columns untouched
     (BCSRMat::zeroRow, src/bpmat/BCSRMat.cpp:2005-2030) */
  if (A && op >= 1)
    for (int b = 0; b < n_bc; b++) {
      int row = bc_nodes[b];
      for (int j = rowp[row]; j < rowp[row + 1]; j++) {
        double *a = &A[36 * (size_t)j];
        for (int ii = 0; ii < 6; ii++)
          if (bc_vars[b] & (1 << ii))
            for (int jj = 0; jj < 6; jj++) a[6 * ii + jj] = 0.0;
        if (cols[j] == row)
          for (int ii = 0; ii < 6; ii++)
            if (bc_vars[b] & (1 << ii)) a[7 * ii] = 1.0;
      }
    }
  return missing;
}
