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
/* Element order: 2 = MITC4 (TACSQuad4Shell / TACSQuad4NonlinearShell, the default build), 3 = MITC9
 * (TACSQuad9Shell / TACSQuad9NonlinearShell, TACSShellElementDefs.h:16-37) — the reference's
 * element template is the same for both (TACSShellElement<quadrature, TACSShellQuadBasis<order>,
 * ...>), so this file is compiled once per order (shell_oracle_q9.c sets ORACLE_ORDER 3 and
 * includes it); the order-3 entry points are called oracle9_*. */
#ifndef ORACLE_ORDER
#define ORACLE_ORDER 2
#endif
#if ORACLE_ORDER == 3
#define oracle_strain oracle9_strain
#define oracle_residual oracle9_residual
#define oracle_jacobian oracle9_jacobian
#define oracle_jacobian_dyn oracle9_jacobian_dyn
#define oracle_mat_type oracle9_mat_type
#define oracle_pattern oracle9_pattern
#define oracle_assemble oracle9_assemble
#define oracle_assemble_dyn oracle9_assemble_dyn
#define oracle_pattern_dep oracle9_pattern_dep
#define oracle_assemble_dep oracle9_assemble_dep
#endif
#include "shell_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORD ORACLE_ORDER
#define NN (ORD * ORD)                                  /* nodes, TACSShellQuadBasis::NUM_NODES :121 */
#define NV (6 * NN)                                     /* element variables */
#define NQ (ORD * ORD)                                  /* quadrature points */
#define NTY (4 * ORD * (ORD - 1) + (ORD - 1) * (ORD - 1)) /* tying points, :125-142 */

/* Gauss abscissae and weights, 15-digit literals: basis/TACSGaussQuadrature.h:23-30 */
static const double GAUSS_PT = 0.577350269189626;
#if ORD == 3
static const double GAUSS_PTS[3] = {-0.774596669241483, 0.0, 0.774596669241483};
static const double GAUSS_WTS[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
#endif

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

/* ---- basis, TACSShellElementQuadBasis.h ---------------------------------- */
typedef struct { double N[NN], Nxi[NN], Neta[NN]; } shape_t;

static void shape_eval(const double pt[2], shape_t *s) { /* :62-114, :171-232 */
#if ORD == 2
  double na[2] = {0.5 * (1.0 - pt[0]), 0.5 * (1.0 + pt[0])};
  double nb[2] = {0.5 * (1.0 - pt[1]), 0.5 * (1.0 + pt[1])};
  const double dna[2] = {-0.5, 0.5}, dnb[2] = {-0.5, 0.5};
#else /* TacsLagrangeLobattoShapeFuncDerivative<3> :96-104 */
  const double u = pt[0], v = pt[1];
  double na[3] = {-0.5 * u * (1.0 - u), (1.0 - u) * (1.0 + u), 0.5 * (1.0 + u) * u};
  double nb[3] = {-0.5 * v * (1.0 - v), (1.0 - v) * (1.0 + v), 0.5 * (1.0 + v) * v};
  const double dna[3] = {-0.5 + u, -2.0 * u, 0.5 + u}, dnb[3] = {-0.5 + v, -2.0 * v, 0.5 + v};
#endif
  for (int j = 0; j < ORD; j++)
    for (int i = 0; i < ORD; i++) {
      s->N[ORD * j + i] = na[i] * nb[j];
      s->Nxi[ORD * j + i] = dna[i] * nb[j];
      s->Neta[ORD * j + i] = na[i] * dnb[j];
    }
}
/* interpFields<stride,3> :171 */
static void interp3(const shape_t *s, const double *v, int stride, double f[3]) {
  f[0] = f[1] = f[2] = 0.0;
  for (int n = 0; n < NN; n++)
    for (int k = 0; k < 3; k++) f[k] += s->N[n] * v[stride * n + k];
}
/* interpFieldsGrad<stride,3> :210 — grad[2k] = d/dxi, grad[2k+1] = d/deta */
static void interp3_grad(const shape_t *s, const double *v, int stride, double g[6]) {
  for (int k = 0; k < 6; k++) g[k] = 0.0;
  for (int n = 0; n < NN; n++)
    for (int k = 0; k < 3; k++) {
      g[2 * k] += s->Nxi[n] * v[stride * n + k];
      g[2 * k + 1] += s->Neta[n] * v[stride * n + k];
    }
}
static void node_point(int n, double pt[2]) { /* getNodePoint :147 */
  pt[0] = -1.0 + (2.0 / (ORD - 1)) * (n % ORD);
  pt[1] = -1.0 + (2.0 / (ORD - 1)) * (n / ORD);
}
/* tying points, getTyingPoint :530-564 with knots {-1,1} / {0}; field per index
   getTyingField :487: 0,1 g11; 2,3 g22; 4 g12; 5,6 g23; 7,8 g13 */
#if ORD == 2
static const double TY_PT[9][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}, {0, 0},
                                   {-1, 0}, {1, 0}, {0, -1}, {0, 1}};
static const int TY_FIELD[9] = {0, 0, 1, 1, 2, 3, 3, 4, 4};
#else
/* order 3: "order" knots = 3-point Gauss, "reduced" knots = 2-point Gauss (getTyingKnots :510-512);
   point of index t within its field, getTyingPoint :530-564; field blocks of 6, 6, 4, 6, 6 */
#define G3 0.774596669241483
#define G2 0.577350269189626
static const double TY_PT[28][2] = {
    {-G2, -G3}, {G2, -G3}, {-G2, 0.0}, {G2, 0.0}, {-G2, G3}, {G2, G3},     /* g11: reduced x order */
    {-G3, -G2}, {0.0, -G2}, {G3, -G2}, {-G3, G2}, {0.0, G2}, {G3, G2},     /* g22: order x reduced */
    {-G2, -G2}, {G2, -G2}, {-G2, G2}, {G2, G2},                            /* g12: reduced x reduced */
    {-G3, -G2}, {0.0, -G2}, {G3, -G2}, {-G3, G2}, {0.0, G2}, {G3, G2},     /* g23: as g22 */
    {-G2, -G3}, {G2, -G3}, {-G2, 0.0}, {G2, 0.0}, {-G2, G3}, {G2, G3}};    /* g13: as g11 */
static const int TY_FIELD[28] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2,
                                 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4};
/* TacsLagrangeShapeFunction<order> :17-30, same operation order */
static void lagrange(int order, double u, const double *knots, double *N) {
  for (int i = 0; i < order; i++) {
    N[i] = 1.0;
    for (int j = 0; j < order; j++)
      if (i != j) {
        double d = 1.0 / (knots[i] - knots[j]);
        N[i] *= (u - knots[j]) * d;
      }
  }
}
#endif
/* interpTyingStrain :651-672 (evalTyingInterp :569-616): gty = [g11 g12 g13 g22 g23 0] */
static void interp_tying(const double pt[2], const double ety[NTY], double gty[6]) {
#if ORD == 3
  static const double ko[3] = {-G3, 0.0, G3}, kr[2] = {-G2, G2};
  static const int index[5] = {0, 3, 1, 4, 2};
  double na[3], nb[3], nar[2], nbr[2], N[28], *w = N;
  lagrange(3, pt[0], ko, na);
  lagrange(3, pt[1], ko, nb);
  lagrange(2, pt[0], kr, nar);
  lagrange(2, pt[1], kr, nbr);
  for (int j = 0; j < 3; j++) for (int i = 0; i < 2; i++) *w++ = nar[i] * nb[j];   /* g11 */
  for (int j = 0; j < 2; j++) for (int i = 0; i < 3; i++) *w++ = na[i] * nbr[j];   /* g22 */
  for (int j = 0; j < 2; j++) for (int i = 0; i < 2; i++) *w++ = nar[i] * nbr[j];  /* g12 */
  for (int j = 0; j < 2; j++) for (int i = 0; i < 3; i++) *w++ = na[i] * nbr[j];   /* g23 */
  for (int j = 0; j < 3; j++) for (int i = 0; i < 2; i++) *w++ = nar[i] * nb[j];   /* g13 */
  static const int npts[5] = {6, 6, 4, 6, 6};
  const double *N0 = N;
  for (int field = 0; field < 5; field++) {
    gty[index[field]] = 0.0;
    for (int k = 0; k < npts[field]; k++, N0++, ety++) gty[index[field]] += N0[0] * ety[0];
  }
  gty[5] = 0.0;
  return;
#else
  double na[2] = {0.5 * (1.0 - pt[0]), 0.5 * (1.0 + pt[0])};
  double nb[2] = {0.5 * (1.0 - pt[1]), 0.5 * (1.0 + pt[1])};
  gty[0] = 0.0; gty[0] += (1.0 * nb[0]) * ety[0]; gty[0] += (1.0 * nb[1]) * ety[1];
  gty[3] = 0.0; gty[3] += (na[0] * 1.0) * ety[2]; gty[3] += (na[1] * 1.0) * ety[3];
  gty[1] = 0.0; gty[1] += (1.0 * 1.0) * ety[4];
  gty[4] = 0.0; gty[4] += (na[0] * 1.0) * ety[5]; gty[4] += (na[1] * 1.0) * ety[6];
  gty[2] = 0.0; gty[2] += (1.0 * nb[0]) * ety[7]; gty[2] += (1.0 * nb[1]) * ety[8];
  gty[5] = 0.0;
#endif
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
  double fn[3 * NN], Xdn[9 * NN], Tn[9 * NN], XdinvTn[9 * NN];
  shape_t sn[NN];                    /* shape functions at the nodes */
  shape_t st[NTY];                   /* ... at the tying points */
  double Xxi_t[NTY][6], n0_t[NTY][3];
  shape_t sq[NQ];                    /* ... at the Gauss points */
  double pt[NQ][2], T[NQ][9], XdinvT[NQ][9], XdinvzT[NQ][9], detXd[NQ];
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

static void geometry(const oracle_comp_t *c, const double *X, geo_t *g) {
  /* TacsShellComputeNodeNormals, TACSShellUtilities.h:301-342 */
  for (int i = 0; i < NN; i++) {
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
  for (int t = 0; t < NTY; t++) {
    shape_eval(TY_PT[t], &g->st[t]);
    interp3_grad(&g->st[t], X, 3, g->Xxi_t[t]);
    interp3(&g->st[t], g->fn, 3, g->n0_t[t]);
  }
  /* Gauss points, TACSShellElement.h:514-534 and TacsShellComputeDispGrad
     (TACSShellUtilities.h:369-393); quadrature order xi fastest
     (TACSShellElementQuadrature.h:22-27, :69-77), weight 1 for order 2 */
  for (int q = 0; q < NQ; q++) {
#if ORD == 2
    const double wq = 1.0;
    g->pt[q][0] = (q % 2 == 0) ? -GAUSS_PT : GAUSS_PT;
    g->pt[q][1] = (q / 2 == 0) ? -GAUSS_PT : GAUSS_PT;
#else
    const double wq = GAUSS_WTS[q % 3] * GAUSS_WTS[q / 3];
    g->pt[q][0] = GAUSS_PTS[q % 3];
    g->pt[q][1] = GAUSS_PTS[q / 3];
    (void)GAUSS_PT;
#endif
    shape_eval(g->pt[q], &g->sq[q]);
    double Xxi[6], n0[3], nxi[6];
    interp3_grad(&g->sq[q], X, 3, Xxi);
    interp3(&g->sq[q], g->fn, 3, n0);
    compute_transform(c, Xxi, n0, g->T[q]);
    interp3_grad(&g->sq[q], g->fn, 3, nxi);
    double Xd[9], Xdz[9], Xdinv[9], neg[9];
    frame_xn(Xxi, n0, Xd);
    frame_x0(nxi, Xdz);
    g->detXd[q] = inv3(Xd, Xdinv) * wq;
    matmul(Xdinv, Xdz, neg);
    for (int k = 0; k < 9; k++) neg[k] *= -1.0;
    matmul(Xdinv, g->T[q], g->XdinvT[q]);
    matmul(neg, g->XdinvT[q], g->XdinvzT[q]);
  }
}

/* ---- quantities that are LINEAR in the element state ---------------------- */
typedef struct {
  double d[3 * NN];                   /* director d = q x fn, TACSDirector.h:244-267 */
  double Uxi_t[NTY][6], d0_t[NTY][3]; /* at the tying points */
  double u0x[NQ][9], u1x[NQ][9];      /* at the Gauss points */
  double u0xn[NN][9], Ctn_lin[NN][9]; /* at the nodes: u0x and T^T(-q^x)T */
} lin_t;

static void linear_maps(const geo_t *g, const double v[NV], lin_t *l) {
  for (int n = 0; n < NN; n++) cross3(&v[6 * n + 3], &g->fn[3 * n], &l->d[3 * n]);
  for (int t = 0; t < NTY; t++) {
    interp3_grad(&g->st[t], v, 6, l->Uxi_t[t]);
    interp3(&g->st[t], l->d, 3, l->d0_t[t]);
  }
  /* TacsShellComputeDispGrad, TACSShellUtilities.h:395-418 */
  for (int q = 0; q < NQ; q++) {
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
  for (int n = 0; n < NN; n++) {
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
static void tying_linear(const geo_t *g, const lin_t *l, double ety[NTY]) {
  for (int t = 0; t < NTY; t++) {
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
static void tying_bilinear(const lin_t *a, const lin_t *b, double ety[NTY]) {
  for (int t = 0; t < NTY; t++) {
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
static void membrane_shear(const geo_t *g, int q, const double ety[NTY], double e[9]) {
  double gty[6], e0ty[6];
  interp_tying(g->pt[q], ety, gty);
  symm_transform_t(g->XdinvT[q], gty, e0ty);
  e[0] = e0ty[0];
  e[1] = e0ty[3];
  e[2] = 2.0 * e0ty[1];
  e[6] = 2.0 * e0ty[4];
  e[7] = 2.0 * e0ty[2];
}
static void strain_linear(const geo_t *g, const lin_t *l, const double ety[NTY], int q,
                          double e[9]) {
  membrane_shear(g, q, ety, e);
  const double *u1x = l->u1x[q];
  e[3] = u1x[0];
  e[4] = u1x[4];
  e[5] = u1x[1] + u1x[3];
  double et = 0.0;
  for (int n = 0; n < NN; n++) {
    double etn = 0.5 * (l->Ctn_lin[n][3] + l->u0xn[n][3] - l->Ctn_lin[n][1] - l->u0xn[n][1]);
    et += g->sq[q].N[n] * etn;
  }
  e[8] = et;
}
/* bilinear part: TACSShellNonlinearModel::evalStrainDeriv (TACSShellElementModel.h:1172-1211) */
static void strain_bilinear(const geo_t *g, const lin_t *a, const lin_t *b,
                            const double ety_ab[NTY], int q, double e[9]) {
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
static void forward_strain(const oracle_comp_t *c, const geo_t *g, const double *q,
                           double e_out[9 * NQ]) {
  /* drill strain at the nodes, TacsShellComputeDrillStrain (TACSShellUtilities.h:651-693) */
  double etn[NN];
  for (int i = 0; i < NN; i++) {
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
  double ety[NTY];
  tying_linear(g, &l, ety);
  if (c->model == 1) {
    double eb[NTY];
    tying_bilinear(&l, &l, eb);
    for (int t = 0; t < NTY; t++) ety[t] += 0.5 * eb[t];
  }
  for (int qp = 0; qp < NQ; qp++) {
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
    for (int n = 0; n < NN; n++) et += g->sq[qp].N[n] * etn[n];
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

void oracle_strain(const oracle_comp_t *c, const double *X, const double *q,
                   double *e, double *detXd) {
  geo_t g;
  geometry(c, X, &g);
  forward_strain(c, &g, q, e);
  for (int i = 0; i < NQ; i++) detXd[i] = g.detXd[i];
}

/* residual and (optionally) tangent by exact differentiation of the energy */
static void res_and_tangent(const oracle_comp_t *c, double alpha, double temperature,
                            const double *X, const double *q, double *res,
                            double *mat) {
  geo_t g;
  geometry(c, X, &g);
  double e[9 * NQ];
  forward_strain(c, &g, q, e);

  /* stresses of the mechanical strain, TACSShellElement.h:549-574 */
  double s[NQ][9];
  for (int qp = 0; qp < NQ; qp++) {
    double em[9];
    for (int i = 0; i < 9; i++) em[i] = e[9 * qp + i] - c->eth[i] * temperature;
    stress(c->Cs, em, s[qp]);
  }

  /* linear maps of the unit directions and of the state */
  static const int NDOF = NV;
  lin_t *lu = (lin_t *)malloc(sizeof(lin_t) * (NV + 1));
  lin_t *lq = &lu[NV];
  double ety_lin[NV][NTY];
  for (int a = 0; a < NDOF; a++) {
    double v[NV];
    memset(v, 0, sizeof(v));
    v[a] = 1.0;
    linear_maps(&g, v, &lu[a]);
    tying_linear(&g, &lu[a], ety_lin[a]);
  }
  linear_maps(&g, q, lq);

  /* B[qp][a][i] = de_i/dq_a at the state */
  double (*B)[NV][9] = (double (*)[NV][9])malloc(sizeof(double) * NQ * NV * 9);
  for (int a = 0; a < NDOF; a++) {
    double etb[NTY];
    if (c->model == 1) tying_bilinear(lq, &lu[a], etb);
    for (int qp = 0; qp < NQ; qp++) {
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
      for (int qp = 0; qp < NQ; qp++) {
        double t = 0.0;
        for (int i = 0; i < 9; i++) t += B[qp][a][i] * s[qp][i];
        r += g.detXd[qp] * t;
      }
      res[a] = r;
    }
  }

  if (mat) {
    double (*CB)[NV][9] = (double (*)[NV][9])malloc(sizeof(double) * NQ * NV * 9);
    for (int qp = 0; qp < NQ; qp++)
      for (int a = 0; a < NDOF; a++) stress(c->Cs, B[qp][a], CB[qp][a]);
    for (int a = 0; a < NDOF; a++) {
      for (int b = a; b < NDOF; b++) {
        double etb[NTY];
        if (c->model == 1) tying_bilinear(&lu[a], &lu[b], etb);
        double k = 0.0;
        for (int qp = 0; qp < NQ; qp++) {
          double t = 0.0;
          for (int i = 0; i < 9; i++) t += B[qp][a][i] * CB[qp][b][i];
          if (c->model == 1) {
            double eb[9];
            strain_bilinear(&g, &lu[a], &lu[b], etb, qp, eb);
            for (int i = 0; i < 8; i++) t += s[qp][i] * eb[i];
          }
          k += g.detXd[qp] * t;
        }
        mat[NV * a + b] = alpha * k;
        mat[NV * b + a] = alpha * k;
      }
    }
    free(CB);
  }
  free(B);
  free(lu);
}

void oracle_residual(const oracle_comp_t *c, const double *X, const double *q,
                     double *res) {
  res_and_tangent(c, 1.0, c->temperature, X, q, res, NULL);
}

void oracle_jacobian(const oracle_comp_t *c, double alpha, const double *X,
                     const double *q, double *res, double *mat) {
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

static void inertia(const oracle_comp_t *c, double gamma, const double *X,
                    const double *qdd, double *res, double *mat) {
  geo_t g;
  geometry(c, X, &g);
  double dddot[3 * NN], dd[3 * NN], d2Tdotd[9 * NN * NN], d2Tdotu[9 * NN * NN];
  static const double zero24[NV] = {0};
  if (!qdd) qdd = zero24;
  for (int i = 0; i < NN; i++) cross3(&qdd[6 * i + 3], &g.fn[3 * i], &dddot[3 * i]);
  memset(dd, 0, sizeof(dd));
  memset(d2Tdotd, 0, sizeof(d2Tdotd));
  memset(d2Tdotu, 0, sizeof(d2Tdotu));
  for (int q = 0; q < NQ; q++) {
    const shape_t *sh = &g.sq[q];
    const double det = g.detXd[q];
    double u0dd[3], d0dd[3];
    interp3(sh, qdd, 6, u0dd);
    interp3(sh, dddot, 3, d0dd);
    for (int i = 0; i < NN; i++)
      for (int k = 0; k < 3; k++) {
        if (res) res[6 * i + k] += sh->N[i] * det * (c->mom[0] * u0dd[k] + c->mom[1] * d0dd[k]);
        dd[3 * i + k] += sh->N[i] * det * (c->mom[1] * u0dd[k] + c->mom[2] * d0dd[k]);
      }
    for (int i = 0; i < NN; i++)
      for (int j = 0; j < NN; j++) {
        const double nn = sh->N[i] * sh->N[j];
        for (int k = 0; k < 3; k++) {
          if (mat) mat[NV * (6 * i + k) + 6 * j + k] += gamma * det * c->mom[0] * nn;
          d2Tdotd[3 * NN * (3 * i + k) + 3 * j + k] += det * c->mom[2] * nn;
          d2Tdotu[3 * NN * (3 * i + k) + 3 * j + k] += det * c->mom[1] * nn;
        }
      }
  }
  for (int i = 0; i < NN; i++) {
    if (res) { /* crossProductAdd(1.0, t, dd, r) */
      double r[3];
      cross3(&g.fn[3 * i], &dd[3 * i], r);
      for (int k = 0; k < 3; k++) res[6 * i + 3 + k] += r[k];
    }
    if (!mat) continue;
    for (int j = 0; j < NN; j++) {
      double d[9], tmp[9];
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) d[3 * a + b] = gamma * d2Tdotd[3 * NN * (3 * i + a) + 3 * j + b];
      skew_mat_skew(&g.fn[3 * i], d, &g.fn[3 * j], tmp);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) mat[NV * (6 * i + 3 + a) + 6 * j + 3 + b] -= tmp[3 * a + b];
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) d[3 * a + b] = gamma * d2Tdotu[3 * NN * (3 * i + a) + 3 * j + b];
      skew_mat(&g.fn[3 * i], d, tmp);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          mat[NV * (6 * i + 3 + a) + 6 * j + b] += tmp[3 * a + b];
          mat[NV * (6 * j + b) + 6 * i + 3 + a] += tmp[3 * a + b];
        }
    }
  }
}

void oracle_jacobian_dyn(const oracle_comp_t *c, double alpha, double gamma, const double *X,
                         const double *q, const double *qdd, double *res,
                         double *mat) {
  oracle_jacobian(c, alpha, X, q, res, mat);
  inertia(c, gamma, X, qdd, res, mat);
}

void oracle_mat_type(const oracle_comp_t *c, int type, const double *X,
                     const double *q, double *mat) {
  if (type == 0) {
    res_and_tangent(c, 1.0, c->temperature, X, q, NULL, mat);
    return;
  }
  if (type == 2) { /* TACS_MASS_MATRIX: alpha = beta = 0, gamma = 1, :700-705, :769 */
    memset(mat, 0, NV * NV * sizeof(double));
    inertia(c, 1.0, X, NULL, NULL, mat);
    return;
  }
  /* geometric stiffness: central difference of the nonlinear twin's tangent
     along the element's own state (and temperature), :705-751 */
  oracle_comp_t nl = *c;
  nl.model = 1;
  const double dh = 1e-4;
  double norm = 0.0;
  for (int i = 0; i < NV; i++) norm += q[i] * q[i];
  norm += c->temperature * c->temperature;
  if (norm == 0.0) norm = 1.0; else norm = sqrt(norm);
  double alpha = 0.5 * norm / dh;
  double path[NV], *mp = (double *)malloc(2 * sizeof(double) * NV * NV), *mm = mp + NV * NV;
  for (int i = 0; i < NV; i++) path[i] = dh * q[i] / norm;
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
  for (int i = 0; i < NV; i++) path[i] = -dh * q[i] / norm;
  res_and_tangent(&nl, -alpha, Tm, X, path, NULL, mm);
  for (int i = 0; i < NV * NV; i++) mat[i] = mp[i] + mm[i];
  free(mp);
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
    for (int i = 0; i < NN; i++) cnt[conn[NN * e + i] + 1] += NN;
  for (int i = 0; i < n_nodes; i++) cnt[i + 1] += cnt[i];
  int *tmp = (int *)malloc(sizeof(int) * (size_t)cnt[n_nodes]);
  int *fill = (int *)malloc(sizeof(int) * (size_t)n_nodes);
  for (int i = 0; i < n_nodes; i++) fill[i] = cnt[i];
  for (int e = 0; e < n_elems; e++)
    for (int i = 0; i < NN; i++) {
      int r = conn[NN * e + i];
      for (int j = 0; j < NN; j++) tmp[fill[r]++] = conn[NN * e + j];
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
    const int *nd = &conn[NN * e];
    const oracle_comp_t *c = &comps[elem_comp ? elem_comp[e] : 0];
    double Xe[3 * NN], qe[NV], qdde[NV], re[NV], me[NV * NV];
    for (int i = 0; i < NN; i++) {
      memcpy(&Xe[3 * i], &X[3 * (size_t)nd[i]], 3 * sizeof(double));
      memcpy(&qe[6 * i], &u[6 * (size_t)nd[i]], 6 * sizeof(double));
      if (udd) memcpy(&qdde[6 * i], &udd[6 * (size_t)nd[i]], 6 * sizeof(double));
      else memset(&qdde[6 * i], 0, 6 * sizeof(double));
    }
    if (op == 0) { oracle_residual(c, Xe, qe, re); if (udd) inertia(c, 0.0, Xe, qdde, re, NULL); }
    else if (op == 1) oracle_jacobian_dyn(c, alpha, gamma, Xe, qe, qdde, re, me);
    else oracle_mat_type(c, op - 2, Xe, qe, me);
    if (res && op <= 1)
      for (int i = 0; i < NN; i++)
        for (int k = 0; k < 6; k++) res[6 * (size_t)nd[i] + k] += re[6 * i + k];
    if (A && op >= 1) {
      /* TACSAssembler::addMatValues -> BCSRMat::addRowValues, block (i,j) row-major */
      for (int i = 0; i < NN; i++)
        for (int j = 0; j < NN; j++) {
          int k = find_col(rowp, cols, nd[i], nd[j]);
          if (k < 0) { missing++; continue; }
          double *a = &A[36 * (size_t)k];
          for (int r = 0; r < 6; r++)
            for (int cc = 0; cc < 6; cc++) a[6 * r + cc] += me[NV * (6 * i + r) + 6 * j + cc];
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

/* ---- dependent nodes --------------------------------------------------------
   A connectivity entry -(d + 1) refers to dependent node d, whose values are the weighted sum
   of the independent nodes dep_conn[dep_ptr[d] .. dep_ptr[d + 1]) (TACSAssembler::setDependentNodes,
   src/TACSAssembler.cpp:716-775). */

/* independent nodes of an element with their weights: the varp / vars / weights arrays of
   TACSAssembler::addMatValues (src/TACSAssembler.h:485-505) */
static int expand_nodes(const int *nd, const int *dep_ptr, const int *dep_conn,
                        const double *dep_w, int *varp, int *vars, double *weights) {
  int k = 0;
  varp[0] = 0;
  for (int i = 0; i < NN; i++) {
    if (nd[i] >= 0) {
      weights[k] = 1.0; vars[k] = nd[i]; k++;
    } else {
      int dep = -nd[i] - 1;
      for (int j = dep_ptr[dep]; j < dep_ptr[dep + 1]; j++, k++) {
        weights[k] = dep_w[j]; vars[k] = dep_conn[j];
      }
    }
    varp[i + 1] = k;
  }
  return k;
}

/* TACSAssembler::computeLocalNodeToNodeCSR with dependent nodes (src/TACSAssembler.cpp:1839-1935:
   every independent node behind an element couples to every other one), sorted and unique */
int oracle_pattern_dep(int n_nodes, int n_elems, const int *conn, const int *dep_ptr,
                       const int *dep_conn, int *rowp, int *cols) {
  /* independent nodes behind every element (nodeCount of the reference, :1858-1875) */
  int *eptr = (int *)malloc(sizeof(int) * ((size_t)n_elems + 1));
  eptr[0] = 0;
  for (int e = 0; e < n_elems; e++) {
    int n = 0;
    for (int i = 0; i < NN; i++) {
      int v = conn[NN * e + i];
      n += v >= 0 ? 1 : dep_ptr[-v] - dep_ptr[-v - 1];
    }
    eptr[e + 1] = eptr[e] + n;
  }
  int *vars = (int *)malloc(sizeof(int) * (size_t)(eptr[n_elems] > 0 ? eptr[n_elems] : 1));
  for (int e = 0, k = 0; e < n_elems; e++)
    for (int i = 0; i < NN; i++) {
      int v = conn[NN * e + i];
      if (v >= 0) vars[k++] = v;
      else for (int j = dep_ptr[-v - 1]; j < dep_ptr[-v]; j++) vars[k++] = dep_conn[j];
    }
  int *cnt = (int *)calloc((size_t)n_nodes + 1, sizeof(int));
  for (int e = 0; e < n_elems; e++)
    for (int a = eptr[e]; a < eptr[e + 1]; a++) cnt[vars[a] + 1] += eptr[e + 1] - eptr[e];
  for (int i = 0; i < n_nodes; i++) cnt[i + 1] += cnt[i];
  int *tmp = (int *)malloc(sizeof(int) * (size_t)(cnt[n_nodes] > 0 ? cnt[n_nodes] : 1));
  int *fill = (int *)malloc(sizeof(int) * (size_t)(n_nodes > 0 ? n_nodes : 1));
  for (int i = 0; i < n_nodes; i++) fill[i] = cnt[i];
  for (int e = 0; e < n_elems; e++)
    for (int a = eptr[e]; a < eptr[e + 1]; a++)
      for (int b = eptr[e]; b < eptr[e + 1]; b++) tmp[fill[vars[a]]++] = vars[b];
  int nnz = 0;
  rowp[0] = 0;
  for (int r = 0; r < n_nodes; r++) {
    int len = cnt[r + 1] - cnt[r];
    int *row = &tmp[cnt[r]];
    qsort(row, len, sizeof(int), cmp_int);
    int last = -1;
    for (int k = 0; k < len; k++)
      if (k == 0 || row[k] != last) {
        if (cols) cols[nnz] = row[k];
        nnz++;
        last = row[k];
      }
    rowp[r + 1] = nnz;
  }
  free(cnt); free(tmp); free(fill); free(eptr); free(vars);
  return nnz;
}

/* oracle_assemble_dyn on a mesh with dependent nodes.  Gather: the values of a dependent node
   are the weighted sum of its independent nodes, accumulated from zero in list order
   (TACSBVec::endDistributeValues, src/bpmat/TACSBVec.cpp:930-975; node locations the same way
   through TACSAssembler::setNodes).  Residual: element rows of a dependent node are summed in
   the dependent slot first and distributed with the weights afterwards (TACSBVec::setValues with
   TACS_ADD_VALUES, then beginSetValues, :855-885).  Matrix: W^T K_e W block by block
   (TACSAssembler::addMatValues -> addWeightValues, src/TACSAssembler.h:485-510,
   BCSRMat::addRowWeightValues).  X, u, udd, res have n_nodes rows (independent nodes). */
int oracle_assemble_dep(int op, double alpha, double gamma, int n_nodes, int n_elems,
                        const int *conn, const int *elem_comp, const oracle_comp_t *comps,
                        const double *X, const double *u, const double *udd, int n_dep,
                        const int *dep_ptr, const int *dep_conn, const double *dep_w, int n_bc,
                        const int *bc_nodes, const int *bc_vars, const double *bc_vals,
                        const int *rowp, const int *cols, double *res, double *A) {
  int missing = 0, max_dep = 1;
  for (int d = 0; d < n_dep; d++)
    if (dep_ptr[d + 1] - dep_ptr[d] > max_dep) max_dep = dep_ptr[d + 1] - dep_ptr[d];
  const int cap = NN * max_dep;
  int *varp = (int *)malloc(sizeof(int) * (NN + 1)), *vars = (int *)malloc(sizeof(int) * cap);
  double *w = (double *)malloc(sizeof(double) * cap);
  /* dependent rows of X, u, udd and of the residual */
  double *Xd = (double *)calloc((size_t)3 * (n_dep + 1), sizeof(double));
  double *ud = (double *)calloc((size_t)6 * (n_dep + 1), sizeof(double));
  double *uddd = (double *)calloc((size_t)6 * (n_dep + 1), sizeof(double));
  double *rd = (double *)calloc((size_t)6 * (n_dep + 1), sizeof(double));
  for (int d = 0; d < n_dep; d++)
    for (int j = dep_ptr[d]; j < dep_ptr[d + 1]; j++) {
      for (int k = 0; k < 3; k++) Xd[3 * d + k] += dep_w[j] * X[3 * (size_t)dep_conn[j] + k];
      for (int k = 0; k < 6; k++) ud[6 * d + k] += dep_w[j] * u[6 * (size_t)dep_conn[j] + k];
      if (udd) for (int k = 0; k < 6; k++) uddd[6 * d + k] += dep_w[j] * udd[6 * (size_t)dep_conn[j] + k];
    }
  if (res) memset(res, 0, sizeof(double) * 6 * (size_t)n_nodes);
  if (A) memset(A, 0, sizeof(double) * 36 * (size_t)rowp[n_nodes]);
  for (int e = 0; e < n_elems; e++) {
    const int *nd = &conn[NN * e];
    const oracle_comp_t *c = &comps[elem_comp ? elem_comp[e] : 0];
    double Xe[3 * NN], qe[NV], qdde[NV], re[NV], me[NV * NV];
    for (int i = 0; i < NN; i++) {
      const double *xs = nd[i] >= 0 ? &X[3 * (size_t)nd[i]] : &Xd[3 * (-nd[i] - 1)];
      const double *us = nd[i] >= 0 ? &u[6 * (size_t)nd[i]] : &ud[6 * (-nd[i] - 1)];
      memcpy(&Xe[3 * i], xs, 3 * sizeof(double));
      memcpy(&qe[6 * i], us, 6 * sizeof(double));
      if (udd) memcpy(&qdde[6 * i], nd[i] >= 0 ? &udd[6 * (size_t)nd[i]] : &uddd[6 * (-nd[i] - 1)], 6 * sizeof(double));
      else memset(&qdde[6 * i], 0, 6 * sizeof(double));
    }
    if (op == 0) { oracle_residual(c, Xe, qe, re); if (udd) inertia(c, 0.0, Xe, qdde, re, NULL); }
    else if (op == 1) oracle_jacobian_dyn(c, alpha, gamma, Xe, qe, qdde, re, me);
    else oracle_mat_type(c, op - 2, Xe, qe, me);
    if (res && op <= 1)
      for (int i = 0; i < NN; i++) {
        double *dst = nd[i] >= 0 ? &res[6 * (size_t)nd[i]] : &rd[6 * (-nd[i] - 1)];
        for (int k = 0; k < 6; k++) dst[k] += re[6 * i + k];
      }
    if (A && op >= 1) {
      expand_nodes(nd, dep_ptr, dep_conn, dep_w, varp, vars, w);
      for (int i = 0; i < NN; i++)
        for (int ii = varp[i]; ii < varp[i + 1]; ii++)
          for (int j = 0; j < NN; j++)
            for (int jj = varp[j]; jj < varp[j + 1]; jj++) {
              int k = find_col(rowp, cols, vars[ii], vars[jj]);
              if (k < 0) { missing++; continue; }
              double *a = &A[36 * (size_t)k];
              const double ww = w[ii] * w[jj];
              for (int r = 0; r < 6; r++)
                for (int cc = 0; cc < 6; cc++) a[6 * r + cc] += ww * me[NV * (6 * i + r) + 6 * j + cc];
            }
    }
  }
  if (res && op <= 1)
    for (int d = 0; d < n_dep; d++)
      for (int j = dep_ptr[d]; j < dep_ptr[d + 1]; j++)
        for (int k = 0; k < 6; k++) res[6 * (size_t)dep_conn[j] + k] += dep_w[j] * rd[6 * d + k];
  if (res && op <= 1)
    for (int b = 0; b < n_bc; b++)
      for (int k = 0; k < 6; k++)
        if (bc_vars[b] & (1 << k))
          res[6 * (size_t)bc_nodes[b] + k] = u[6 * (size_t)bc_nodes[b] + k] - bc_vals[6 * b + k];
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
  free(varp); free(vars); free(w); free(Xd); free(ud); free(uddd); free(rd);
  return missing;
}
