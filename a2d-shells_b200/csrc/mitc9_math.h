// mitc9_math.h — element math of the 9-node MITC shell (TACSQuad9Shell =
// TACSShellElement<TACSQuadQuadraticQuadrature, TACSShellQuadBasis<3>, TACSLinearizedRotation,
// TACSShellLinearModel>, src/elements/shell/TACSShellElementDefs.h:16-18), written as the work
// items of k_assemble9 (assemble9_kernels.cuh): one thread block per element, every function
// below is the job of ONE thread for one node / tying point / Gauss point / table entry /
// element column.  Host-callable as well: tests/host_emul.cpp steps the same functions on the
// CPU against the reference-derived fixtures (tests/golden/quad9.npz).
//
// Strain order as the reference (TACSShellElementModel.h:33-77, TACSShellConstitutive.h:125):
//   0,1,2 membrane (e11, e22, 2e12)   3,4,5 bending   6,7 transverse shear (2e23, 2e13)   8 drill
// Rows 0,1,2,6,7 come from the 28 tying strains (6 g11, 6 g22, 4 g12, 6 g23, 6 g13;
// TACSShellElementQuadBasis.h:125-142, 487-616).  The 9 Gauss points are the tensor product of
// the 3-point rule, which is also the "order" knot set of the tying interpolation (:510-512): at
// a Gauss point the full-order Lagrange factors are 0 / 1, and each interpolated tying
// component is a combination of 2 (g11, g22, g23, g13) or 4 (g12) tying points only.
//
// The tangent is K = sum_qp w det B^T C B with B = d(strain)/d(state) written out in closed
// form (linear strain model: B does not depend on the state); the residual is
// sum_qp w det B^T C (e - T e_th) with the strains e of the state evaluated forward exactly as
// the reference orders it for the drilling strain (see drill_strain_state in mitc4_math.h).
#ifndef A2DS_MITC9_MATH_H
#define A2DS_MITC9_MATH_H

#include "mitc4_math.h"

namespace a2ds {

static const int Q9_NN = 9, Q9_NV = 54, Q9_NTY = 28, Q9_LD = 56, Q9_KROWS = 12;
// leading dimension of the B / CB tables: 68 = 4 (mod 16) puts the four strain rows a DMMA
// fragment load touches (lane & 3) into disjoint bank groups for the 8 columns (lane >> 2)
static const int Q9_LDB = 68;

// geometry + state record of one element in shared memory (written by the producer warp)
struct Elem9 {
  double X[27], q[54];
  double fn[27];                    // node normals (TacsShellComputeNodeNormals)
  double dr[27];                    // directors d_n = theta_n x fn_n (TACSDirector.h:244-267)
  double a1[27], a2[27], t2n[27];   // drill derivative coefficients per node, node normal axis
  double etn[9];                    // drill strain of the state at the nodes
  double tXxi[28][3], tXeta[28][3], tn0[28][3];   // tying-point frames
  double tUxi[28][3], tUeta[28][3], td0[28][3];   // the same of the state: u,xi  u,eta  director
  double ety[28];                   // tying strains of the state
  double qT[9][9], qA[9][9], qZ[9][9];            // Gauss point: T, Xd^-1 T, -Xd^-1 Xdz Xd^-1 T
  double qw[9];                     // quadrature weight * det(Xd)
  double qP[9][9];                  // T T^T (the natural transform's frame is not orthogonal, see q9_transform)
  double sq[9][9];                  // w det C (e - T e_th) at the Gauss points
  double qU0[9][6], qU1[9][6];      // columns 0, 1 of u0x, u1x of the state, [2 i + j]
  double shat[9][5];                // sq pulled back to the tying components (g11 g22 g12 g23 g13)
  double sigt[28];                  // ... and summed onto the tying points
};
// derivative tables and the strain-displacement tables of the Gauss point being contracted
struct Tab9 {
  double Gt[Q9_NTY][Q9_LD];         // d(tying strain t) / d(dof)
  double Gt1[Q9_NTY][Q9_LD];        // bilinear tying form with the state: Bil_t(q, 1_dof)
  double Dn[Q9_NN][Q9_LD];          // d(drill strain at node n) / d(dof)
  double B[Q9_KROWS][Q9_LDB];       // B of the current Gauss point (rows 9..11, columns 54.. zero)
  double CB[Q9_KROWS][Q9_LDB];      // w det C B (tangent) or w det C B1 (geometric stiffness)
};

// shape function tables of the element class (the same for every element: filled once per
// thread block): values and parametric derivatives at the 9 Gauss points, the 28 tying points
// and the 9 node points, node index 3 j + i
struct Shape9 {
  double Nq[9][9], Nxq[9][9], Neq[9][9];
  double Nt[Q9_NTY][9], Nxt[Q9_NTY][9], Net[Q9_NTY][9];
  double Nxn[9][9], Nen[9][9];
  int tfield[Q9_NTY];
};

// quadratic Lagrange functions on {-1, 0, 1} (TacsLagrangeLobattoShapeFuncDerivative<3>, :96-104)
A2DS_HD void q9_lag(double u, double N[3], double dN[3]) {
  N[0] = -0.5 * u * (1.0 - u); N[1] = (1.0 - u) * (1.0 + u); N[2] = 0.5 * (1.0 + u) * u;
  dN[0] = -0.5 + u; dN[1] = -2.0 * u; dN[2] = 0.5 + u;
}
// 15-digit literals of basis/TACSGaussQuadrature.h:26-30
#define Q9_G3 0.774596669241483
#define Q9_G2 0.577350269189626
A2DS_HD double q9_gauss3(int i) { return i == 0 ? -Q9_G3 : (i == 1 ? 0.0 : Q9_G3); }
A2DS_HD double q9_wt3(int i) { return i == 1 ? 8.0 / 9.0 : 5.0 / 9.0; }
// linear Lagrange functions on the reduced knots {-g2, g2}
A2DS_HD void q9_red(double u, double r[2]) {
  r[0] = (Q9_G2 - u) * (0.5 / Q9_G2);
  r[1] = (u + Q9_G2) * (0.5 / Q9_G2);
}
// tying point t -> field (0 g11, 1 g22, 2 g12, 3 g23, 4 g13) and parametric point
// (getTyingField :487, getTyingPoint :530-564)
A2DS_HD int q9_ty_point(int t, double pt[2]) {
  int field, k;
  if (t < 6) { field = 0; k = t; }
  else if (t < 12) { field = 1; k = t - 6; }
  else if (t < 16) { field = 2; k = t - 12; }
  else if (t < 22) { field = 3; k = t - 16; }
  else { field = 4; k = t - 22; }
  if (field == 0 || field == 4) {          // reduced in xi, full order in eta
    pt[0] = (k % 2) ? Q9_G2 : -Q9_G2; pt[1] = q9_gauss3(k / 2);
  } else if (field == 1 || field == 3) {   // full order in xi, reduced in eta
    pt[0] = q9_gauss3(k % 3); pt[1] = (k / 3) ? Q9_G2 : -Q9_G2;
  } else {
    pt[0] = (k % 2) ? Q9_G2 : -Q9_G2; pt[1] = (k / 2) ? Q9_G2 : -Q9_G2;
  }
  return field;
}
// shape functions and their parametric derivatives at a point, node index 3 j + i
A2DS_HD void q9_shape(const double pt[2], double N[9], double Nx[9], double Ne[9]) {
  double na[3], nb[3], da[3], db[3];
  q9_lag(pt[0], na, da);
  q9_lag(pt[1], nb, db);
  #pragma unroll
  for (int j = 0; j < 3; j++)
    #pragma unroll
    for (int i = 0; i < 3; i++) {
      N[3 * j + i] = na[i] * nb[j];
      Nx[3 * j + i] = da[i] * nb[j];
      Ne[3 * j + i] = na[i] * db[j];
    }
}
// table rows of point p: 0..8 Gauss points, 9..36 tying points, 37..45 node points
A2DS_HD void q9_shape_tables(Shape9 &H, int p) {
  double pt[2], N[9], Nx[9], Ne[9];
  if (p < 9) {
    pt[0] = q9_gauss3(p % 3); pt[1] = q9_gauss3(p / 3);
    q9_shape(pt, N, Nx, Ne);
    #pragma unroll
    for (int n = 0; n < 9; n++) { H.Nq[p][n] = N[n]; H.Nxq[p][n] = Nx[n]; H.Neq[p][n] = Ne[n]; }
  } else if (p < 9 + Q9_NTY) {
    const int t = p - 9;
    H.tfield[t] = q9_ty_point(t, pt);
    q9_shape(pt, N, Nx, Ne);
    #pragma unroll
    for (int n = 0; n < 9; n++) { H.Nt[t][n] = N[n]; H.Nxt[t][n] = Nx[n]; H.Net[t][n] = Ne[n]; }
  } else {
    const int m = p - 9 - Q9_NTY;
    pt[0] = -1.0 + (m % 3); pt[1] = -1.0 + (m / 3);
    q9_shape(pt, N, Nx, Ne);
    #pragma unroll
    for (int n = 0; n < 9; n++) { H.Nxn[m][n] = Nx[n]; H.Nen[m][n] = Ne[n]; }
  }
}
A2DS_HD void q9_interp3(const double w[9], const double *v, int ld, double f[3]) {
  f[0] = f[1] = f[2] = 0.0;
  #pragma unroll
  for (int n = 0; n < 9; n++)
    #pragma unroll
    for (int k = 0; k < 3; k++) f[k] += w[n] * v[ld * n + k];
}

// frame of a point: T = [t1 | t2 | n] (row major, columns are the axes), from X,xi and the
// (not necessarily unit) normal; TACSShellElementTransform.h:25-92 (natural, including the
// t1[0]-only projection of :42-44) and :116-213 (reference axis).  strict: reference's rounding.
A2DS_HD void q9_transform(const CompData &c, const double Xxi[3], const double n0[3], double t1[3],
                          double t2[3], double n[3]) {
  double inv = 1.0 / sqrt(sdot(n0, n0));
  n[0] = A2DS_MUL(n0[0], inv); n[1] = A2DS_MUL(n0[1], inv); n[2] = A2DS_MUL(n0[2], inv);
  if (c.transform == 0) {
    t1[0] = Xxi[0]; t1[1] = Xxi[1]; t1[2] = Xxi[2];
    const double d = sdot(n, t1);
    t1[0] = A2DS_ADD(t1[0], -A2DS_MUL(d, n[0]));
    t1[0] = A2DS_ADD(t1[0], -A2DS_MUL(d, n[0]));
    t1[0] = A2DS_ADD(t1[0], -A2DS_MUL(d, n[0]));
  } else {
    const double an = sdot(c.axis, n);
    t1[0] = A2DS_ADD(c.axis[0], -A2DS_MUL(an, n[0]));
    t1[1] = A2DS_ADD(c.axis[1], -A2DS_MUL(an, n[1]));
    t1[2] = A2DS_ADD(c.axis[2], -A2DS_MUL(an, n[2]));
  }
  inv = 1.0 / sqrt(sdot(t1, t1));
  t1[0] = A2DS_MUL(t1[0], inv); t1[1] = A2DS_MUL(t1[1], inv); t1[2] = A2DS_MUL(t1[2], inv);
  scross(n, t1, t2);
}
// inverse of Xd = [a | b | c] (columns), inv3x3 (TACSElementAlgebra.h:1980); returns det
A2DS_HD double q9_frame_inverse(const double a[3], const double b[3], const double c[3], double Xi[9]) {
  const double A[9] = {a[0], b[0], c[0], a[1], b[1], c[1], a[2], b[2], c[2]};
  const double det = A2DS_ADD(A2DS_ADD(A2DS_MUL(A[8], sdet2(A[0], A[4], A[3], A[1])),
                                       -A2DS_MUL(A[7], sdet2(A[0], A[5], A[3], A[2]))),
                              A2DS_MUL(A[6], sdet2(A[1], A[5], A[2], A[4])));
  const double di = 1.0 / det;
  Xi[0] = A2DS_MUL(sdet2(A[4], A[8], A[5], A[7]), di);
  Xi[1] = A2DS_MUL(-sdet2(A[1], A[8], A[2], A[7]), di);
  Xi[2] = A2DS_MUL(sdet2(A[1], A[5], A[2], A[4]), di);
  Xi[3] = A2DS_MUL(-sdet2(A[3], A[8], A[5], A[6]), di);
  Xi[4] = A2DS_MUL(sdet2(A[0], A[8], A[2], A[6]), di);
  Xi[5] = A2DS_MUL(-sdet2(A[0], A[5], A[2], A[3]), di);
  Xi[6] = A2DS_MUL(sdet2(A[3], A[7], A[4], A[6]), di);
  Xi[7] = A2DS_MUL(-sdet2(A[0], A[7], A[1], A[6]), di);
  Xi[8] = A2DS_MUL(sdet2(A[0], A[4], A[1], A[3]), di);
  return det;
}
A2DS_HD void q9_matmul_strict(const double A[9], const double B[9], double C[9]) {
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++)
      C[3 * i + j] = A2DS_ADD(A2DS_ADD(A2DS_MUL(A[3 * i], B[j]), A2DS_MUL(A[3 * i + 1], B[3 + j])),
                              A2DS_MUL(A[3 * i + 2], B[6 + j]));
}
// interpFieldsGrad in the reference's operation order (TACSShellElementQuadBasis.h:210-232)
A2DS_HD void q9_grad_strict(const double pt[2], const double *v, int ld, double gxi[3], double geta[3]) {
  double na[3], nb[3], da[3], db[3];
  q9_lag(pt[0], na, da);
  q9_lag(pt[1], nb, db);
  #pragma unroll
  for (int k = 0; k < 3; k++) gxi[k] = geta[k] = 0.0;
  #pragma unroll
  for (int j = 0; j < 3; j++)
    #pragma unroll
    for (int i = 0; i < 3; i++) {
      const double wx = A2DS_MUL(da[i], nb[j]), we = A2DS_MUL(na[i], db[j]);
      #pragma unroll
      for (int k = 0; k < 3; k++) {
        gxi[k] = A2DS_ADD(gxi[k], A2DS_MUL(wx, v[ld * (3 * j + i) + k]));
        geta[k] = A2DS_ADD(geta[k], A2DS_MUL(we, v[ld * (3 * j + i) + k]));
      }
    }
}

// ---- node n: normal, frame, drill strain of the state and its derivative coefficients -------
// TacsShellComputeNodeNormals (TACSShellUtilities.h:301-342), TacsShellComputeDrillStrain
// (:651-693); d(et_n)/d(u_m) = 1/2 (a0 t2 - a1 t1) with a_j = N_m,xi(n) S[0][j] + N_m,eta(n) S[1][j]
// = N_m,xi(n) a1 + N_m,eta(n) a2;  d(et_n)/d(theta_n) = -normal axis of the node frame.
A2DS_HD void q9_node(const CompData &c, Elem9 &E, int n) {
  const double pt[2] = {-1.0 + (n % 3), -1.0 + (n / 3)};
  double Xxi[3], Xeta[3], fn[3];
  q9_grad_strict(pt, E.X, 3, Xxi, Xeta);
  scross(Xxi, Xeta, fn);
  const double nrm = sqrt(sdot(fn, fn));
  if (nrm != 0.0) {
    const double inv = 1.0 / nrm;
    fn[0] = A2DS_MUL(fn[0], inv); fn[1] = A2DS_MUL(fn[1], inv); fn[2] = A2DS_MUL(fn[2], inv);
  }
  double t1[3], t2[3], nn[3];
  q9_transform(c, Xxi, fn, t1, t2, nn);
  double Xi[9], S[9];
  q9_frame_inverse(Xxi, Xeta, fn, Xi);
  const double T[9] = {t1[0], t2[0], nn[0], t1[1], t2[1], nn[1], t1[2], t2[2], nn[2]};
  q9_matmul_strict(Xi, T, S);
  double uxi[3], ueta[3];
  q9_grad_strict(pt, E.q, 6, uxi, ueta);
  const double th[3] = {E.q[6 * n + 3], E.q[6 * n + 4], E.q[6 * n + 5]};
  E.etn[n] = drill_strain_state(T, S, uxi, ueta, th);
  double w[3];
  cross(t1, t2, w);
  #pragma unroll
  for (int k = 0; k < 3; k++) {
    E.fn[3 * n + k] = fn[k];
    E.a1[3 * n + k] = 0.5 * (S[0] * t2[k] - S[1] * t1[k]);
    E.a2[3 * n + k] = 0.5 * (S[3] * t2[k] - S[4] * t1[k]);
    E.t2n[3 * n + k] = w[k];
  }
  scross(th, fn, &E.dr[3 * n]);
}

// ---- tying point t: frame vectors and the tying strain of the state ---------------------------
// TACSShellLinearModel::computeTyingStrain (TACSShellElementModel.h:33-77)
A2DS_HD void q9_tying(Elem9 &E, const Shape9 &H, int t, bool nonlinear = false) {
  const double *N = H.Nt[t], *Nx = H.Nxt[t], *Ne = H.Net[t];
  const int field = H.tfield[t];
  double Xxi[3], Xeta[3], n0[3], Uxi[3], Ueta[3], d0[3];
  q9_interp3(Nx, E.X, 3, Xxi);
  q9_interp3(Ne, E.X, 3, Xeta);
  q9_interp3(N, E.fn, 3, n0);
  q9_interp3(Nx, E.q, 6, Uxi);
  q9_interp3(Ne, E.q, 6, Ueta);
  q9_interp3(N, E.dr, 3, d0);
  #pragma unroll
  for (int k = 0; k < 3; k++) {
    E.tXxi[t][k] = Xxi[k]; E.tXeta[t][k] = Xeta[k]; E.tn0[t][k] = n0[k];
    E.tUxi[t][k] = Uxi[k]; E.tUeta[t][k] = Ueta[k]; E.td0[t][k] = d0[k];
  }
  double e;
  if (field == 0) e = dot(Uxi, Xxi);
  else if (field == 1) e = dot(Ueta, Xeta);
  else if (field == 2) e = 0.5 * (dot(Uxi, Xeta) + dot(Ueta, Xxi));
  else if (field == 3) e = 0.5 * (dot(Xeta, d0) + dot(n0, Ueta));
  else e = 0.5 * (dot(Xxi, d0) + dot(n0, Uxi));
  if (nonlinear) {   // + 1/2 Bil_t(q, q), TACSShellNonlinearModel::computeTyingStrain (:644-700)
    if (field == 0) e += 0.5 * dot(Uxi, Uxi);
    else if (field == 1) e += 0.5 * dot(Ueta, Ueta);
    else if (field == 2) e += 0.5 * dot(Uxi, Ueta);
    else if (field == 3) e += 0.5 * dot(d0, Ueta);
    else e += 0.5 * dot(d0, Uxi);
  }
  E.ety[t] = e;
}

// ---- Gauss point q: frame, Xd^-1 T, the through-thickness term, weight * det -------------------
// TACSShellElement.h:514-534, TacsShellComputeDispGrad (TACSShellUtilities.h:369-393)
A2DS_HD void q9_qp(const CompData &c, Elem9 &E, const Shape9 &H, int q, bool bil = true) {
  const double *N = H.Nq[q], *Nx = H.Nxq[q], *Ne = H.Neq[q];
  double Xxi[3], Xeta[3], n0[3], nxi[3], neta[3];
  q9_interp3(Nx, E.X, 3, Xxi);
  q9_interp3(Ne, E.X, 3, Xeta);
  q9_interp3(N, E.fn, 3, n0);
  q9_interp3(Nx, E.fn, 3, nxi);
  q9_interp3(Ne, E.fn, 3, neta);
  double t1[3], t2[3], nn[3];
  q9_transform(c, Xxi, n0, t1, t2, nn);
  double Xi[9];
  const double det = q9_frame_inverse(Xxi, Xeta, n0, Xi);
  const double T[9] = {t1[0], t2[0], nn[0], t1[1], t2[1], nn[1], t1[2], t2[2], nn[2]};
  // A = Xd^-1 T;  Z = -(Xd^-1 Xdz) A with Xdz = [n,xi | n,eta | 0]
  double A[9], P[9];
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++)
      A[3 * i + j] = Xi[3 * i] * T[j] + Xi[3 * i + 1] * T[3 + j] + Xi[3 * i + 2] * T[6 + j];
  #pragma unroll
  for (int i = 0; i < 3; i++) {
    P[3 * i] = Xi[3 * i] * nxi[0] + Xi[3 * i + 1] * nxi[1] + Xi[3 * i + 2] * nxi[2];
    P[3 * i + 1] = Xi[3 * i] * neta[0] + Xi[3 * i + 1] * neta[1] + Xi[3 * i + 2] * neta[2];
    P[3 * i + 2] = 0.0;
  }
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++) {
      E.qT[q][3 * i + j] = T[3 * i + j];
      E.qA[q][3 * i + j] = A[3 * i + j];
      E.qZ[q][3 * i + j] = -(P[3 * i] * A[j] + P[3 * i + 1] * A[3 + j] + P[3 * i + 2] * A[6 + j]);
    }
  E.qw[q] = det * (q9_wt3(q % 3) * q9_wt3(q / 3));
  if (bil)
  #pragma unroll
  for (int k = 0; k < 3; k++)
    #pragma unroll
    for (int l = 0; l < 3; l++)
      E.qP[q][3 * k + l] = T[3 * k] * T[3 * l] + T[3 * k + 1] * T[3 * l + 1] + T[3 * k + 2] * T[3 * l + 2];
}

// the 5 interpolated tying components (g11, g22, g12, g23, g13) at Gauss point q of a set of 28
// tying values with stride ld (interpTyingStrain :651-672 at a point of the "order" knot set)
A2DS_HD void q9_interp_tying(int q, const double *ty, int ld, double g[5]) {
  const int a = q % 3, b = q / 3;
  double ra[2], rb[2];
  q9_red(q9_gauss3(a), ra);
  q9_red(q9_gauss3(b), rb);
  g[0] = ra[0] * ty[ld * (2 * b)] + ra[1] * ty[ld * (2 * b + 1)];
  g[1] = rb[0] * ty[ld * (6 + a)] + rb[1] * ty[ld * (6 + 3 + a)];
  g[2] = rb[0] * (ra[0] * ty[ld * 12] + ra[1] * ty[ld * 13]) +
         rb[1] * (ra[0] * ty[ld * 14] + ra[1] * ty[ld * 15]);
  g[3] = rb[0] * ty[ld * (16 + a)] + rb[1] * ty[ld * (16 + 3 + a)];
  g[4] = ra[0] * ty[ld * (22 + 2 * b)] + ra[1] * ty[ld * (22 + 2 * b + 1)];
}
// e0ty = A^T gty A (mat3x3SymmTransformTranspose) -> strains 0, 1, 2, 6, 7
A2DS_HD void q9_membrane_shear(const double A[9], const double g[5], double e[9]) {
  // gty = [[g11 g12 g13] [g12 g22 g23] [g13 g23 0]]
  const double G[9] = {g[0], g[2], g[4], g[2], g[1], g[3], g[4], g[3], 0.0};
  double W[9];
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++) W[3 * i + j] = G[3 * i] * A[j] + G[3 * i + 1] * A[3 + j] + G[3 * i + 2] * A[6 + j];
  const double e00 = A[0] * W[0] + A[3] * W[3] + A[6] * W[6];
  const double e01 = A[0] * W[1] + A[3] * W[4] + A[6] * W[7];
  const double e02 = A[0] * W[2] + A[3] * W[5] + A[6] * W[8];
  const double e11 = A[1] * W[1] + A[4] * W[4] + A[7] * W[7];
  const double e12 = A[1] * W[2] + A[4] * W[5] + A[7] * W[8];
  e[0] = e00; e[1] = e11; e[2] = 2.0 * e01; e[6] = 2.0 * e12; e[7] = 2.0 * e02;
}
// TACSShellConstitutive::computeStress (TACSShellConstitutive.h:125-147)
A2DS_HD void q9_stress(const double Cs[22], const double e[9], double s[9]) {
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

// ---- Gauss point q: strains of the state and the weighted stresses ----------------------------
// TACSShellElement::addResidual, TACSShellElement.h:314-373 (thermal: :549-574); nonlinear
// model: TACSShellNonlinearModel::evalStrain (TACSShellElementModel.h:1112-1130).  Also the
// state's u0x / u1x columns and the stresses pulled back to the tying components, which the
// bilinear strain terms (geometric stiffness, nonlinear tangent) use.
A2DS_HD void q9_qp_state(const CompData &c, Elem9 &E, const Shape9 &H, int q, double thermal,
                         bool nonlinear = false, bool bil = true) {
  const double *N = H.Nq[q], *Nx = H.Nxq[q], *Ne = H.Neq[q];
  double e[9], g[5];
  q9_interp_tying(q, E.ety, 1, g);
  q9_membrane_shear(E.qA[q], g, e);
  double u0xi[3], u0eta[3], d0[3], d0xi[3], d0eta[3];
  q9_interp3(Nx, E.q, 6, u0xi);
  q9_interp3(Ne, E.q, 6, u0eta);
  q9_interp3(N, E.dr, 3, d0);
  q9_interp3(Nx, E.dr, 3, d0xi);
  q9_interp3(Ne, E.dr, 3, d0eta);
  const double *T = E.qT[q], *A = E.qA[q], *Z = E.qZ[q];
  double u0x[3][2], u1x[3][2];   // columns 0, 1 of T^T (u0d A) and T^T (u1d A + u0d Z)
  #pragma unroll
  for (int j = 0; j < 2; j++) {
    double v0[3], v1[3];
    #pragma unroll
    for (int k = 0; k < 3; k++) {
      v0[k] = u0xi[k] * A[j] + u0eta[k] * A[3 + j] + d0[k] * A[6 + j];
      v1[k] = d0xi[k] * A[j] + d0eta[k] * A[3 + j] + u0xi[k] * Z[j] + u0eta[k] * Z[3 + j] + d0[k] * Z[6 + j];
    }
    #pragma unroll
    for (int i = 0; i < 3; i++) {
      u0x[i][j] = T[i] * v0[0] + T[3 + i] * v0[1] + T[6 + i] * v0[2];
      u1x[i][j] = T[i] * v1[0] + T[3 + i] * v1[1] + T[6 + i] * v1[2];
      if (bil) {
        E.qU0[q][2 * i + j] = u0x[i][j];
        E.qU1[q][2 * i + j] = u1x[i][j];
      }
    }
  }
  e[3] = u1x[0][0]; e[4] = u1x[1][1]; e[5] = u1x[0][1] + u1x[1][0];
  if (nonlinear) {
    #pragma unroll
    for (int i = 0; i < 3; i++) {
      e[3] += u0x[i][0] * u1x[i][0];
      e[4] += u0x[i][1] * u1x[i][1];
      e[5] += u0x[i][0] * u1x[i][1] + u1x[i][0] * u0x[i][1];
    }
  }
  double et = 0.0;
  #pragma unroll
  for (int n = 0; n < 9; n++) et += N[n] * E.etn[n];
  e[8] = et;
  #pragma unroll
  for (int k = 0; k < 9; k++) e[k] -= thermal * c.temperature * c.eth[k];
  double s[9];
  q9_stress(c.Cs, e, s);
  #pragma unroll
  for (int k = 0; k < 9; k++) { s[k] *= E.qw[q]; E.sq[q][k] = s[k]; }
  if (!bil) return;
  // P = A Se A^T with Se = [[s0 s2 s7] [s2 s1 s6] [s7 s6 0]]: sum_i s_i e_i = sum P_mn gty_mn
  const double Se[9] = {s[0], s[2], s[7], s[2], s[1], s[6], s[7], s[6], 0.0};
  double W[9];
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++) W[3 * i + j] = A[3 * i] * Se[j] + A[3 * i + 1] * Se[3 + j] + A[3 * i + 2] * Se[6 + j];
  auto P = [&](int m, int n) { return W[3 * m] * A[3 * n] + W[3 * m + 1] * A[3 * n + 1] + W[3 * m + 2] * A[3 * n + 2]; };
  E.shat[q][0] = P(0, 0); E.shat[q][1] = P(1, 1); E.shat[q][2] = 2.0 * P(0, 1);
  E.shat[q][3] = 2.0 * P(1, 2); E.shat[q][4] = 2.0 * P(0, 2);
}

// weight of tying point t in the interpolation at Gauss point q (transpose of q9_interp_tying)
A2DS_HD double q9_ty_weight(int q, int t) {
  const int a = q % 3, b = q / 3;
  double ra[2], rb[2];
  q9_red(q9_gauss3(a), ra);
  q9_red(q9_gauss3(b), rb);
  if (t < 6) return (t / 2 == b) ? ra[t % 2] : 0.0;
  if (t < 12) { const int k = t - 6; return (k % 3 == a) ? rb[k / 3] : 0.0; }
  if (t < 16) { const int k = t - 12; return ra[k % 2] * rb[k / 2]; }
  if (t < 22) { const int k = t - 16; return (k % 3 == a) ? rb[k / 3] : 0.0; }
  const int k = t - 22;
  return (k / 2 == b) ? ra[k % 2] : 0.0;
}
// tying-level stress of tying point t: sum over the Gauss points of weight * pulled-back stress
A2DS_HD void q9_sigt(Elem9 &E, const Shape9 &H, int t) {
  const int field = H.tfield[t];
  double s = 0.0;
  #pragma unroll
  for (int q = 0; q < 9; q++) s += q9_ty_weight(q, t) * E.shat[q][field];
  E.sigt[t] = s;
}

// ---- table entries ----------------------------------------------------------------------------
// d(tying strain t)/d(dof) for the frame (Fxi, Feta, Fn): with the geometry frame (X,xi  X,eta
// n0) this is Gt, the linear tying rows; with the state's frame (u,xi  u,eta  d0) it is the
// bilinear form Bil_t(q, 1_dof) of the nonlinear tying strains
// (TACSShellNonlinearModel::computeTyingStrainDeriv, TACSShellElementModel.h:1033-1109) —
// the two have the same structure.  dof = 6 m + k (k < 3 displacement, k >= 3 rotation)
A2DS_HD double q9_gt_frame(const Elem9 &E, const Shape9 &H, const double (*Fxi)[3], const double (*Feta)[3],
                           const double (*Fn)[3], int t, int dof) {
  const int field = H.tfield[t];
  const int m = dof / 6, k = dof % 6;
  const double N = H.Nt[t][m], Nx = H.Nxt[t][m], Ne = H.Net[t][m];
  if (k < 3) {
    if (field == 0) return Nx * Fxi[t][k];
    if (field == 1) return Ne * Feta[t][k];
    if (field == 2) return 0.5 * (Nx * Feta[t][k] + Ne * Fxi[t][k]);
    if (field == 3) return 0.5 * Ne * Fn[t][k];
    return 0.5 * Nx * Fn[t][k];
  }
  if (field < 3) return 0.0;
  // X,a . (theta x fn) = theta . (fn x X,a)
  const double *v = field == 3 ? Feta[t] : Fxi[t], *f = &E.fn[3 * m];
  const int c = k - 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
  return 0.5 * N * (f[c1] * v[c2] - f[c2] * v[c1]);
}
A2DS_HD double q9_gt(const Elem9 &E, const Shape9 &H, int t, int dof) {
  return q9_gt_frame(E, H, E.tXxi, E.tXeta, E.tn0, t, dof);
}
A2DS_HD double q9_gt1(const Elem9 &E, const Shape9 &H, int t, int dof) {
  return q9_gt_frame(E, H, E.tUxi, E.tUeta, E.td0, t, dof);
}
// Dn[n][dof] = d(drill strain at node n)/d(dof)
A2DS_HD double q9_dn(const Elem9 &E, const Shape9 &H, int n, int dof) {
  const int m = dof / 6, k = dof % 6;
  if (k >= 3) return m == n ? -E.t2n[3 * n + k - 3] : 0.0;
  return H.Nxn[n][m] * E.a1[3 * n + k] + H.Nen[n][m] * E.a2[3 * n + k];
}

// ---- column `dof` at Gauss point q: Bk = d(strain)/d(dof) of the linear model (9 rows) and,
// when B1k is given, B1k = Bil(q, 1_dof), the state-dependent part of the nonlinear model's B
// (evalStrainDeriv, TACSShellElementModel.h:1172-1211; its drill row is zero) ---------------------
// The unit fields of a dof are rank one: u0x(1_dof)_ij = v_i alpha_j, u1x(1_dof)_ij = v_i beta_j
// with v_i = t_i[k] (displacement) or t_i . (e_c x fn_m) (rotation).
A2DS_HD void q9_bcol(const Elem9 &E, const Tab9 &Tb, const Shape9 &H, int q, int dof, double Bk[9],
                     double *B1k = nullptr) {
  const int m = dof / 6, k = dof % 6;
  const double N = H.Nq[q][m], Nx = H.Nxq[q][m], Ne = H.Neq[q][m];
  double g[5];
  q9_interp_tying(q, &Tb.Gt[0][dof], Q9_LD, g);
  q9_membrane_shear(E.qA[q], g, Bk);
  const double *T = E.qT[q], *A = E.qA[q], *Z = E.qZ[q];
  double al0, al1, be0, be1, v[3];
  if (k < 3) {
    al0 = Nx * A[0] + Ne * A[3];
    al1 = Nx * A[1] + Ne * A[4];
    be0 = Nx * Z[0] + Ne * Z[3];
    be1 = Nx * Z[1] + Ne * Z[4];
    v[0] = T[3 * k]; v[1] = T[3 * k + 1]; v[2] = T[3 * k + 2];
  } else {
    al0 = N * A[6];
    al1 = N * A[7];
    be0 = Nx * A[0] + Ne * A[3] + N * Z[6];
    be1 = Nx * A[1] + Ne * A[4] + N * Z[7];
    const double *f = &E.fn[3 * m];
    const int c = k - 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
    // t_i . (theta x fn) = theta . (fn x t_i)
    #pragma unroll
    for (int i = 0; i < 3; i++) v[i] = f[c1] * T[3 * c2 + i] - f[c2] * T[3 * c1 + i];
  }
  Bk[3] = be0 * v[0];
  Bk[4] = be1 * v[1];
  Bk[5] = be1 * v[0] + be0 * v[1];
  double d = 0.0;
  #pragma unroll
  for (int n = 0; n < 9; n++) d += H.Nq[q][n] * Tb.Dn[n][dof];
  Bk[8] = d;
  if (B1k) {
    q9_interp_tying(q, &Tb.Gt1[0][dof], Q9_LD, g);
    q9_membrane_shear(E.qA[q], g, B1k);
    const double *U0 = E.qU0[q], *U1 = E.qU1[q];
    double p0 = 0.0, p1 = 0.0, r0 = 0.0, r1 = 0.0;
    #pragma unroll
    for (int i = 0; i < 3; i++) {
      p0 += v[i] * U0[2 * i]; p1 += v[i] * U0[2 * i + 1];
      r0 += v[i] * U1[2 * i]; r1 += v[i] * U1[2 * i + 1];
    }
    B1k[3] = al0 * r0 + be0 * p0;
    B1k[4] = al1 * r1 + be1 * p1;
    B1k[5] = al0 * r1 + be0 * p1 + be1 * p0 + al1 * r0;
    B1k[8] = 0.0;
  }
}

// ---- geometric term of node pair (ma, mb): sum_i s_i d2(e_i)/d(dof_a) d(dof_b) ------------------
// (the second derivatives of the nonlinear strains, TACSShellElementModel.h:1033-1109, 1172-1211,
// contracted with the weighted stresses E.sq).  With the rank-one unit fields of q9_bcol every
// bilinear form is a scalar times v_a . v_b, and v_a . v_b is an entry of P = T T^T
// (displacement dofs) or of P applied to the director derivatives d_c = e_c x fn (rotation dofs);
// the tying forms use Cartesian dot products directly (P = I).  So the 6 x 6 block is
//   [ M_uu         M_ut D_b       ]      M_hh' = sum_q coef_q(h, h') P_q + (tying part) I,
//   [ D_a^T M_tu   D_a^T M_tt D_b ]      D[k][c] = (e_c x fn)[k]
// One work item = one 3 x 3 quadrant (ha, hb) of the block: out[9] row major, rows = the
// three dofs of kind ha (0 displacement, 1 rotation) of node ma.
A2DS_HD void q9_geo_quadrant(const Elem9 &E, const Shape9 &H, int ma, int mb, int ha, int hb, double out[9]) {
  double M[9];
  #pragma unroll
  for (int k = 0; k < 9; k++) M[k] = 0.0;
  // tying part: g11, g22, g12 couple displacements; g23, g13 couple a rotation with a displacement
  if (!(ha == 1 && hb == 1)) {
    double c = 0.0;
    for (int t = 0; t < Q9_NTY; t++) {
      const int field = H.tfield[t];
      const double s = E.sigt[t];
      if (ha == 0 && hb == 0) {
        if (field == 0) c += s * H.Nxt[t][ma] * H.Nxt[t][mb];
        else if (field == 1) c += s * H.Net[t][ma] * H.Net[t][mb];
        else if (field == 2) c += s * 0.5 * (H.Nxt[t][ma] * H.Net[t][mb] + H.Net[t][ma] * H.Nxt[t][mb]);
      } else {
        // the rotation node carries N, the displacement node the derivative
        const int mr = ha == 1 ? ma : mb, md = ha == 1 ? mb : ma;
        if (field == 3) c += s * 0.5 * H.Nt[t][mr] * H.Net[t][md];
        else if (field == 4) c += s * 0.5 * H.Nt[t][mr] * H.Nxt[t][md];
      }
    }
    M[0] = M[4] = M[8] = c;
  }
  // bending part
  for (int q = 0; q < 9; q++) {
    const double *A = E.qA[q], *Z = E.qZ[q], *P = E.qP[q];
    const double s3 = E.sq[q][3], s4 = E.sq[q][4], s5 = E.sq[q][5];
    const double Na = H.Nq[q][ma], Nxa = H.Nxq[q][ma], Nea = H.Neq[q][ma];
    const double Nb = H.Nq[q][mb], Nxb = H.Nxq[q][mb], Neb = H.Neq[q][mb];
    const double ua0 = Nxa * A[0] + Nea * A[3], ua1 = Nxa * A[1] + Nea * A[4];
    const double ub0 = Nxb * A[0] + Neb * A[3], ub1 = Nxb * A[1] + Neb * A[4];
    double ala0, ala1, bea0, bea1, alb0, alb1, beb0, beb1;
    if (ha == 0) { ala0 = ua0; ala1 = ua1; bea0 = Nxa * Z[0] + Nea * Z[3]; bea1 = Nxa * Z[1] + Nea * Z[4]; }
    else { ala0 = Na * A[6]; ala1 = Na * A[7]; bea0 = ua0 + Na * Z[6]; bea1 = ua1 + Na * Z[7]; }
    if (hb == 0) { alb0 = ub0; alb1 = ub1; beb0 = Nxb * Z[0] + Neb * Z[3]; beb1 = Nxb * Z[1] + Neb * Z[4]; }
    else { alb0 = Nb * A[6]; alb1 = Nb * A[7]; beb0 = ub0 + Nb * Z[6]; beb1 = ub1 + Nb * Z[7]; }
    const double cf = s3 * (ala0 * beb0 + alb0 * bea0) + s4 * (ala1 * beb1 + alb1 * bea1) +
                      s5 * (alb0 * bea1 + beb0 * ala1 + ala0 * beb1 + bea0 * alb1);
    #pragma unroll
    for (int k = 0; k < 9; k++) M[k] += cf * P[k];
  }
  // D[k][c] = (e_c x f)[k] = eps_{k c l} f_l
  const double *fa = &E.fn[3 * ma], *fb = &E.fn[3 * mb];
  const double Da[9] = {0.0, fa[2], -fa[1], -fa[2], 0.0, fa[0], fa[1], -fa[0], 0.0};
  const double Db[9] = {0.0, fb[2], -fb[1], -fb[2], 0.0, fb[0], fb[1], -fb[0], 0.0};
  double R[9];   // M D_b for a rotation column kind, else M
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++)
      R[3 * i + j] = hb == 1 ? M[3 * i] * Db[j] + M[3 * i + 1] * Db[3 + j] + M[3 * i + 2] * Db[6 + j] : M[3 * i + j];
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++)
      out[3 * i + j] = ha == 1 ? Da[i] * R[j] + Da[3 + i] * R[3 + j] + Da[6 + i] * R[6 + j] : R[3 * i + j];
}
// the whole 6 x 6 block (host emulation): blk[36] row major, rows = dofs of node ma
A2DS_HD void q9_geo_pair(const Elem9 &E, const Shape9 &H, int ma, int mb, double blk[36]) {
  #pragma unroll
  for (int ha = 0; ha < 2; ha++)
    #pragma unroll
    for (int hb = 0; hb < 2; hb++) {
      double o[9];
      q9_geo_quadrant(E, H, ma, mb, ha, hb, o);
      #pragma unroll
      for (int i = 0; i < 3; i++)
        #pragma unroll
        for (int j = 0; j < 3; j++) blk[6 * (3 * ha + i) + 3 * hb + j] = o[3 * i + j];
    }
}

// ---- mass path: M_e = sum_q w det N_a N_b [[m0 I, m1 D_b], [m1 D_a^T, m2 D_a^T D_b]] --------
// (TACSShellElement.h:410-447, 614-648 with TACSLinearizedRotation: the director rate is
// d'' = theta'' x fn, TACSDirector.h:232-262).  Needs only the node normals and w det of the
// Gauss points; blk[36] row major, rows = dofs of node ma.
A2DS_HD void q9_mass_pair(const CompData &c, const Elem9 &E, const Shape9 &H, int ma, int mb, double blk[36]) {
  double cf = 0.0;
  #pragma unroll
  for (int q = 0; q < 9; q++) cf += E.qw[q] * H.Nq[q][ma] * H.Nq[q][mb];
  const double *fa = &E.fn[3 * ma], *fb = &E.fn[3 * mb];
  const double Da[9] = {0.0, fa[2], -fa[1], -fa[2], 0.0, fa[0], fa[1], -fa[0], 0.0};
  const double Db[9] = {0.0, fb[2], -fb[1], -fb[2], 0.0, fb[0], fb[1], -fb[0], 0.0};
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++) {
      blk[6 * i + j] = (i == j) ? cf * c.mom[0] : 0.0;
      blk[6 * i + 3 + j] = cf * c.mom[1] * Db[3 * i + j];
      blk[6 * (3 + i) + j] = cf * c.mom[1] * Da[3 * j + i];
      blk[6 * (3 + i) + 3 + j] = cf * c.mom[2] * (Da[i] * Db[j] + Da[3 + i] * Db[3 + j] + Da[6 + i] * Db[6 + j]);
    }
}
// node normal only (first part of q9_node) and w det of Gauss point q only (part of q9_qp)
A2DS_HD void q9_node_normal(Elem9 &E, int n) {
  const double pt[2] = {-1.0 + (n % 3), -1.0 + (n / 3)};
  double Xxi[3], Xeta[3], fn[3];
  q9_grad_strict(pt, E.X, 3, Xxi, Xeta);
  scross(Xxi, Xeta, fn);
  const double nrm = sqrt(sdot(fn, fn));
  if (nrm != 0.0) {
    const double inv = 1.0 / nrm;
    fn[0] = A2DS_MUL(fn[0], inv); fn[1] = A2DS_MUL(fn[1], inv); fn[2] = A2DS_MUL(fn[2], inv);
  }
  #pragma unroll
  for (int k = 0; k < 3; k++) E.fn[3 * n + k] = fn[k];
}
A2DS_HD void q9_qp_det(Elem9 &E, const Shape9 &H, int q) {
  double Xxi[3], Xeta[3], n0[3], Xi[9];
  q9_interp3(H.Nxq[q], E.X, 3, Xxi);
  q9_interp3(H.Neq[q], E.X, 3, Xeta);
  q9_interp3(H.Nq[q], E.fn, 3, n0);
  E.qw[q] = q9_frame_inverse(Xxi, Xeta, n0, Xi) * (q9_wt3(q % 3) * q9_wt3(q / 3));
}

}  // namespace a2ds
#endif
