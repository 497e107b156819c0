// mitc4_math.h — per-lane arithmetic of the MITC4 director-shell assembly kernel.
//
// A warp walks through a fixed sequence of phases; inside a phase every lane executes
// the same straight-line code on its own work item (a node, a Gauss point, three columns
// of the strain matrix, a pair of nodes ...) and the phases communicate through ElemGeom /
// ElemWork records that live in shared memory.  Node and Gauss-point phases are run for a
// small batch of elements at once (one lane per node / per Gauss point), the column,
// contraction and scatter phases then take the elements of the batch one at a time with
// all 32 lanes.  The functions below are the phase bodies.  They are written against plain
// pointers so that the very same code can be stepped lane by lane on the host
// (tests/host_emul.cpp) — the CUDA kernel in a2ds.cu only adds the warp synchronisation
// between phases, the DMMA contraction and the scatter.
//
// What is computed (reference: TACSShellElement<Quad2x2, QuadBasis<2>,
// LinearizedRotation, Linear|NonlinearModel>, src/elements/shell/
// TACSShellElement.h:303-771).  With q the 24 element DOFs, every strain of the
// element is a polynomial of degree <= 2 in q:
//       e(q) = B0 q + 1/2 B1(q) q,          B(q) = de/dq = B0 + B1(q)
// and the reference's hand-differentiated code evaluates the exact gradient and
// Hessian of U = 1/2 sum_qp w (e - T eth)^T C (e - T eth).  We form those
// directly:
//       r  = sum_qp w B^T s                      s = C (e - T eth)
//       K  = sum_qp w B^T C B + Kgeo(s)          (B = B0 for the linear model)
//       G  = sum_qp w (B1^T C B0 + B0^T C B1) + Kgeo(s0),  s0 = C (B0 q - T eth)
// G is the part of the nonlinear tangent that is linear in (q, T): exactly what
// the reference's central difference (TACSShellElement.h:705-751) converges to,
// without its cancellation noise.
//
// Strain rows per Gauss point (TACSShellElementModel.h:440-455):
//   0,1,2 membrane (e11, e22, 2e12)   6,7 transverse shear (2e23, 2e13)
//   3,4,5 bending                     8   drilling penalty strain
// Rows 0,1,2,6,7 come from the 9 MITC tying strains (QuadBasis.h:530-672).
#ifndef A2DS_MITC4_MATH_H
#define A2DS_MITC4_MATH_H

#include <math.h>

#ifdef __CUDACC__
#define A2DS_HD __host__ __device__ __forceinline__
#else
#define A2DS_HD inline
#endif

namespace a2ds {

// Products and sums that must not be contracted into FMAs: used where the
// reference's own rounding has to be reproduced (see drill_strain_state).
#ifdef __CUDA_ARCH__
#define A2DS_MUL(a, b) __dmul_rn((a), (b))
#define A2DS_ADD(a, b) __dadd_rn((a), (b))
#else
// host build (tests/host_emul.cpp): a volatile temporary keeps the compiler from
// fusing, so the emulation can be built with -mfma -ffp-contract=fast to mimic nvcc
static inline double a2ds_host_mul(double a, double b) { volatile double r = a * b; return r; }
static inline double a2ds_host_add(double a, double b) { volatile double r = a + b; return r; }
#define A2DS_MUL(a, b) a2ds_host_mul((a), (b))
#define A2DS_ADD(a, b) a2ds_host_add((a), (b))
#endif

// Component (constitutive + element class) record, device resident.
// Cs/eth are produced on the host by the reference's own constitutive object
// (TACSShellConstitutive.h:34; TACSIsoShellConstitutive.cpp:192-226, 438-456).
struct CompData {
  double Cs[22];       // A[6] B[6] D[6] As[3] drill
  double eth[9];       // thermal strain per unit temperature
  double temperature;  // TACSShellElement::temperature
  double axis[3];      // normalised reference axis (transform == 1)
  double mom[3];       // mass moments (TACSShellConstitutive::evalMassMoments)
  int model;           // 0 linear strain model, 1 nonlinear
  int transform;       // 0 natural, 1 reference axis
  int coupled;         // 0: the B block of Cs is identically zero
  int pad_;
};

// leading dimension of the staged 24x24 element matrix (odd: conflict-free rows and columns)
static const int KE_LD = 25;

// Lane -> work item.  lane = 4 c + qp with c = 2 m + h:
//   qp = lane & 3   Gauss point,  m = (lane >> 3) & 3   node,  h = (lane >> 2) & 1   0: u, 1: theta
// This is the m8n8k4 DMMA fragment layout read backwards: with the contraction index
// ordered k = 4 s + qp (s = strain row) and matrix row/column  8 t + c  standing for
// DOF 6 m + 3 h + t, lane (c, qp) of an A or B fragment for tile t, k-step s holds exactly
// B_qp[s][6 m + 3 h + t] — an entry of the three columns this lane computes.  The
// operands of the contraction therefore never leave the registers of the lane that made
// them.
A2DS_HD int lane_qp(int lane) { return lane & 3; }
A2DS_HD int lane_m(int lane) { return (lane >> 3) & 3; }
A2DS_HD int lane_h(int lane) { return (lane >> 2) & 1; }

// Per Gauss point data, produced once per element by phase_qp and read by the lanes of
// that Gauss point in the column phase.
struct QpData {
  double t0[3], t1[3];  // T columns 0, 1
  double S[6], Sz[6];   // (Xd^-1 T)[i][j] and its thickness derivative, i = 0..2, j = 0..1
  double M[25];         // (e0,e1,e2,e6,e7) = M (g11,g12,g13,g22,g23)
  double w;             // det(Xd) * quadrature weight
  double P0[6], P1[6];  // T u0x[:,j], T u1x[:,j] (j = 0,1) for the state
  double e[9];          // strains of the state at this point (e[8]: drilling strain)
  double pad_[4];       // 69 doubles = 5 (mod 16): see the bank note in ElemGeom
};

// Element geometry: filled by the node phase (one lane per node) and the Gauss point
// phase (one lane per Gauss point).  A warp prepares a small batch of elements at once so
// that those phases run on distinct work items instead of 8 redundant copies.
struct ElemGeom {
  double X[12], q[24];
  double fn[12];   // unit node normals            (TacsShellComputeNodeNormals)
  double dr[12];   // directors d_m = theta_m x fn_m (TACSDirector.h:244-267)
  double wn[12];   // per node: t0 x t1 of the node frame (rotation part of the drill row)
  double cdr[36];  // per node n, slot (m ^ n) in 0..2: d(drill strain at n)/d(u_m), a 3-vector
                   // (zero for the diagonally opposite node, slot 3, which is not stored)
  double etn[4];   // nodal drill strain of the state, evaluated in the reference's order
  QpData qp[4];
  // sizeof(QpData) = 69 doubles = 5 (mod 16) and sizeof(ElemGeom) = 388 doubles = 4
  // (mod 16): the 16 (element, node) / (element, Gauss point) lanes of the batched phases
  // then fall into 16 different 8-byte shared-memory banks (measured: 8 % on the residual
  // kernel against an unlucky stride)
};

// Working set of the geometric-stiffness phase of the element currently being processed
struct ElemWork {
  double ca[4][8][2], cb[4][8][2];  // per Gauss point, per generalised node
                                    // (u_0..u_3, d_0..d_3): coefficient pairs
  double sg[4][3];     // per Gauss point: w*s3, w*s4, w*s5 (bending resultants)
  double sig[4][9];    // per Gauss point contribution to the tying-point stresses
  double sigsum[9];    // tying-point stresses summed over the Gauss points
};

A2DS_HD void cross(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
A2DS_HD double dot(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// 3x3 inverse, adjugate form as TACSElementAlgebra.h:1980 (row major)
A2DS_HD double inv3(const double A[9], double Ai[9]) {
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

// Local frame T = [t1 | t2 | n] (columns) from X,xi and the normal.
// Natural transform reproduces TACSShellElementTransform.h:25-92 including the
// projection that only updates the first component (lines 42-44, three
// successive subtractions); reference-axis transform follows :116-213.
A2DS_HD void shell_transform(const CompData &c, const double Xxi[3], const double n0[3],
                             double t1[3], double t2[3], double n[3]) {
  double inv = 1.0 / sqrt(dot(n0, n0));
  n[0] = n0[0] * inv; n[1] = n0[1] * inv; n[2] = n0[2] * inv;
  if (c.transform == 0) {
    t1[0] = Xxi[0]; t1[1] = Xxi[1]; t1[2] = Xxi[2];
    double d = dot(n, t1);
    t1[0] = t1[0] - d * n[0];
    t1[0] = t1[0] - d * n[0];
    t1[0] = t1[0] - d * n[0];
  } else {
    double an = dot(c.axis, n);
    t1[0] = c.axis[0] - an * n[0];
    t1[1] = c.axis[1] - an * n[1];
    t1[2] = c.axis[2] - an * n[2];
  }
  inv = 1.0 / sqrt(dot(t1, t1));
  t1[0] *= inv; t1[1] *= inv; t1[2] *= inv;
  cross(n, t1, t2);
}

// half edge vectors of a nodal field with stride `ld`:
//   a[i] = d/dxi at eta = -1,+1     b[i] = d/deta at xi = -1,+1
A2DS_HD void edge_vectors(const double *v, int ld, double a0[3], double a1[3], double b0[3],
                          double b1[3]) {
  for (int k = 0; k < 3; k++) {
    a0[k] = 0.5 * (v[ld + k] - v[k]);
    a1[k] = 0.5 * (v[3 * ld + k] - v[2 * ld + k]);
    b0[k] = 0.5 * (v[2 * ld + k] - v[k]);
    b1[k] = 0.5 * (v[3 * ld + k] - v[ld + k]);
  }
}

// d/dxi of a nodal field on the edge eta = -1 (side 0) / +1 (side 1), and d/deta on the
// edge xi = -1 / +1.  The side is lane dependent: read the two nodes straight from
// (shared) memory so that no register array is indexed dynamically.
A2DS_HD void edge_xi(const double *v, int ld, int side, double o[3]) {
  const double *lo = v + 2 * side * ld, *hi = lo + ld;
  o[0] = 0.5 * (hi[0] - lo[0]); o[1] = 0.5 * (hi[1] - lo[1]); o[2] = 0.5 * (hi[2] - lo[2]);
}
A2DS_HD void edge_eta(const double *v, int ld, int side, double o[3]) {
  const double *lo = v + side * ld, *hi = lo + 2 * ld;
  o[0] = 0.5 * (hi[0] - lo[0]); o[1] = 0.5 * (hi[1] - lo[1]); o[2] = 0.5 * (hi[2] - lo[2]);
}

// Drill strain of the state at node m, in the operation order of the reference
// (TacsShellComputeDrillStrain, TACSShellUtilities.h:651-693; evalDrillStrain,
// TACSDirector.h:560-564):  et = 1/2 (Ct[3] + u0x[3] - Ct[1] - u0x[1]) with
// Ct = (T^T C) T, C = I - theta^x and u0x = T^T ([u,xi | u,eta | 0] (Xd^-1 T)).
// Ct[3] and Ct[1] carry the O(1) entries of T^T T (T is not orthonormal, see
// shell_transform), so their difference loses digits relative to the rotation
// size.  The residual inherits that rounding noise; evaluating the value the same
// way keeps our residual within 1e-12 of the reference instead of 1e-12 of the
// exact value.  The derivative rows (strain_columns) do not need this.
A2DS_HD double drill_strain_state(const double T[9], const double S[9], const double uxi[3],
                                  const double ueta[3], const double th[3]) {
  // C = I - theta^x  (setMatSkew(-1, q, C), TACSElementAlgebra.h:1518)
  const double C[9] = {1.0, th[2], -th[1], -th[2], 1.0, th[0], th[1], -th[0], 1.0};
  // tmp = T^T C  (rows 0,1 only), mat3x3TransMatMult: C[3i+j] = A[i]B[j]+A[3+i]B[3+j]+A[6+i]B[6+j]
  double tc[6];
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++)
      tc[3 * i + j] = A2DS_ADD(A2DS_ADD(A2DS_MUL(T[i], C[j]), A2DS_MUL(T[3 + i], C[3 + j])),
                               A2DS_MUL(T[6 + i], C[6 + j]));
  // Ct = tmp T, entries [1] = (0,1) and [3] = (1,0), mat3x3MatMult
  const double ct1 = A2DS_ADD(A2DS_ADD(A2DS_MUL(tc[0], T[1]), A2DS_MUL(tc[1], T[4])),
                              A2DS_MUL(tc[2], T[7]));
  const double ct3 = A2DS_ADD(A2DS_ADD(A2DS_MUL(tc[3], T[0]), A2DS_MUL(tc[4], T[3])),
                              A2DS_MUL(tc[5], T[6]));
  // tmp2 = u0d S with u0d = [u,xi | u,eta | 0] (columns); columns 0,1 needed
  double us[6];
  for (int k = 0; k < 3; k++)
    for (int j = 0; j < 2; j++)
      us[2 * k + j] = A2DS_ADD(A2DS_ADD(A2DS_MUL(uxi[k], S[j]), A2DS_MUL(ueta[k], S[3 + j])),
                               A2DS_MUL(0.0, S[6 + j]));
  // u0x = T^T tmp2, entries [1] = (0,1), [3] = (1,0)
  const double u1 = A2DS_ADD(A2DS_ADD(A2DS_MUL(T[0], us[1]), A2DS_MUL(T[3], us[3])),
                             A2DS_MUL(T[6], us[5]));
  const double u3 = A2DS_ADD(A2DS_ADD(A2DS_MUL(T[1], us[0]), A2DS_MUL(T[4], us[2])),
                             A2DS_MUL(T[7], us[4]));
  return A2DS_MUL(0.5, A2DS_ADD(A2DS_ADD(A2DS_ADD(ct3, u3), -ct1), -u1));
}

// ---- phase 1: node m (lane & 3) ---------------------------------------------
// Node normal, node frame T_m, S_m = Xd^-1 T_m, director, and the drill strain of
// the state (TACSShellUtilities.h:301-342, 651-693).  This phase is written with
// non-contracted products/sums in the reference's operation order so that T_m and
// S_m — and with them the rounding of drill_strain_state — are reproduced to the
// bit (see the note there).  It is ~200 flops per node; everything downstream is
// free to use FMAs.
A2DS_HD double sdot(const double x[3], const double y[3]) {
  return A2DS_ADD(A2DS_ADD(A2DS_MUL(x[0], y[0]), A2DS_MUL(x[1], y[1])), A2DS_MUL(x[2], y[2]));
}
A2DS_HD void scross(const double x[3], const double y[3], double o[3]) {
  o[0] = A2DS_ADD(A2DS_MUL(x[1], y[2]), -A2DS_MUL(x[2], y[1]));
  o[1] = A2DS_ADD(A2DS_MUL(x[2], y[0]), -A2DS_MUL(x[0], y[2]));
  o[2] = A2DS_ADD(A2DS_MUL(x[0], y[1]), -A2DS_MUL(x[1], y[0]));
}
A2DS_HD double sdet2(double a, double b, double c, double d) {  // a*b - c*d
  return A2DS_ADD(A2DS_MUL(a, b), -A2DS_MUL(c, d));
}

template <class Rec>
A2DS_HD void phase_node(const CompData &c, Rec &s, int m) {
  double Xxi[3], Xeta[3];
  edge_xi(s.X, 3, m / 2, Xxi);    // X,xi at the node (eta = +-1)
  edge_eta(s.X, 3, m % 2, Xeta);  // X,eta at the node (xi = +-1)
  double fn[3];
  scross(Xxi, Xeta, fn);
  double nrm = sqrt(sdot(fn, fn));
  if (nrm != 0.0) {
    double inv = 1.0 / nrm;
    fn[0] = A2DS_MUL(fn[0], inv); fn[1] = A2DS_MUL(fn[1], inv); fn[2] = A2DS_MUL(fn[2], inv);
  }
  // transform (TACSShellElementTransform.h:25-92 / :116-213), strict order
  double t1[3], t2[3], n[3];
  {
    double inv = 1.0 / sqrt(sdot(fn, fn));
    n[0] = A2DS_MUL(fn[0], inv); n[1] = A2DS_MUL(fn[1], inv); n[2] = A2DS_MUL(fn[2], inv);
    if (c.transform == 0) {
      t1[0] = Xxi[0]; t1[1] = Xxi[1]; t1[2] = Xxi[2];
      double d = sdot(n, t1);
      t1[0] = A2DS_ADD(t1[0], -A2DS_MUL(d, n[0]));
      t1[0] = A2DS_ADD(t1[0], -A2DS_MUL(d, n[0]));
      t1[0] = A2DS_ADD(t1[0], -A2DS_MUL(d, n[0]));
    } else {
      double an = sdot(c.axis, n);
      t1[0] = A2DS_ADD(c.axis[0], -A2DS_MUL(an, n[0]));
      t1[1] = A2DS_ADD(c.axis[1], -A2DS_MUL(an, n[1]));
      t1[2] = A2DS_ADD(c.axis[2], -A2DS_MUL(an, n[2]));
    }
    inv = 1.0 / sqrt(sdot(t1, t1));
    t1[0] = A2DS_MUL(t1[0], inv); t1[1] = A2DS_MUL(t1[1], inv); t1[2] = A2DS_MUL(t1[2], inv);
    scross(n, t1, t2);
  }
  // Xd = [Xxi | Xeta | fn] (columns), row major; inverse as inv3x3
  // (TACSElementAlgebra.h:1980)
  const double A[9] = {Xxi[0], Xeta[0], fn[0], Xxi[1], Xeta[1], fn[1], Xxi[2], Xeta[2], fn[2]};
  double Xi[9];
  {
    double det = A2DS_ADD(A2DS_ADD(A2DS_MUL(A[8], sdet2(A[0], A[4], A[3], A[1])),
                                   -A2DS_MUL(A[7], sdet2(A[0], A[5], A[3], A[2]))),
                          A2DS_MUL(A[6], sdet2(A[1], A[5], A[2], A[4])));
    double di = 1.0 / det;
    Xi[0] = A2DS_MUL(sdet2(A[4], A[8], A[5], A[7]), di);
    Xi[1] = A2DS_MUL(-sdet2(A[1], A[8], A[2], A[7]), di);
    Xi[2] = A2DS_MUL(sdet2(A[1], A[5], A[2], A[4]), di);
    Xi[3] = A2DS_MUL(-sdet2(A[3], A[8], A[5], A[6]), di);
    Xi[4] = A2DS_MUL(sdet2(A[0], A[8], A[2], A[6]), di);
    Xi[5] = A2DS_MUL(-sdet2(A[0], A[5], A[2], A[3]), di);
    Xi[6] = A2DS_MUL(sdet2(A[3], A[7], A[4], A[6]), di);
    Xi[7] = A2DS_MUL(-sdet2(A[0], A[7], A[1], A[6]), di);
    Xi[8] = A2DS_MUL(sdet2(A[0], A[4], A[1], A[3]), di);
  }
  // S = Xd^-1 T (mat3x3MatMult), T row major with columns t1, t2, n
  const double T[9] = {t1[0], t2[0], n[0], t1[1], t2[1], n[1], t1[2], t2[2], n[2]};
  double S[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      S[3 * i + j] = A2DS_ADD(A2DS_ADD(A2DS_MUL(Xi[3 * i], T[j]), A2DS_MUL(Xi[3 * i + 1], T[3 + j])),
                              A2DS_MUL(Xi[3 * i + 2], T[6 + j]));
  // derivative of this node's drill strain w.r.t. the displacements of node mm:
  //   1/2 (a0 t2 - a1 t1),  a_j = N_mm,xi(node) S[0][j] + N_mm,eta(node) S[1][j]
  // (TacsShellComputeDrillStrain + evalDrillStrainSens, TACSShellUtilities.h:651-693, 780-818)
#pragma unroll
  for (int slot = 0; slot < 3; slot++) {
    const int mm = m ^ slot;  // the node itself, its xi neighbour, its eta neighbour
    const double Nxi = (mm / 2 == m / 2) ? ((mm % 2) ? 0.5 : -0.5) : 0.0;
    const double Neta = (mm % 2 == m % 2) ? ((mm / 2) ? 0.5 : -0.5) : 0.0;
    const double a0 = Nxi * S[0] + Neta * S[3], a1 = Nxi * S[1] + Neta * S[4];
#pragma unroll
    for (int k = 0; k < 3; k++) s.cdr[9 * m + 3 * slot + k] = 0.5 * (a0 * t2[k] - a1 * t1[k]);
  }
  {
    double uxi[3], ueta[3];
    edge_xi(s.q, 6, m / 2, uxi);
    edge_eta(s.q, 6, m % 2, ueta);
    const double th[3] = {s.q[6 * m + 3], s.q[6 * m + 4], s.q[6 * m + 5]};
    s.etn[m] = drill_strain_state(T, S, uxi, ueta, th);
  }
  double w[3], d[3];
  cross(t1, t2, w);
  scross(&s.q[6 * m + 3], fn, d);
  for (int k = 0; k < 3; k++) {
    s.fn[3 * m + k] = fn[k];
    s.wn[3 * m + k] = w[k];
    s.dr[3 * m + k] = d[k];
  }
}

// Gauss point geometry kept in registers by the lanes of a Gauss point
struct QpGeom {
  double na[2], nb[2];      // 1D shape functions at the point
  double t0[3], t1[3], tn[3];
  double S[9], Sz[9];       // Xd^-1 T and -Xd^-1 Xdz Xd^-1 T
  double M[25];             // (e0,e1,e2,e6,e7) = M (g11,g12,g13,g22,g23)
  double w;                 // det(Xd) * quadrature weight
  double P0[6], P1[6];      // T u0x[:,j], T u1x[:,j] (j = 0,1) for the state
  double e[9];              // strains of the state
};

// 2-point Gauss rule, 15-digit abscissa as the reference (TACSGaussQuadrature.h:26)
#define A2DS_GAUSS_PT 0.577350269189626

// ---- phase 2a: Gauss point geometry (lane >> 3) ------------------------------
// TACSShellElement.h:520-534 and TacsShellComputeDispGrad (TACSShellUtilities.h:361-421)
template <class Rec>
A2DS_HD void qp_geometry(const CompData &c, const Rec &s, int qp, bool want_e, bool want_P,
                         bool nonlinear, QpGeom &g) {
  const double xi = (qp & 1) ? A2DS_GAUSS_PT : -A2DS_GAUSS_PT;
  const double eta = (qp & 2) ? A2DS_GAUSS_PT : -A2DS_GAUSS_PT;
  g.na[0] = 0.5 * (1.0 - xi); g.na[1] = 0.5 * (1.0 + xi);
  g.nb[0] = 0.5 * (1.0 - eta); g.nb[1] = 0.5 * (1.0 + eta);
  double a0[3], a1[3], b0[3], b1[3];
  edge_vectors(s.X, 3, a0, a1, b0, b1);
  double Xxi[3], Xeta[3], n0[3], nxi[3], neta[3];
  double f0[3], f1[3], h0[3], h1[3];
  edge_vectors(s.fn, 3, f0, f1, h0, h1);
  const double N[4] = {g.na[0] * g.nb[0], g.na[1] * g.nb[0], g.na[0] * g.nb[1],
                       g.na[1] * g.nb[1]};
  for (int k = 0; k < 3; k++) {
    Xxi[k] = g.nb[0] * a0[k] + g.nb[1] * a1[k];
    Xeta[k] = g.na[0] * b0[k] + g.na[1] * b1[k];
    nxi[k] = g.nb[0] * f0[k] + g.nb[1] * f1[k];
    neta[k] = g.na[0] * h0[k] + g.na[1] * h1[k];
    n0[k] = N[0] * s.fn[k] + N[1] * s.fn[3 + k] + N[2] * s.fn[6 + k] + N[3] * s.fn[9 + k];
  }
  shell_transform(c, Xxi, n0, g.t0, g.t1, g.tn);
  double Xd[9] = {Xxi[0], Xeta[0], n0[0], Xxi[1], Xeta[1], n0[1], Xxi[2], Xeta[2], n0[2]};
  double Xi[9];
  g.w = inv3(Xd, Xi);
  // S = Xd^-1 T
  for (int i = 0; i < 3; i++) {
    g.S[3 * i] = Xi[3 * i] * g.t0[0] + Xi[3 * i + 1] * g.t0[1] + Xi[3 * i + 2] * g.t0[2];
    g.S[3 * i + 1] = Xi[3 * i] * g.t1[0] + Xi[3 * i + 1] * g.t1[1] + Xi[3 * i + 2] * g.t1[2];
    g.S[3 * i + 2] = Xi[3 * i] * g.tn[0] + Xi[3 * i + 1] * g.tn[1] + Xi[3 * i + 2] * g.tn[2];
  }
  // Sz = -(Xd^-1 Xdz) S with Xdz = [n,xi | n,eta | 0]
  double Z[9];
  for (int i = 0; i < 3; i++) {
    Z[3 * i] = Xi[3 * i] * nxi[0] + Xi[3 * i + 1] * nxi[1] + Xi[3 * i + 2] * nxi[2];
    Z[3 * i + 1] = Xi[3 * i] * neta[0] + Xi[3 * i + 1] * neta[1] + Xi[3 * i + 2] * neta[2];
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      g.Sz[3 * i + j] = -(Z[3 * i] * g.S[j] + Z[3 * i + 1] * g.S[3 + j]);
  // M: e0ty = S^T gty S (mat3x3SymmTransformTranspose, TACSElementAlgebra.h:1094),
  // e = (e0ty[00], e0ty[11], 2 e0ty[01], 2 e0ty[12], 2 e0ty[02]); columns ordered
  // (g11, g12, g13, g22, g23) with gty symmetric, g33 = 0.
  const int ra[5] = {0, 1, 0, 1, 0}, rb[5] = {0, 1, 1, 2, 2};
  const double fac[5] = {1.0, 1.0, 2.0, 2.0, 2.0};
  for (int r = 0; r < 5; r++) {
    const int a = ra[r], b = rb[r];
    const double f = fac[r];
    const double *S = g.S;
    g.M[5 * r + 0] = f * (S[a] * S[b]);
    g.M[5 * r + 1] = f * (S[a] * S[3 + b] + S[3 + a] * S[b]);
    g.M[5 * r + 2] = f * (S[a] * S[6 + b] + S[6 + a] * S[b]);
    g.M[5 * r + 3] = f * (S[3 + a] * S[3 + b]);
    g.M[5 * r + 4] = f * (S[3 + a] * S[6 + b] + S[6 + a] * S[3 + b]);
  }
  if (want_e || want_P) {
    // u0d = [u,xi | u,eta | d0], u1d = [d,xi | d,eta | 0] at the point
    double ua0[3], ua1[3], ub0[3], ub1[3], da0[3], da1[3], db0[3], db1[3];
    edge_vectors(s.q, 6, ua0, ua1, ub0, ub1);
    edge_vectors(s.dr, 3, da0, da1, db0, db1);
    double Y0[6], Y1[6];  // (u0d S)[:,0..1], (u1d S + u0d Sz)[:,0..1], row major 3x2
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double uxi = g.nb[0] * ua0[k] + g.nb[1] * ua1[k];
      double ueta = g.na[0] * ub0[k] + g.na[1] * ub1[k];
      double dxi = g.nb[0] * da0[k] + g.nb[1] * da1[k];
      double deta = g.na[0] * db0[k] + g.na[1] * db1[k];
      double d0 = N[0] * s.dr[k] + N[1] * s.dr[3 + k] + N[2] * s.dr[6 + k] + N[3] * s.dr[9 + k];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        Y0[2 * k + j] = uxi * g.S[j] + ueta * g.S[3 + j] + d0 * g.S[6 + j];
        Y1[2 * k + j] = dxi * g.S[j] + deta * g.S[3 + j] + uxi * g.Sz[j] + ueta * g.Sz[3 + j] +
                        d0 * g.Sz[6 + j];
      }
    }
    // local displacement gradients u0x = T^T Y0, u1x = T^T Y1 (columns j = 0,1)
    double u0x[3][2], u1x[3][2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const double y0[3] = {Y0[j], Y0[2 + j], Y0[4 + j]}, y1[3] = {Y1[j], Y1[2 + j], Y1[4 + j]};
      u0x[0][j] = dot(g.t0, y0); u0x[1][j] = dot(g.t1, y0); u0x[2][j] = dot(g.tn, y0);
      u1x[0][j] = dot(g.t0, y1); u1x[1][j] = dot(g.t1, y1); u1x[2][j] = dot(g.tn, y1);
      // P = T (T^T Y): T is not orthonormal in general, keep both factors
      if (want_P) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          g.P0[3 * j + k] = g.t0[k] * u0x[0][j] + g.t1[k] * u0x[1][j] + g.tn[k] * u0x[2][j];
          g.P1[3 * j + k] = g.t0[k] * u1x[0][j] + g.t1[k] * u1x[1][j] + g.tn[k] * u1x[2][j];
        }
      }
    }
    if (want_e) {
      // Strains of the state, evaluated forward as the reference does (computeTyingStrain
      // TACSShellElementModel.h:33/:644, interpTyingStrain, e0ty = S^T gty S, evalStrain
      // :440/:1115).  Tying strains at the 9 MITC points from the edge / centre vectors:
      double fa0[3], fa1[3], fb0[3], fb1[3];  // director averaged on the four edges
#pragma unroll
      for (int k = 0; k < 3; k++) {
        fa0[k] = 0.5 * (s.dr[k] + s.dr[3 + k]);      // eta = -1: nodes 0,1
        fa1[k] = 0.5 * (s.dr[6 + k] + s.dr[9 + k]);  // eta = +1: nodes 2,3
        fb0[k] = 0.5 * (s.dr[k] + s.dr[6 + k]);      // xi = -1: nodes 0,2
        fb1[k] = 0.5 * (s.dr[3 + k] + s.dr[9 + k]);  // xi = +1: nodes 1,3
      }
      double nA0[3], nA1[3], nB0[3], nB1[3];  // node normal averaged on the four edges
#pragma unroll
      for (int k = 0; k < 3; k++) {
        nA0[k] = 0.5 * (s.fn[k] + s.fn[3 + k]);
        nA1[k] = 0.5 * (s.fn[6 + k] + s.fn[9 + k]);
        nB0[k] = 0.5 * (s.fn[k] + s.fn[6 + k]);
        nB1[k] = 0.5 * (s.fn[3 + k] + s.fn[9 + k]);
      }
      double xc[3], yc[3], uxc[3], uyc[3];  // X,xi  X,eta  u,xi  u,eta at the centre
#pragma unroll
      for (int k = 0; k < 3; k++) {
        xc[k] = 0.5 * (a0[k] + a1[k]); yc[k] = 0.5 * (b0[k] + b1[k]);
        uxc[k] = 0.5 * (ua0[k] + ua1[k]); uyc[k] = 0.5 * (ub0[k] + ub1[k]);
      }
      double g11a = dot(ua0, a0), g11b = dot(ua1, a1);
      double g22a = dot(ub0, b0), g22b = dot(ub1, b1);
      double g12c = 0.5 * (dot(uxc, yc) + dot(uyc, xc));
      double g23a = 0.5 * (dot(b0, fb0) + dot(nB0, ub0)), g23b = 0.5 * (dot(b1, fb1) + dot(nB1, ub1));
      double g13a = 0.5 * (dot(a0, fa0) + dot(nA0, ua0)), g13b = 0.5 * (dot(a1, fa1) + dot(nA1, ua1));
      if (nonlinear) {
        g11a += 0.5 * dot(ua0, ua0); g11b += 0.5 * dot(ua1, ua1);
        g22a += 0.5 * dot(ub0, ub0); g22b += 0.5 * dot(ub1, ub1);
        g12c += 0.5 * dot(uxc, uyc);
        g23a += 0.5 * dot(fb0, ub0); g23b += 0.5 * dot(fb1, ub1);
        g13a += 0.5 * dot(fa0, ua0); g13b += 0.5 * dot(fa1, ua1);
      }
      const double g5[5] = {g.nb[0] * g11a + g.nb[1] * g11b, g12c, g.nb[0] * g13a + g.nb[1] * g13b,
                            g.na[0] * g22a + g.na[1] * g22b, g.na[0] * g23a + g.na[1] * g23b};
      const int row[5] = {0, 1, 2, 6, 7};
#pragma unroll
      for (int r = 0; r < 5; r++)
        g.e[row[r]] = g.M[5 * r] * g5[0] + g.M[5 * r + 1] * g5[1] + g.M[5 * r + 2] * g5[2] +
                      g.M[5 * r + 3] * g5[3] + g.M[5 * r + 4] * g5[4];
      g.e[3] = u1x[0][0];
      g.e[4] = u1x[1][1];
      g.e[5] = u1x[0][1] + u1x[1][0];
      if (nonlinear) {
        g.e[3] += u0x[0][0] * u1x[0][0] + u0x[1][0] * u1x[1][0] + u0x[2][0] * u1x[2][0];
        g.e[4] += u0x[0][1] * u1x[0][1] + u0x[1][1] * u1x[1][1] + u0x[2][1] * u1x[2][1];
        g.e[5] += u0x[0][0] * u1x[0][1] + u0x[1][0] * u1x[1][1] + u0x[2][0] * u1x[2][1] +
                  u1x[0][0] * u0x[0][1] + u1x[1][0] * u0x[1][1] + u1x[2][0] * u0x[2][1];
      }
      // drilling strain: nodal values (reference operation order) interpolated as
      // interpFields<1,1> does (TACSShellElement.h:350)
      double et = 0.0;
#pragma unroll
      for (int n = 0; n < 4; n++)
        et = A2DS_ADD(et, A2DS_MUL(A2DS_MUL(g.na[n % 2], g.nb[n / 2]), s.etn[n]));
      g.e[8] = et;
    }
  }
}

// ---- phase 2: Gauss point qp of one element (one lane per Gauss point) -------------
A2DS_HD void phase_qp(const CompData &c, ElemGeom &s, int qp, bool want_e, bool want_P,
                      bool nonlinear, double *Pq) {
  QpGeom g;
  qp_geometry(c, s, qp, want_e, want_P, nonlinear, g);
  QpData &d = s.qp[qp];
#pragma unroll
  for (int k = 0; k < 3; k++) { d.t0[k] = g.t0[k]; d.t1[k] = g.t1[k]; }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) { d.S[2 * i + j] = g.S[3 * i + j]; d.Sz[2 * i + j] = g.Sz[3 * i + j]; }
#pragma unroll
  for (int i = 0; i < 25; i++) d.M[i] = g.M[i];
  d.w = g.w;
  if (want_P) {
#pragma unroll
    for (int i = 0; i < 6; i++) { d.P0[i] = g.P0[i]; d.P1[i] = g.P1[i]; }
  }
  if (want_e) {
#pragma unroll
    for (int i = 0; i < 9; i++) d.e[i] = g.e[i];
  }
  if (Pq) {  // T T^T (symmetric 6) for the geometric-stiffness phase
    const double *t0 = g.t0, *t1 = g.t1, *tn = g.tn;
    Pq[0] = t0[0] * t0[0] + t1[0] * t1[0] + tn[0] * tn[0];
    Pq[1] = t0[0] * t0[1] + t1[0] * t1[1] + tn[0] * tn[1];
    Pq[2] = t0[0] * t0[2] + t1[0] * t1[2] + tn[0] * tn[2];
    Pq[3] = t0[1] * t0[1] + t1[1] * t1[1] + tn[1] * tn[1];
    Pq[4] = t0[1] * t0[2] + t1[1] * t1[2] + tn[1] * tn[2];
    Pq[5] = t0[2] * t0[2] + t1[2] * t1[2] + tn[2] * tn[2];
  }
}

// 1D shape functions at Gauss point qp
A2DS_HD void qp_shape(int qp, double na[2], double nb[2]) {
  const double xi = (qp & 1) ? A2DS_GAUSS_PT : -A2DS_GAUSS_PT;
  const double eta = (qp & 2) ? A2DS_GAUSS_PT : -A2DS_GAUSS_PT;
  na[0] = 0.5 * (1.0 - xi); na[1] = 0.5 * (1.0 + xi);
  nb[0] = 0.5 * (1.0 - eta); nb[1] = 0.5 * (1.0 + eta);
}

// Coefficients of node m at a Gauss point: derivative of the local displacement
// gradients w.r.t. the node's displacement / director:
//   d u0x[i][j] / d u_mk = T[k][i] a[j]     d u1x[i][j] / d u_mk = T[k][i] az[j]
//   d u0x[i][j] / d d_mk = T[k][i] b[j]     d u1x[i][j] / d d_mk = T[k][i] cc[j]
struct NodeCoef { double a[2], az[2], b[2], cc[2]; };

A2DS_HD void node_coef(const QpData &g, const double na[2], const double nb[2], int m,
                       NodeCoef &n) {
  const double dN = (m % 2) ? 0.5 : -0.5, dM = (m / 2) ? 0.5 : -0.5;
  const double nam = (m % 2) ? na[1] : na[0], nbm = (m / 2) ? nb[1] : nb[0];
  const double Nxi = dN * nbm, Neta = nam * dM, N = nam * nbm;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    n.a[j] = Nxi * g.S[j] + Neta * g.S[2 + j];
    n.az[j] = Nxi * g.Sz[j] + Neta * g.Sz[2 + j];
    n.b[j] = N * g.S[4 + j];
    n.cc[j] = n.a[j] + N * g.Sz[4 + j];
  }
}

// Three columns (node m, h = 0: displacements, h = 1: rotations) of the 9-row
// strain matrix at one Gauss point.  With (field, normal) = (X, fn) this is B0; with
// (u, director) it is the state dependent part B1(q):
//   field,xi / field,eta on the node's two edges and at the centre feed the tying rows,
//   `normal` averaged on the edges feeds the transverse-shear tying rows,
//   bending rows:  B0 uses the T columns, B1 uses T u1x[:,j] / T u0x[:,j].
A2DS_HD void strain_columns(const ElemGeom &s, const QpData &g, const double na[2],
                            const double nb[2], const NodeCoef &nc, int m, int h,
                            const double *field, int ld, const double *normal,
                            bool has_a, const double *pa0, const double *pa1, const double *pz0,
                            const double *pz1, bool drill, double B[9][3]) {
  const int sx = m % 2, sy = m / 2;  // which xi / eta side the node sits on
  const double dN = sx ? 0.5 : -0.5, dM = sy ? 0.5 : -0.5;
  const double nas = sx ? na[1] : na[0], nbs = sy ? nb[1] : nb[0];
  const double fn[3] = {s.fn[3 * m], s.fn[3 * m + 1], s.fn[3 * m + 2]};
  double fxi[3], feta[3];  // field,xi on the node's eta edge; field,eta on its xi edge
  edge_xi(field, ld, sy, fxi);
  edge_eta(field, ld, sx, feta);
  double G5[5][3];  // columns of (g11, g12, g13, g22, g23) interpolated to the point
  if (h == 0) {
    const double c11 = nbs * dN;          // g11 tying point on this node's eta edge
    const double c22 = nas * dM;          // g22 tying point on this node's xi edge
    const double c12x = 0.5 * (0.5 * dN), c12e = 0.5 * (0.5 * dM);  // centre point
    const double c23 = nas * (0.5 * dM);  // g23: 1/2 N,eta n0 on the xi edge
    const double c13 = nbs * (0.5 * dN);  // g13: 1/2 N,xi n0 on the eta edge
    // edge-averaged normals: xi edge -> nodes (sx, sx+2); eta edge -> nodes (2 sy, 2 sy + 1)
    const double *n23a = normal + 3 * sx, *n13a = normal + 6 * sy;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      // field,xi / field,eta at the element centre
      const double cxi = 0.25 * (field[ld + k] - field[k] + field[3 * ld + k] - field[2 * ld + k]);
      const double ceta = 0.25 * (field[2 * ld + k] - field[k] + field[3 * ld + k] - field[ld + k]);
      G5[0][k] = c11 * fxi[k];
      G5[3][k] = c22 * feta[k];
      G5[1][k] = c12x * ceta + c12e * cxi;
      G5[4][k] = c23 * (0.5 * (n23a[k] + n23a[6 + k]));
      G5[2][k] = c13 * (0.5 * (n13a[k] + n13a[3 + k]));
    }
  } else {
    // 1/2 N_m(tying point) (fn_m x field,eta|xi); N_m = 1/2 at the edge mid points
    double x23[3], x13[3];
    cross(fn, feta, x23);
    cross(fn, fxi, x13);
    const double c23 = nas * 0.25, c13 = nbs * 0.25;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      G5[0][k] = 0.0; G5[1][k] = 0.0; G5[3][k] = 0.0;
      G5[4][k] = c23 * x23[k];
      G5[2][k] = c13 * x13[k];
    }
  }
#pragma unroll
  for (int r = 0; r < 5; r++) {
    const int row = (r < 3) ? r : r + 3;  // strain rows 0,1,2,6,7
#pragma unroll
    for (int k = 0; k < 3; k++)
      B[row][k] = g.M[5 * r] * G5[0][k] + g.M[5 * r + 1] * G5[1][k] +
                  g.M[5 * r + 2] * G5[2][k] + g.M[5 * r + 3] * G5[3][k] +
                  g.M[5 * r + 4] * G5[4][k];
  }
  // bending rows: e3 = u1x[0][0], e4 = u1x[1][1], e5 = u1x[0][1] + u1x[1][0]
  // (+ u0x/u1x products for the nonlinear part)
  const double ca0 = h ? nc.b[0] : nc.a[0], ca1 = h ? nc.b[1] : nc.a[1];      // "u0x" partner
  const double cz0 = h ? nc.cc[0] : nc.az[0], cz1 = h ? nc.cc[1] : nc.az[1];  // "u1x" partner
  double A0[3] = {0.0, 0.0, 0.0}, A1[3] = {0.0, 0.0, 0.0}, Z0[3], Z1[3];
  if (h == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (has_a) { A0[k] = pa0[k]; A1[k] = pa1[k]; }
      Z0[k] = pz0[k]; Z1[k] = pz1[k];
    }
  } else {
    if (has_a) { cross(fn, pa0, A0); cross(fn, pa1, A1); }
    cross(fn, pz0, Z0);
    cross(fn, pz1, Z1);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    B[3][k] = ca0 * A0[k] + cz0 * Z0[k];
    B[4][k] = ca1 * A1[k] + cz1 * Z1[k];
    B[5][k] = ca0 * A1[k] + cz1 * Z0[k] + ca1 * A0[k] + cz0 * Z1[k];
  }
  // drilling strain row: et = sum_n N_n etn_n,
  // etn_n = 1/2 (u0x[1][0] - u0x[0][1]) - theta_n . (t0n x t1n)   (TACSDirector.h:560-564)
  if (!drill) {
    B[8][0] = B[8][1] = B[8][2] = 0.0;
  } else if (h == 0) {
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int n = 0; n < 4; n++) {
      const int slot = m ^ n;  // 3: diagonally opposite node, no contribution
      const double Nq = (slot == 3) ? 0.0 : na[n % 2] * nb[n / 2];
      const double *cd = &s.cdr[9 * n + 3 * (slot & ~(slot >> 1))];  // slot 3 -> any valid slot
#pragma unroll
      for (int k = 0; k < 3; k++) acc[k] += Nq * cd[k];
    }
    B[8][0] = acc[0]; B[8][1] = acc[1]; B[8][2] = acc[2];
  } else {
    const double Nq = nas * nbs;
#pragma unroll
    for (int k = 0; k < 3; k++) B[8][k] = -Nq * s.wn[3 * m + k];
  }
}

// s = C e (TACSShellConstitutive::computeStress, TACSShellConstitutive.h:125-147)
// `coupled` == false: the B (membrane-bending coupling) block is identically zero (isotropic
// shell without offset) and its 18 products are skipped; the flag is uniform over a warp.
A2DS_HD void apply_C(const double Cs[22], const double e[9], double s[9], bool coupled = true) {
  const double *A = &Cs[0], *B = &Cs[6], *D = &Cs[12], *As = &Cs[18];
  if (!coupled) {
    s[0] = A[0] * e[0] + A[1] * e[1] + A[2] * e[2];
    s[1] = A[1] * e[0] + A[3] * e[1] + A[4] * e[2];
    s[2] = A[2] * e[0] + A[4] * e[1] + A[5] * e[2];
    s[3] = D[0] * e[3] + D[1] * e[4] + D[2] * e[5];
    s[4] = D[1] * e[3] + D[3] * e[4] + D[4] * e[5];
    s[5] = D[2] * e[3] + D[4] * e[4] + D[5] * e[5];
    s[6] = As[0] * e[6] + As[1] * e[7];
    s[7] = As[1] * e[6] + As[2] * e[7];
    s[8] = Cs[21] * e[8];
    return;
  }
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

// What the launch is asked to produce
struct Want {
  bool res, kmat, gmat;  // residual, tangent, geometric stiffness
  bool nonlinear;        // element uses the nonlinear strain model (tangent/residual)
  double thermal;        // 1: mechanical strain e - T eth;  0: plain B q (matrix-free product)
};

// ---- column phase: lane = (qp, m, h) ------------------------------------------------
// The lane's three columns are returned IN REGISTERS, indexed [strain row][component]
// (they are the DMMA fragments, see lane_qp above):
//   lane_b1 : Bq = B1(q), the state dependent part (only when G or the nonlinear model is
//             asked for); also publishes the coefficient pairs of the geometric phase
//   lane_b0w: Bc = B0 (+ Bq for the nonlinear model) and Wc = w C Bc
// They are separate so that the kernel can run the tangent contraction between them and
// never holds B0, W and B1 at the same time.
A2DS_HD void lane_b1(const ElemGeom &s, ElemWork &wk, int lane, double Bq[9][3]) {
  const int qp = lane_qp(lane), m = lane_m(lane), h = lane_h(lane);
  const QpData &g = s.qp[qp];
  double na[2], nb[2];
  qp_shape(qp, na, nb);
  NodeCoef nc;
  node_coef(g, na, nb, m, nc);
  // coefficient pairs for the geometric stiffness phase (generalised nodes u_m, d_m)
  const int p = 4 * h + m;
  wk.ca[qp][p][0] = h ? nc.b[0] : nc.a[0]; wk.ca[qp][p][1] = h ? nc.b[1] : nc.a[1];
  wk.cb[qp][p][0] = h ? nc.cc[0] : nc.az[0]; wk.cb[qp][p][1] = h ? nc.cc[1] : nc.az[1];
  strain_columns(s, g, na, nb, nc, m, h, s.q, 6, s.dr, true, &g.P1[0], &g.P1[3], &g.P0[0],
                 &g.P0[3], false, Bq);
}

A2DS_HD void lane_b0w(const CompData &c, const ElemGeom &s, int lane, const Want &w,
                      const double Bq[9][3], double Bc[9][3], double Wc[9][3]) {
  const int qp = lane_qp(lane), m = lane_m(lane), h = lane_h(lane);
  const QpData &g = s.qp[qp];
  double na[2], nb[2];
  qp_shape(qp, na, nb);
  NodeCoef nc;
  node_coef(g, na, nb, m, nc);
  strain_columns(s, g, na, nb, nc, m, h, s.X, 3, s.fn, false, g.t0, g.t1, g.t0, g.t1, true, Bc);
  if (w.nonlinear) {  // B = B0 + B1(q)
#pragma unroll
    for (int r = 0; r < 9; r++)
#pragma unroll
      for (int k = 0; k < 3; k++) Bc[r][k] += Bq[r][k];
  }
  const double gw = g.w;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double b[9], cb[9];
#pragma unroll
    for (int r = 0; r < 9; r++) b[r] = Bc[r][k];
    apply_C(c.Cs, b, cb, c.coupled != 0);
#pragma unroll
    for (int r = 0; r < 9; r++) Wc[r][k] = gw * cb[r];
  }
}

// ---- phase 3: lane = (qp, m, h): mechanical strain -> residual partials, stresses ---
// The strains of the state at the Gauss point come from phase_qp.  Returns the lane's three
// residual entries for this Gauss point, r = W^T (e - T eth) (to be summed over the four
// Gauss points), and publishes the stresses of the geometric phase.
A2DS_HD void lane_stress(const CompData &c, const ElemGeom &s, ElemWork &wk, int lane,
                         const Want &w, const double Wc[9][3], double r3[3]) {
  const int qp = lane_qp(lane), m = lane_m(lane), h = lane_h(lane);
  double na[2], nb[2];
  qp_shape(qp, na, nb);
  const double qp_w = s.qp[qp].w;
  const double Tth = w.thermal * c.temperature;
  double e[9];
#pragma unroll
  for (int r = 0; r < 9; r++) e[r] = s.qp[qp].e[r] - Tth * c.eth[r];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double rr = 0.0;
#pragma unroll
    for (int r = 0; r < 9; r++) rr += Wc[r][k] * e[r];
    r3[k] = rr;
  }
  if ((w.gmat || w.nonlinear) && m == 0 && h == 0) {
    double st[9];
    apply_C(c.Cs, e, st);
    // stresses feeding the geometric terms
    wk.sg[qp][0] = qp_w * st[3]; wk.sg[qp][1] = qp_w * st[4]; wk.sg[qp][2] = qp_w * st[5];
    // pull the membrane/shear stresses back to the tying points:
    // dU/dg5 = M^T (w s_ms), then the tying interpolation transposed
    const double sm[5] = {qp_w * st[0], qp_w * st[1], qp_w * st[2], qp_w * st[6], qp_w * st[7]};
    const double *M = s.qp[qp].M;
    double dg[5];
#pragma unroll
    for (int cidx = 0; cidx < 5; cidx++)
      dg[cidx] = M[cidx] * sm[0] + M[5 + cidx] * sm[1] + M[10 + cidx] * sm[2] +
                 M[15 + cidx] * sm[3] + M[20 + cidx] * sm[4];
    // tying point order (QuadBasis.h:530-564): g11 @eta=-+1, g22 @xi=-+1, g12 centre,
    // g23 @xi=-+1, g13 @eta=-+1; dg order (g11, g12, g13, g22, g23)
    wk.sig[qp][0] = nb[0] * dg[0]; wk.sig[qp][1] = nb[1] * dg[0];
    wk.sig[qp][2] = na[0] * dg[3]; wk.sig[qp][3] = na[1] * dg[3];
    wk.sig[qp][4] = dg[1];
    wk.sig[qp][5] = na[0] * dg[4]; wk.sig[qp][6] = na[1] * dg[4];
    wk.sig[qp][7] = nb[0] * dg[2]; wk.sig[qp][8] = nb[1] * dg[2];
  }
}

// tying-point stresses summed over the Gauss points (lanes 0..8, after lane_stress)
A2DS_HD void sum_tying_stress(ElemWork &s, int t) {
  s.sigsum[t] = s.sig[0][t] + s.sig[1][t] + s.sig[2][t] + s.sig[3][t];
}

// Fold a 3x3 block given in terms of the generalised nodes (u_m, d_m) onto the rotation
// DOFs of the linearised rotation d = theta x fn:
//   rows of a director node:    skew(fn_m) * blk
//   columns of a director node: blk * skew(fn_mm)^T
// (TACSLinearizedRotation::addDirectorJacobian, TACSDirector.h:369-486)
A2DS_HD void fold_director_block(const ElemGeom &gm, bool pd, int m, bool ppd, int mm,
                                 double blk[9]) {
  if (pd) {  // rows: skew(fn_m) * blk
    const double *f = &gm.fn[3 * m];
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
  if (ppd) {  // columns: blk * skew(fn_mm)^T, i.e. row_i -> fn x row_i
    const double *f = &gm.fn[3 * mm];
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

// ---- geometric stiffness: one 3x3 block for the generalised node pair (p, pp) ---
// p, pp in 0..7: 0..3 displacement of node p, 4..7 director of node p-4.
// Returns the block already folded onto the rotation DOFs:
//   rows of a director node:    skew(fn_m) * blk      (d = theta x fn)
//   columns of a director node: blk * skew(fn_m)^T
// (TACSLinearizedRotation::addDirectorJacobian, TACSDirector.h:369-486)
A2DS_HD void geo_block(const ElemGeom &gm, const ElemWork &s, const double *Pq4, int p, int pp,
                       double out[9]) {
  const double *sig = s.sigsum;  // tying point stresses (summed over the Gauss points)
  const int m = p & 3, mm = pp & 3;
  const bool pd = p >= 4, ppd = pp >= 4;
  const double dN = (m % 2) ? 0.5 : -0.5, dM = (m / 2) ? 0.5 : -0.5;
  const double dNN = (mm % 2) ? 0.5 : -0.5, dMM = (mm / 2) ? 0.5 : -0.5;
  double sc = 0.0;  // scalar multiplying the identity (tying part)
  if (!pd && !ppd) {
    // g11: N,xi N,xi at eta = -+1 ; g22: N,eta N,eta at xi = -+1 ; g12 at the centre
    if (m / 2 == mm / 2) sc += sig[m / 2] * dN * dNN;
    if (m % 2 == mm % 2) sc += sig[2 + m % 2] * dM * dMM;
    sc += sig[4] * 0.5 * ((0.5 * dN) * (0.5 * dMM) + (0.5 * dM) * (0.5 * dNN));
  } else if (pd != ppd) {
    // g23 = 1/2 d0 . u,eta (xi = -+1), g13 = 1/2 d0 . u,xi (eta = -+1)
    const int md = pd ? m : mm, mu = pd ? mm : m;
    const double dNu = (mu % 2) ? 0.5 : -0.5, dMu = (mu / 2) ? 0.5 : -0.5;
    if (md % 2 == mu % 2) sc += sig[5 + md % 2] * 0.5 * 0.5 * dMu;
    if (md / 2 == mu / 2) sc += sig[7 + md / 2] * 0.5 * 0.5 * dNu;
  }
  double blk[9] = {sc, 0.0, 0.0, 0.0, sc, 0.0, 0.0, 0.0, sc};
  // bending part: sum_qp (alpha_p^T Sigma beta_pp + beta_p^T Sigma alpha_pp) T T^T
#pragma unroll
  for (int qp = 0; qp < 4; qp++) {
    const double *ap = s.ca[qp][p], *bp = s.cb[qp][p], *app = s.ca[qp][pp], *bpp = s.cb[qp][pp];
    const double s3 = s.sg[qp][0], s4 = s.sg[qp][1], s5 = s.sg[qp][2];
    const double mq = ap[0] * (s3 * bpp[0] + s5 * bpp[1]) + ap[1] * (s5 * bpp[0] + s4 * bpp[1]) +
                      bp[0] * (s3 * app[0] + s5 * app[1]) + bp[1] * (s5 * app[0] + s4 * app[1]);
    const double *P = Pq4 + 6 * qp;
    blk[0] += mq * P[0]; blk[1] += mq * P[1]; blk[2] += mq * P[2];
    blk[3] += mq * P[1]; blk[4] += mq * P[3]; blk[5] += mq * P[4];
    blk[6] += mq * P[2]; blk[7] += mq * P[4]; blk[8] += mq * P[5];
  }
  fold_director_block(gm, pd, m, ppd, mm, blk);
#pragma unroll
  for (int i = 0; i < 9; i++) out[i] = blk[i];
}

// ---- light geometry phases of the mass path: node normal and quadrature weight only ----
// (the same expressions as phase_node / qp_geometry evaluate for fn and det)
A2DS_HD void mass_node(ElemGeom &s, int m) {
  double Xxi[3], Xeta[3];
  edge_xi(s.X, 3, m / 2, Xxi);
  edge_eta(s.X, 3, m % 2, Xeta);
  double fn[3];
  scross(Xxi, Xeta, fn);
  double nrm = sqrt(sdot(fn, fn));
  if (nrm != 0.0) {
    double inv = 1.0 / nrm;
    fn[0] = A2DS_MUL(fn[0], inv); fn[1] = A2DS_MUL(fn[1], inv); fn[2] = A2DS_MUL(fn[2], inv);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) s.fn[3 * m + k] = fn[k];
}

A2DS_HD void mass_qp(ElemGeom &s, int qp) {
  double na[2], nb[2];
  qp_shape(qp, na, nb);
  double a0[3], a1[3], b0[3], b1[3];
  edge_vectors(s.X, 3, a0, a1, b0, b1);
  const double N[4] = {na[0] * nb[0], na[1] * nb[0], na[0] * nb[1], na[1] * nb[1]};
  double A[9];  // Xd = [X,xi | X,eta | n0] (columns), row major
#pragma unroll
  for (int k = 0; k < 3; k++) {
    A[3 * k] = nb[0] * a0[k] + nb[1] * a1[k];
    A[3 * k + 1] = na[0] * b0[k] + na[1] * b1[k];
    A[3 * k + 2] = N[0] * s.fn[k] + N[1] * s.fn[3 + k] + N[2] * s.fn[6 + k] + N[3] * s.fn[9 + k];
  }
  s.qp[qp].w = (A[8] * (A[0] * A[4] - A[3] * A[1]) - A[7] * (A[0] * A[5] - A[3] * A[2]) +
                A[6] * (A[1] * A[5] - A[2] * A[4]));
}

// ---- mass matrix: one 3x3 block for the generalised node pair (p, pp) -----------
// Kinetic energy density 1/2 (m0 u'.u' + 2 m1 u'.d' + m2 d'.d') with u and d interpolated
// bilinearly (TACSShellElement.h:614-648): the block is (sum_qp w N_m N_mm) m_k I with
// k = number of directors in the pair, folded onto the rotations like the geometric blocks.
A2DS_HD void mass_block(const CompData &c, const ElemGeom &gm, int p, int pp, double out[9]) {
  const int m = p & 3, mm = pp & 3;
  const bool pd = p >= 4, ppd = pp >= 4;
  double cc = 0.0;
#pragma unroll
  for (int qp = 0; qp < 4; qp++) {
    double na[2], nb[2];
    qp_shape(qp, na, nb);
    cc += gm.qp[qp].w * (na[m % 2] * nb[m / 2]) * (na[mm % 2] * nb[mm / 2]);
  }
  cc *= c.mom[(pd ? 1 : 0) + (ppd ? 1 : 0)];
  double blk[9] = {cc, 0.0, 0.0, 0.0, cc, 0.0, 0.0, 0.0, cc};
  fold_director_block(gm, pd, m, ppd, mm, blk);
#pragma unroll
  for (int i = 0; i < 9; i++) out[i] = blk[i];
}

}  // namespace a2ds
#endif
