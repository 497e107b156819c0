/*
 * shell_oracle.h — CPU restatement of the reference's MITC4 shell assembly path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (a2d-shells_b200/, include/,
 * bench.py's GPU arm) may link, import or call this.  Allowed callers: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
 * and there only as the checker.
 *
 * Parity status: PINNED against the reference itself — the unmodified reference
 * is compiled into oracle/_ref (oracle/Makefile) and tests/test_oracle.py checks
 * every function here against it on seeded inputs, plus against the fixtures in
 * tests/golden/ that were generated from it (tests/golden/make_golden.py).
 * The reference ships no golden vectors or known-answer tests of its own
 * (SURVEY.md §4).
 */
#ifndef A2DS_SHELL_ORACLE_H
#define A2DS_SHELL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int model;          /* 0 = TACSShellLinearModel, 1 = TACSShellNonlinearModel */
  int transform;      /* 0 = TACSShellNaturalTransform, 1 = TACSShellRefAxisTransform */
  double axis[3];     /* normalised reference axis (transform == 1) */
  double Cs[22];      /* A[6] B[6] D[6] As[3] drill  (TACSShellConstitutive.h:34) */
  double eth[9];      /* thermal strain per unit temperature */
  double mom[3];      /* mass moments (TACSShellConstitutive::evalMassMoments) */
  double temperature; /* TACSShellElement::temperature (TACSShellElement.h:35) */
} oracle_comp_t;

/* Array sizes below are for the 4-node element (order 2): X 12, q 24, res 24, mat 576, e 36,
   detXd 4.  The same entry points compiled for the 9-node element (shell_oracle_q9.c) are
   called oracle9_*: X 27, q 54, res 54, mat 2916 (54 x 54), e 81, detXd 9, conn 9 per element. */
/* strains e[4][9] and detXd*weight at the 4 Gauss points for state q[24] */
void oracle_strain(const oracle_comp_t *c, const double *X, const double *q,
                   double *e, double *detXd);

/* TACSShellElement::addResidual (static terms): res[24] (overwritten) */
void oracle_residual(const oracle_comp_t *c, const double *X, const double *q,
                     double *res);

/* TACSShellElement::addJacobian with beta = gamma = 0: res[24], mat[576] = alpha*dR/dq
   (both overwritten; either may be NULL) */
void oracle_jacobian(const oracle_comp_t *c, double alpha, const double *X,
                     const double *q, double *res, double *mat);

/* the same with the inertial terms: res += M qdd, mat += gamma * M
   (TACSShellElement.h:410-447, 614-648; qdd may be NULL = zero) */
void oracle_jacobian_dyn(const oracle_comp_t *c, double alpha, double gamma, const double *X,
                         const double *q, const double *qdd, double *res,
                         double *mat);

/* TACSShellElement::getMatType: type 0 = stiffness, 1 = geometric stiffness, 2 = mass */
void oracle_mat_type(const oracle_comp_t *c, int type, const double *X,
                     const double *q, double *mat);

/* Non-zero pattern of the node-to-node matrix, columns sorted per row
   (TACSAssembler::computeLocalNodeToNodeCSR).  Two-pass: call with cols == NULL
   to fill rowp[n_nodes+1] and get the block count, then again with cols. */
int oracle_pattern(int n_nodes, int n_elems, const int *conn, int *rowp, int *cols);

/*
 * Assembly over a mesh (TACSAssembler::assembleRes / assembleJacobian /
 * assembleMatType, single rank): op 0 = residual, 1 = Jacobian (res + alpha*K),
 * 2 = matType(K), 3 = matType(G), 4 = matType(M).  res[6*n_nodes] and A[36*nnz] are zeroed first
 * (may be NULL when the op does not produce them); boundary conditions are
 * applied exactly as the reference does (TACSBVec::applyBCs, BCSRMat::zeroRow).
 * Returns 0, or the number of element blocks not found in the pattern.
 */
int oracle_assemble(int op, double alpha, int n_nodes, int n_elems, const int *conn,
                    const int *elem_comp, const oracle_comp_t *comps, const double *X,
                    const double *u, int n_bc, const int *bc_nodes, const int *bc_vars,
                    const double *bc_vals, const int *rowp, const int *cols, double *res,
                    double *A);

/* the same with second time derivatives udd (may be NULL) and the gamma term of the
   Jacobian (op 1): res includes M udd, A = alpha K + gamma M */
int oracle_assemble_dyn(int op, double alpha, double gamma, int n_nodes, int n_elems,
                        const int *conn, const int *elem_comp, const oracle_comp_t *comps,
                        const double *X, const double *u, const double *udd, int n_bc,
                        const int *bc_nodes, const int *bc_vars, const double *bc_vals,
                        const int *rowp, const int *cols, double *res, double *A);

/* order 3 (TACSQuad9Shell / TACSQuad9NonlinearShell): same semantics, sizes as stated above */
void oracle9_strain(const oracle_comp_t *c, const double *X, const double *q, double *e, double *detXd);
void oracle9_residual(const oracle_comp_t *c, const double *X, const double *q, double *res);
void oracle9_jacobian(const oracle_comp_t *c, double alpha, const double *X, const double *q,
                      double *res, double *mat);
void oracle9_jacobian_dyn(const oracle_comp_t *c, double alpha, double gamma, const double *X,
                          const double *q, const double *qdd, double *res, double *mat);
void oracle9_mat_type(const oracle_comp_t *c, int type, const double *X, const double *q, double *mat);
int oracle9_pattern(int n_nodes, int n_elems, const int *conn, int *rowp, int *cols);
int oracle9_assemble(int op, double alpha, int n_nodes, int n_elems, const int *conn,
                     const int *elem_comp, const oracle_comp_t *comps, const double *X,
                     const double *u, int n_bc, const int *bc_nodes, const int *bc_vars,
                     const double *bc_vals, const int *rowp, const int *cols, double *res,
                     double *A);
int oracle9_assemble_dyn(int op, double alpha, double gamma, int n_nodes, int n_elems,
                         const int *conn, const int *elem_comp, const oracle_comp_t *comps,
                         const double *X, const double *u, const double *udd, int n_bc,
                         const int *bc_nodes, const int *bc_vars, const double *bc_vals,
                         const int *rowp, const int *cols, double *res, double *A);

/* Meshes with dependent nodes (TACSAssembler::setDependentNodes, src/TACSAssembler.cpp:716-775):
   a connectivity entry -(d + 1) is dependent node d = sum_j dep_w[j] * node dep_conn[j],
   j in dep_ptr[d] .. dep_ptr[d + 1].  Pattern over the independent nodes behind every element
   (computeLocalNodeToNodeCSR :1850-1935), assembly with the weights (TACSBVec.cpp:855-885,
   :930-975; TACSAssembler.h:485-510).  Vectors have n_nodes (independent) rows. */
int oracle_pattern_dep(int n_nodes, int n_elems, const int *conn, const int *dep_ptr,
                       const int *dep_conn, int *rowp, int *cols);
int oracle_assemble_dep(int op, double alpha, double gamma, int n_nodes, int n_elems,
                        const int *conn, const int *elem_comp, const oracle_comp_t *comps,
                        const double *X, const double *u, const double *udd, int n_dep,
                        const int *dep_ptr, const int *dep_conn, const double *dep_w, int n_bc,
                        const int *bc_nodes, const int *bc_vars, const double *bc_vals,
                        const int *rowp, const int *cols, double *res, double *A);
int oracle9_pattern_dep(int n_nodes, int n_elems, const int *conn, const int *dep_ptr,
                        const int *dep_conn, int *rowp, int *cols);
int oracle9_assemble_dep(int op, double alpha, double gamma, int n_nodes, int n_elems,
                         const int *conn, const int *elem_comp, const oracle_comp_t *comps,
                         const double *X, const double *u, const double *udd, int n_dep,
                         const int *dep_ptr, const int *dep_conn, const double *dep_w, int n_bc,
                         const int *bc_nodes, const int *bc_vars, const double *bc_vals,
                         const int *rowp, const int *cols, double *res, double *A);

#ifdef __cplusplus
}
#endif
#endif
