/*
 * ref_driver.cpp — flat C interface onto the UNMODIFIED reference build
 * (oracle/_ref/liba2dshells_ref.so).  TEST INFRASTRUCTURE ONLY.
 *
 * Used by: tests/ (checker), tests/golden/make_golden.py (fixture generator),
 * bench.py --impl reference / cpu_baseline (the CPU arm that is timed beside the
 * GPU path).  Never linked into, imported by, or called from the product.
 *
 * Everything here calls the reference's own public API:
 *   TACSCreator::setGlobalConnectivity/…/createTACS   src/TACSCreator.h:45-110
 *   TACSAssembler::assembleRes/assembleJacobian/assembleMatType
 *                                                     src/TACSAssembler.h:213-220
 *   TACSElement::addResidual/addJacobian/getMatType   src/elements/TACSElement.h:422,450,526
 *   TACSLinearBuckling::solve                         src/TACSBuckling.cpp:197
 * It is compiled with -fno-access-control only so that the buckling helper can
 * re-run the tail of TACSLinearBuckling::solve (src/TACSBuckling.cpp:262-277)
 * on externally supplied K and G values.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "KSM.h"
#include "TACSAssembler.h"
#include "TACSBuckling.h"
#include "TACSCreator.h"
#include "TACSIsoShellConstitutive.h"
#include "TACSMaterialProperties.h"
#include "TACSMeshLoader.h"
#include "TACSParallelMat.h"
#include "TACSSchurMat.h"
#include "TACSShellElementDefs.h"
#include "TACSShellElementTransform.h"

#define NPROP 40

/* A shell constitutive object with an arbitrary (point-independent) 22-entry
   tangent: stands in for the archived composite / blade-stiffened classes
   (archive/constitutive/*, not compiled by the reference) — same interface,
   TACSShellConstitutive.h:31-147. */
class GeneralShellCon : public TACSShellConstitutive {
 public:
  GeneralShellCon(const double *p) {
    memcpy(Cs, &p[9], 22 * sizeof(double));
    memcpy(eth, &p[31], 3 * sizeof(double));
    memcpy(mom, &p[34], 3 * sizeof(double));
  }
  TacsScalar evalDensity(int, const double[], const TacsScalar[]) { return mom[0]; }
  TacsScalar evalSpecificHeat(int, const double[], const TacsScalar[]) { return 0.0; }
  void evalMassMoments(int, const double[], const TacsScalar[], TacsScalar m[]) {
    m[0] = mom[0]; m[1] = mom[1]; m[2] = mom[2];
  }
  void evalStress(int, const double[], const TacsScalar[], const TacsScalar e[], TacsScalar s[]) {
    TacsScalar drill;
    const TacsScalar *A, *B, *D, *As;
    extractTangentStiffness(Cs, &A, &B, &D, &As, &drill);
    computeStress(A, B, D, As, drill, e, s);
  }
  void evalTangentStiffness(int, const double[], const TacsScalar[], TacsScalar C[]) {
    memcpy(C, Cs, 22 * sizeof(TacsScalar));
  }
  void evalThermalStrain(int, const double[], const TacsScalar[], TacsScalar theta, TacsScalar e[]) {
    e[0] = eth[0] * theta; e[1] = eth[1] * theta; e[2] = eth[2] * theta;
    e[3] = e[4] = e[5] = e[6] = e[7] = e[8] = 0.0;
  }
  double Cs[22], eth[3], mom[3];
};

static TACSShellTransform *make_transform(int kind, const double *axis) {
  if (kind == 1) return new TACSShellRefAxisTransform(axis);
  return new TACSShellNaturalTransform();
}

static TACSShellConstitutive *make_con(const double *p) {
  if ((int)p[1] == 1) return new GeneralShellCon(p);
  TACSMaterialProperties *mat =
      new TACSMaterialProperties(p[5], 0.0, p[3], p[4], 1e11, p[6], 0.0);
  return new TACSIsoShellConstitutive(mat, p[7], -1, 0.0, 1.0, p[8]);
}

/* temperature lives on the element object (TACSShellElement.h:35,76) */
static TACSElement *make_element(const double *p, TACSShellTransform *tr) {
  TACSShellConstitutive *con = make_con(p);
  /* p[0]: 0 TACSQuad4Shell, 1 TACSQuad4NonlinearShell, 2 TACSQuad9Shell,
     3 TACSQuad9NonlinearShell (TACSShellElementDefs.h:12-37) */
  if ((int)p[0] == 1) {
    TACSQuad4NonlinearShell *e = new TACSQuad4NonlinearShell(tr, con);
    e->setTemperature(p[2]);
    return e;
  }
  if ((int)p[0] == 2) {
    TACSQuad9Shell *e = new TACSQuad9Shell(tr, con);
    e->setTemperature(p[2]);
    return e;
  }
  if ((int)p[0] == 3) {
    TACSQuad9NonlinearShell *e = new TACSQuad9NonlinearShell(tr, con);
    e->setTemperature(p[2]);
    return e;
  }
  TACSQuad4Shell *e = new TACSQuad4Shell(tr, con);
  e->setTemperature(p[2]);
  return e;
}

struct RefMat {
  int kind;  // 0 = TACSParallelMat, 1 = TACSSchurMat
  TACSMat *mat;
  BCSRMat *blk[4];
};

struct RefCtx {
  TACSCreator *creator;
  TACSAssembler *assembler;
  int n_nodes, n_elems;
  std::vector<int> new_nodes;  // original id -> reference numbering
  std::vector<RefMat> mats;
  TACSBVec *res, *u, *ud, *udd;
};

extern "C" {

int refdrv_nprop() { return NPROP; }

/* Evaluate the constitutive tables exactly as the reference does on the host:
   Cs[22] (TACSIsoShellConstitutive.cpp:192-226), unit thermal strain eth[9]
   (:438-456, theta = 1) and mass moments[3] (:120-129). */
int refdrv_con_tables(const double *p, double *Cs, double *eth, double *mom) {
  TACSShellConstitutive *con = make_con(p);
  con->incref();
  double pt[3] = {0, 0, 0}, X[3] = {0, 0, 0};
  con->evalTangentStiffness(0, pt, X, Cs);
  con->evalThermalStrain(0, pt, X, 1.0, eth);
  con->evalMassMoments(0, pt, X, mom);
  con->decref();
  return 0;
}

/* One element, one call: op 0 addResidual, 1 addJacobian, 2 getMatType(K),
   3 getMatType(G), 4 getMatType(M).  res/mat are zeroed here first (the
   assembler does the same, TACSAssembler.cpp:4144-4145). */
int refdrv_element(const double *p, int transform_kind, const double *axis, int op,
                   double alpha, double beta, double gamma, const double *X,
                   const double *vars, const double *dvars, const double *ddvars,
                   double *res, double *mat) {
  TACSShellTransform *tr = make_transform(transform_kind, axis);
  tr->incref();
  TACSElement *e = make_element(p, tr);
  e->incref();
  const int nv = e->getNumVariables();   /* 24 (4 nodes) or 54 (9 nodes) */
  if (res) memset(res, 0, nv * sizeof(double));
  if (mat) memset(mat, 0, nv * nv * sizeof(double));
  switch (op) {
    case 0: e->addResidual(0, 0.0, X, vars, dvars, ddvars, res); break;
    case 1: e->addJacobian(0, 0.0, alpha, beta, gamma, X, vars, dvars, ddvars, res, mat); break;
    case 2: e->getMatType(TACS_STIFFNESS_MATRIX, 0, 0.0, X, vars, mat); break;
    case 3: e->getMatType(TACS_GEOMETRIC_STIFFNESS_MATRIX, 0, 0.0, X, vars, mat); break;
    case 4: e->getMatType(TACS_MASS_MATRIX, 0, 0.0, X, vars, mat); break;
    default: return 1;
  }
  e->decref();
  tr->decref();
  return 0;
}

/* Same as refdrv_element for a batch of n elements sharing one component
   (X: n*12, vars: n*24, out res: n*24, mat: n*576); returns seconds spent in
   the element calls only. */
double refdrv_element_batch(const double *p, int transform_kind, const double *axis, int op,
                            double alpha, double beta, double gamma, int n, const double *X,
                            const double *vars, double *res, double *mat) {
  TACSShellTransform *tr = make_transform(transform_kind, axis);
  tr->incref();
  TACSElement *e = make_element(p, tr);
  e->incref();
  double zero[24];
  memset(zero, 0, sizeof(zero));
  double lres[24], lmat[576];
  double t0 = MPI_Wtime();
  for (int i = 0; i < n; i++) {
    double *r = res ? &res[24 * i] : lres;
    double *m = mat ? &mat[576 * (size_t)i] : lmat;
    memset(r, 0, 24 * sizeof(double));
    memset(m, 0, 576 * sizeof(double));
    const double *Xe = &X[12 * i], *ve = &vars[24 * i];
    switch (op) {
      case 0: e->addResidual(i, 0.0, Xe, ve, zero, zero, r); break;
      case 1: e->addJacobian(i, 0.0, alpha, beta, gamma, Xe, ve, zero, zero, r, m); break;
      case 2: e->getMatType(TACS_STIFFNESS_MATRIX, i, 0.0, Xe, ve, m); break;
      case 3: e->getMatType(TACS_GEOMETRIC_STIFFNESS_MATRIX, i, 0.0, Xe, ve, m); break;
      case 4: e->getMatType(TACS_MASS_MATRIX, i, 0.0, Xe, ve, m); break;
    }
  }
  double t1 = MPI_Wtime();
  e->decref();
  tr->decref();
  return t1 - t0;
}

/* Build a reference TACSAssembler for a quad mesh given in the caller's
   ("original") numbering, through TACSCreator exactly as
   TACSMeshLoader::createTACS does (src/io/TACSMeshLoader.cpp:1130-1184):
   NATURAL_ORDER / DIRECT_SCHUR defaults (TACSMeshLoader.h:75-78). */
void *refdrv_create_n(int nodes_per_elem, int n_nodes, int n_elems, const int *conn,
                      const int *elem_comp, const double *X, int n_bc, const int *bc_nodes,
                      const int *bc_ptr, const int *bc_vars, const double *bc_vals, int n_comp,
                      const double *comp_props, int transform_kind, const double *axis);
/* dependent nodes for the NEXT refdrv_create / refdrv_create_n call
   (TACSCreator::setDependentNodes, src/TACSCreator.h:65-67): connectivity entries -(d + 1) */
static std::vector<int> g_dep_ptr, g_dep_conn;
static std::vector<double> g_dep_w;
int refdrv_next_dependent_nodes(int n_dep, const int *dep_ptr, const int *dep_conn,
                                const double *dep_w) {
  g_dep_ptr.assign(dep_ptr, dep_ptr + n_dep + 1);
  g_dep_conn.assign(dep_conn, dep_conn + dep_ptr[n_dep]);
  g_dep_w.assign(dep_w, dep_w + dep_ptr[n_dep]);
  return 0;
}
void *refdrv_create(int n_nodes, int n_elems, const int *conn, const int *elem_comp,
                    const double *X, int n_bc, const int *bc_nodes, const int *bc_ptr,
                    const int *bc_vars, const double *bc_vals, int n_comp,
                    const double *comp_props, int transform_kind, const double *axis) {
  return refdrv_create_n(4, n_nodes, n_elems, conn, elem_comp, X, n_bc, bc_nodes, bc_ptr, bc_vars,
                         bc_vals, n_comp, comp_props, transform_kind, axis);
}
/* the same for elements of nodes_per_elem nodes (9: TACSQuad9Shell, node order of
   TACSShellQuadBasis<3>::getNodePoint, xi fastest) */
void *refdrv_create_n(int nodes_per_elem, int n_nodes, int n_elems, const int *conn,
                      const int *elem_comp, const double *X, int n_bc, const int *bc_nodes,
                      const int *bc_ptr, const int *bc_vars, const double *bc_vals, int n_comp,
                      const double *comp_props, int transform_kind, const double *axis) {
  if (!TacsIsInitialized()) {
    MPI_Init(NULL, NULL);
    TacsInitialize();
  }
  RefCtx *c = new RefCtx();
  c->n_nodes = n_nodes;
  c->n_elems = n_elems;
  MPI_Comm comm = MPI_COMM_WORLD;
  c->creator = new TACSCreator(comm, 6);
  c->creator->incref();
  c->creator->setReorderingType(TACSAssembler::NATURAL_ORDER, TACSAssembler::DIRECT_SCHUR);

  std::vector<int> ptr(n_elems + 1);
  for (int i = 0; i <= n_elems; i++) ptr[i] = nodes_per_elem * i;
  c->creator->setGlobalConnectivity(n_nodes, n_elems, ptr.data(), conn, elem_comp);
  c->creator->setBoundaryConditions(n_bc, bc_nodes, bc_ptr, bc_vars, bc_vals);
  if (g_dep_ptr.size() > 1) {
    c->creator->setDependentNodes((int)g_dep_ptr.size() - 1, g_dep_ptr.data(), g_dep_conn.data(),
                                  g_dep_w.data());
    g_dep_ptr.clear(); g_dep_conn.clear(); g_dep_w.clear();
  }
  c->creator->setNodes(X);

  TACSShellTransform *tr = make_transform(transform_kind, axis);
  tr->incref();
  std::vector<TACSElement *> elems(n_comp);
  for (int i = 0; i < n_comp; i++) {
    elems[i] = make_element(&comp_props[NPROP * i], tr);
    elems[i]->incref();
  }
  c->creator->setElements(n_comp, elems.data());
  c->assembler = c->creator->createTACS();
  c->assembler->incref();

  const int *nn;
  c->creator->getNodeNums(&nn);
  c->new_nodes.assign(nn, nn + n_nodes);

  c->res = c->assembler->createVec(); c->res->incref();
  c->u = c->assembler->createVec(); c->u->incref();
  c->ud = c->assembler->createVec(); c->ud->incref();
  c->udd = c->assembler->createVec(); c->udd->incref();
  return c;
}

void refdrv_destroy(void *h) {
  RefCtx *c = (RefCtx *)h;
  for (size_t i = 0; i < c->mats.size(); i++) c->mats[i].mat->decref();
  c->res->decref(); c->u->decref(); c->ud->decref(); c->udd->decref();
  c->assembler->decref();
  c->creator->decref();
  delete c;
}

int refdrv_get_node_nums(void *h, int *out) {
  RefCtx *c = (RefCtx *)h;
  memcpy(out, c->new_nodes.data(), c->n_nodes * sizeof(int));
  return 0;
}

/* element connectivity and node locations in the reference's own numbering */
int refdrv_get_conn(void *h, int *conn) {
  RefCtx *c = (RefCtx *)h;
  const int *ptr, *cn;
  c->assembler->getElementConnectivity(&ptr, &cn);
  memcpy(conn, cn, (size_t)ptr[c->n_elems] * sizeof(int));
  return 0;
}

/* dependent nodes as the assembler holds them (reference numbering); returns their number */
int refdrv_get_dep(void *h, int *ptr, int *conn, double *w) {
  RefCtx *c = (RefCtx *)h;
  TACSBVecDepNodes *d = c->assembler->getBVecDepNodes();
  if (!d) return 0;
  const int *dp, *dc;
  const double *dw;
  int nd = d->getDepNodes(&dp, &dc, &dw);
  if (ptr) memcpy(ptr, dp, (nd + 1) * sizeof(int));
  if (conn) memcpy(conn, dc, dp[nd] * sizeof(int));
  if (w) memcpy(w, dw, dp[nd] * sizeof(double));
  return nd;
}

int refdrv_get_nodes(void *h, double *X) {
  RefCtx *c = (RefCtx *)h;
  TACSBVec *xv;
  c->assembler->getNodes(&xv);
  TacsScalar *x;
  int n = xv->getArray(&x);
  memcpy(X, x, n * sizeof(double));
  return n;
}

int refdrv_get_bcs(void *h, int *nodes, int *vars, double *vals) {
  RefCtx *c = (RefCtx *)h;
  const int *n, *v;
  TacsScalar *val;
  int nb = c->assembler->getBcMap()->getBCs(&n, &v, &val);
  if (nodes) memcpy(nodes, n, nb * sizeof(int));
  if (vars) memcpy(vars, v, nb * sizeof(int));
  if (vals) memcpy(vals, val, 6 * (size_t)nb * sizeof(double));
  return nb;
}

int refdrv_set_threads(void *h, int nt) {
  ((RefCtx *)h)->assembler->setNumThreads(nt);
  return 0;
}

/* state in the reference's numbering; NULL = leave zero */
int refdrv_set_state(void *h, const double *u, const double *ud, const double *udd) {
  RefCtx *c = (RefCtx *)h;
  TacsScalar *a;
  int n = c->u->getArray(&a);
  if (u) memcpy(a, u, n * sizeof(double)); else memset(a, 0, n * sizeof(double));
  c->ud->getArray(&a);
  if (ud) memcpy(a, ud, n * sizeof(double)); else memset(a, 0, n * sizeof(double));
  c->udd->getArray(&a);
  if (udd) memcpy(a, udd, n * sizeof(double)); else memset(a, 0, n * sizeof(double));
  c->assembler->setVariables(c->u, c->ud, c->udd);
  return n;
}

/* set the per-element temperature on every element object (shared per component) */
int refdrv_set_temperature(void *h, double T) {
  RefCtx *c = (RefCtx *)h;
  int ne = c->assembler->getNumElements();
  for (int i = 0; i < ne; i++) {
    TACSElement *e = c->assembler->getElement(i);
    TACSQuad4Shell *l = dynamic_cast<TACSQuad4Shell *>(e);
    if (l) l->setTemperature(T);
    TACSQuad4NonlinearShell *nl = dynamic_cast<TACSQuad4NonlinearShell *>(e);
    if (nl) nl->setTemperature(T);
    TACSQuad9Shell *l9 = dynamic_cast<TACSQuad9Shell *>(e);
    if (l9) l9->setTemperature(T);
    TACSQuad9NonlinearShell *nl9 = dynamic_cast<TACSQuad9NonlinearShell *>(e);
    if (nl9) nl9->setTemperature(T);
  }
  return 0;
}

int refdrv_mat_create(void *h, int kind) {
  RefCtx *c = (RefCtx *)h;
  RefMat m;
  m.kind = kind == 2 ? 1 : kind;   /* both Schur flavours */
  m.blk[0] = m.blk[1] = m.blk[2] = m.blk[3] = NULL;
  if (kind == 1 || kind == 2) {
    /* 1: as the shipped examples call it (TACS_AMD_ORDER, mechBuckling.cpp:118-120); 2: the same
       matrix class with the local nodes left in their natural order — no AMD pass over the
       interior block, which at 1 M nodes takes minutes and is set-up, not assembly */
    TACSSchurMat *s = kind == 1 ? c->assembler->createSchurMat()
                                : c->assembler->createSchurMat(TACSAssembler::NATURAL_ORDER);
    s->incref();
    s->getBCSRMat(&m.blk[0], &m.blk[1], &m.blk[2], &m.blk[3]);
    m.mat = s;
  } else {
    TACSParallelMat *p = c->assembler->createMat();
    p->incref();
    p->getBCSRMat(&m.blk[0], &m.blk[1]);
    m.mat = p;
  }
  c->mats.push_back(m);
  return (int)c->mats.size() - 1;
}

/* sizes of BCSR block `which` (ParallelMat: 0=Aloc 1=Bext; SchurMat: 0=B 1=E 2=F 3=C) */
int refdrv_mat_info(void *h, int mat, int which, int *nrows, int *ncols, int *nnz) {
  RefCtx *c = (RefCtx *)h;
  BCSRMat *b = c->mats[mat].blk[which];
  if (!b) { *nrows = *ncols = *nnz = 0; return 1; }
  int bs;
  const int *rowp, *cols;
  TacsScalar *A;
  b->getArrays(&bs, nrows, ncols, &rowp, &cols, &A);
  *nnz = rowp[*nrows];
  return 0;
}

int refdrv_mat_get(void *h, int mat, int which, int *rowp_out, int *cols_out, double *A_out) {
  RefCtx *c = (RefCtx *)h;
  BCSRMat *b = c->mats[mat].blk[which];
  if (!b) return 1;
  int bs, nr, nc;
  const int *rowp, *cols;
  TacsScalar *A;
  b->getArrays(&bs, &nr, &nc, &rowp, &cols, &A);
  if (rowp_out) memcpy(rowp_out, rowp, (nr + 1) * sizeof(int));
  if (cols_out) memcpy(cols_out, cols, rowp[nr] * sizeof(int));
  if (A_out) memcpy(A_out, A, 36 * (size_t)rowp[nr] * sizeof(double));
  return 0;
}

int refdrv_mat_set(void *h, int mat, int which, const double *A_in) {
  RefCtx *c = (RefCtx *)h;
  BCSRMat *b = c->mats[mat].blk[which];
  if (!b) return 1;
  int bs, nr, nc;
  const int *rowp, *cols;
  TacsScalar *A;
  b->getArrays(&bs, &nr, &nc, &rowp, &cols, &A);
  memcpy(A, A_in, 36 * (size_t)rowp[nr] * sizeof(double));
  return 0;
}

/* SchurMat only: the node (reference numbering) behind each local row of B
   (which=0) / C (which=1) — TACSSchurMat.cpp:453-531 routes by these sets. */
int refdrv_schur_index(void *h, int mat, int which, int *out) {
  RefCtx *c = (RefCtx *)h;
  if (c->mats[mat].kind != 1) return -1;
  TACSSchurMat *s = (TACSSchurMat *)c->mats[mat].mat;
  TACSBVecIndices *idx = which ? s->getSchurMap()->getIndices() : s->getLocalMap()->getIndices();
  const int *ind;
  int n = idx->getIndices(&ind);
  if (out) memcpy(out, ind, n * sizeof(int));
  return n;
}

int refdrv_assemble_res(void *h, double *res) {
  RefCtx *c = (RefCtx *)h;
  c->assembler->assembleRes(c->res);
  TacsScalar *a;
  int n = c->res->getArray(&a);
  if (res) memcpy(res, a, n * sizeof(double));
  return n;
}

int refdrv_assemble_jacobian(void *h, double alpha, double beta, double gamma, int mat, double *res) {
  RefCtx *c = (RefCtx *)h;
  c->assembler->assembleJacobian(alpha, beta, gamma, c->res, c->mats[mat].mat);
  TacsScalar *a;
  int n = c->res->getArray(&a);
  if (res) memcpy(res, a, n * sizeof(double));
  return n;
}

/* the same with TACS_MAT_TRANSPOSE (element matrices transposed on their way into A) */
int refdrv_assemble_jacobian_transpose(void *h, double alpha, double beta, double gamma, int mat,
                                       double *res) {
  RefCtx *c = (RefCtx *)h;
  c->assembler->assembleJacobian(alpha, beta, gamma, c->res, c->mats[mat].mat, TACS_MAT_TRANSPOSE);
  TacsScalar *a;
  int n = c->res->getArray(&a);
  if (res) memcpy(res, a, n * sizeof(double));
  return n;
}

/* type: 0 K, 1 G, 2 M */
int refdrv_assemble_mat_type(void *h, int type, int mat) {
  RefCtx *c = (RefCtx *)h;
  ElementMatrixType t = TACS_STIFFNESS_MATRIX;
  if (type == 1) t = TACS_GEOMETRIC_STIFFNESS_MATRIX;
  if (type == 2) t = TACS_MASS_MATRIX;
  c->assembler->assembleMatType(t, c->mats[mat].mat);
  return 0;
}

/* wall seconds of one call; op 0 res, 1 jacobian(alpha=1), 2 K, 3 G */
double refdrv_time(void *h, int op, int mat) {
  RefCtx *c = (RefCtx *)h;
  double t0 = MPI_Wtime();
  if (op == 0) c->assembler->assembleRes(c->res);
  else if (op == 1) c->assembler->assembleJacobian(1.0, 0.0, 0.0, c->res, c->mats[mat].mat);
  else if (op == 2) c->assembler->assembleMatType(TACS_STIFFNESS_MATRIX, c->mats[mat].mat);
  else if (op == 3) c->assembler->assembleMatType(TACS_GEOMETRIC_STIFFNESS_MATRIX, c->mats[mat].mat);
  return MPI_Wtime() - t0;
}

/*
 * Linear buckling with the reference solver stack exactly as
 * examples/cylinder-buckling/mechBuckling.cpp:118-153 sets it up
 * (TACSSchurPc(kmat, 1e6, 10, 1), GMRES(aux, pc, 10, 15), tolerances 1e-12).
 * mode 0: TACSLinearBuckling::solve does everything (reference assembly).
 * mode 1: kmat and gmat already hold externally assembled values (set with
 *         refdrv_mat_set); only the tail of solve() is run
 *         (src/TACSBuckling.cpp:240,262-263,269-277).
 * u0 (reference numbering) is the load-path state, as the shipped examples pass;
 * NULL in mode 0 lets the reference solve for the path.  path_out (optional)
 * receives the load path actually used (after setBCs).
 */
int refdrv_buckling(void *h, int kmat, int gmat, int aux, int mode, double sigma,
                    int max_lanczos, int num_eigs, double tol, const double *u0,
                    double *eigs, double *errs, double *path_out) {
  RefCtx *c = (RefCtx *)h;
  TACSSchurMat *K = (TACSSchurMat *)c->mats[kmat].mat;
  TACSSchurMat *G = (TACSSchurMat *)c->mats[gmat].mat;
  TACSSchurMat *A = (TACSSchurMat *)c->mats[aux].mat;
  TACSSchurPc *pc = new TACSSchurPc(K, 1000000, 10.0, 1);
  pc->incref();
  GMRES *solver = new GMRES(A, pc, 10, 15, 0);
  solver->incref();
  solver->setTolerances(1e-12, 1e-12);
  TACSLinearBuckling *b =
      new TACSLinearBuckling(c->assembler, sigma, G, K, A, solver, max_lanczos, num_eigs, tol);
  b->incref();
  TACSBVec *f = c->assembler->createVec(); f->incref();
  TACSBVec *u = c->assembler->createVec(); u->incref();
  if (u0) {
    TacsScalar *a;
    int n = u->getArray(&a);
    memcpy(a, u0, n * sizeof(double));
  }
  if (mode == 0) {
    /* u0 == NULL: the reference computes the load path itself by a linear static
       solve, path = -K^{-1} r (src/TACSBuckling.cpp:244-258) */
    b->solve(f, u0 ? u : NULL, NULL);
  } else {
    c->assembler->zeroVariables();
    A->copyValues(K);
    b->path->copyValues(u);
    c->assembler->setBCs(b->path);
    c->assembler->setVariables(b->path);
    A->axpy(sigma, G);
    A->applyBCs(c->assembler->getBcMap());
    b->pc->factor();
    b->sep->solve(NULL);
  }
  if (path_out) {
    TacsScalar *a;
    int n = b->path->getArray(&a);
    memcpy(path_out, a, n * sizeof(double));
  }
  for (int i = 0; i < num_eigs; i++) {
    TacsScalar err;
    eigs[i] = b->extractEigenvalue(i, &err);
    if (errs) errs[i] = err;
  }
  f->decref(); u->decref();
  b->decref(); solver->decref(); pc->decref();
  return 0;
}

/* TACSFrequencyAnalysis (Lanczos, shift-invert): the reference assembles K and M through
   TACSAssembler::assembleMatType, forms K - sigma M, applies the BCs, factors and iterates
   (src/TACSBuckling.cpp:823-870).  Returns the eigenvalues omega^2. */
int refdrv_frequency(void *h, int kmat, int mmat, double sigma, int max_lanczos, int num_eigs,
                     double tol, double *eigs, double *errs) {
  RefCtx *c = (RefCtx *)h;
  TACSSchurMat *K = (TACSSchurMat *)c->mats[kmat].mat;
  TACSSchurMat *M = (TACSSchurMat *)c->mats[mmat].mat;
  TACSSchurPc *pc = new TACSSchurPc(K, 1000000, 10.0, 1);
  pc->incref();
  GMRES *solver = new GMRES(K, pc, 10, 15, 0);
  solver->incref();
  solver->setTolerances(1e-12, 1e-12);
  TACSFrequencyAnalysis *f =
      new TACSFrequencyAnalysis(c->assembler, sigma, M, K, solver, max_lanczos, num_eigs, tol);
  f->incref();
  f->solve(NULL, 0);
  for (int i = 0; i < num_eigs; i++) {
    TacsScalar err;
    eigs[i] = f->extractEigenvalue(i, &err);
    if (errs) errs[i] = err;
  }
  f->decref(); solver->decref(); pc->decref();
  return 0;
}

/* load-path state the buckling flow assembles G about: zero + setBCs
   (src/TACSBuckling.cpp:262) in the reference numbering */
int refdrv_bc_state(void *h, double *u) {
  RefCtx *c = (RefCtx *)h;
  TACSBVec *v = c->assembler->createVec();
  v->incref();
  c->assembler->setBCs(v);
  TacsScalar *a;
  int n = v->getArray(&a);
  memcpy(u, a, n * sizeof(double));
  v->decref();
  return n;
}

/* The reference's BDF reader (TACSMeshLoader::scanBDFFile, src/io/TACSMeshLoader.cpp:570):
   scan, then hand out what getConnectivity / getBCs return.  Two-call protocol: with
   elem_ptr == NULL only the sizes are filled.  Returns scanBDFFile's own return value;
   the loader stays alive in *h until refdrv_bdf_free.  secs: wall time of the scan. */
int refdrv_bdf_scan(const char *path, void **h, int *sizes, double *secs) {
  TACSMeshLoader *ml = new TACSMeshLoader(MPI_COMM_WORLD);
  ml->incref();
  double t0 = MPI_Wtime();
  int fail = ml->scanBDFFile(path);
  *secs = MPI_Wtime() - t0;
  int nn = 0, ne = 0, nb = 0;
  const int *eptr = NULL, *bptr = NULL;
  ml->getConnectivity(&nn, &ne, &eptr, NULL, NULL, NULL);
  ml->getBCs(&nb, NULL, NULL, &bptr, NULL);
  sizes[0] = nn; sizes[1] = ne; sizes[2] = (eptr ? eptr[ne] : 0);
  sizes[3] = nb; sizes[4] = (bptr ? bptr[nb] : 0); sizes[5] = ml->getNumComponents();
  *h = ml;
  return fail;
}

void refdrv_bdf_arrays(void *h, int *elem_ptr, int *elem_conn, int *elem_comp, double *X,
                       int *bc_nodes, int *bc_ptr, int *bc_vars, double *bc_vals,
                       char *elem_descript /* 9 per component */,
                       char *comp_descript /* 33 per component */) {
  TACSMeshLoader *ml = (TACSMeshLoader *)h;
  int nn, ne, nb;
  const int *eptr, *econn, *ecomp, *bn, *bv, *bp;
  const TacsScalar *Xp, *bvals;
  ml->getConnectivity(&nn, &ne, &eptr, &econn, &ecomp, &Xp);
  ml->getBCs(&nb, &bn, &bv, &bp, &bvals);
  memcpy(elem_ptr, eptr, (ne + 1) * sizeof(int));
  memcpy(elem_conn, econn, eptr[ne] * sizeof(int));
  memcpy(elem_comp, ecomp, ne * sizeof(int));
  memcpy(X, Xp, 3 * nn * sizeof(double));
  memcpy(bc_nodes, bn, nb * sizeof(int));
  memcpy(bc_ptr, bp, (nb + 1) * sizeof(int));
  memcpy(bc_vars, bv, bp[nb] * sizeof(int));
  memcpy(bc_vals, bvals, bp[nb] * sizeof(double));
  for (int k = 0; k < ml->getNumComponents(); k++) {
    memset(&elem_descript[9 * k], 0, 9);
    memset(&comp_descript[33 * k], 0, 33);
    strncpy(&elem_descript[9 * k], ml->getElementDescript(k), 8);
    strncpy(&comp_descript[33 * k], ml->getComponentDescript(k), 32);
  }
}

void refdrv_bdf_free(void *h) { ((TACSMeshLoader *)h)->decref(); }

}  // extern "C"
