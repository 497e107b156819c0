/*
 * a2ds.h — C ABI of the B200 shell-assembly library (liba2ds_b200.so).
 *
 * This is the drop-in boundary for ONE path of the reference (a2d-shells, a TACS
 * mini-app): assembly of the residual, tangent stiffness and geometric stiffness
 * of MITC4 director shells into 6x6-block BCSR matrices.  Every entry point
 * below names the reference interface it stands in for (paths relative to the
 * reference root).  Plain pointers and sizes only; all `const double*` / `const
 * int*` arguments are HOST pointers unless the name ends in `_dev`.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error;
 *     a2ds_last_error() returns a description of the last failure of the calling
 *     thread (the reference prints to stderr and carries on,
 *     src/bpmat/BCSRMat.cpp:1803; a C ABI cannot, so it reports);
 *   - one context per GPU, driven by one host thread; work is enqueued on the
 *     context's stream; calls that hand data back to the host synchronise;
 *   - node indices are LOCAL indices of this rank: owned nodes first
 *     [0, n_owned), then ghost nodes (TACSBVec x / x_ext split,
 *     src/bpmat/TACSBVec.h:141-158);
 *   - DOF layout per node [u v w rx ry rz]; element matrix blocks are 6x6,
 *     row-major inside the block, block k at A[36 k]
 *     (src/bpmat/BCSRMatImpl.h:27-46, BCSRMat.cpp:1806-1817).
 *   - there is no CPU fallback: if no CUDA device is usable a2ds_create fails.
 */
#ifndef A2DS_H
#define A2DS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct a2ds_ctx a2ds_ctx;

/* ElementMatrixType, src/elements/TACSElementTypes.h:101-107 */
#define A2DS_STIFFNESS_MATRIX 0
#define A2DS_GEOMETRIC_STIFFNESS_MATRIX 1
#define A2DS_MASS_MATRIX 2

/* element class of a component: TACSQuad4Shell / TACSQuad4NonlinearShell
   (src/elements/shell/TACSShellElementDefs.h:12-14, 27-29) */
#define A2DS_QUAD4_SHELL 0
#define A2DS_QUAD4_NONLINEAR_SHELL 1

/* TACSShellNaturalTransform / TACSShellRefAxisTransform
   (src/elements/shell/TACSShellElementTransform.h:22, 95) */
#define A2DS_TRANSFORM_NATURAL 0
#define A2DS_TRANSFORM_REF_AXIS 1

/* scatter mode: atomics in launch order, or one launch per element colour with a
   fixed colour order (bit-reproducible run to run) */
#define A2DS_SCATTER_ATOMIC 0
#define A2DS_SCATTER_COLORED 1
/* atomics, one launch, elements visited colour by colour: elements in flight together rarely
   share a node (fewer collisions of the atomic adds); not bit-reproducible */
#define A2DS_SCATTER_ATOMIC_COLOR_ORDER 2

const char *a2ds_last_error(void);
const char *a2ds_version(void);

/* ---- context --------------------------------------------------------------- */
int a2ds_create(int device, a2ds_ctx **ctx);
int a2ds_destroy(a2ds_ctx *ctx);
int a2ds_synchronize(a2ds_ctx *ctx);

/* ---- mesh, replaces the element loops' inputs ---------------------------------
 * TACSAssembler::setElementConnectivity / setElements (src/TACSAssembler.h:89-91):
 * conn[4 e + i] is the local node of corner i of element e in the reference's
 * tensor order (src/elements/shell/TACSShellElementQuadBasis.h:147-150);
 * elem_comp[e] selects the component record below. */
int a2ds_set_mesh(a2ds_ctx *ctx, int n_nodes, int n_owned, int n_elems, const int *conn,
                  const int *elem_comp);
/* The same for elements of order x order nodes.  order 2 = a2ds_set_mesh.  order 3: 9-node
 * shells (TACSQuad9Shell = TACSShellElement<TACSQuadQuadraticQuadrature, TACSShellQuadBasis<3>,
 * TACSLinearizedRotation, TACSShellLinearModel>, src/elements/shell/TACSShellElementDefs.h:16-18);
 * conn[9 e + 3 j + i] in the tensor order of TACSShellQuadBasis<3>::getNodePoint
 * (TACSShellElementQuadBasis.h:147-150).  A 9-node mesh supports every assembly entry point of
 * both strain models (TACSQuad9NonlinearShell = elem_class 1): a2ds_assemble_res,
 * a2ds_assemble_jacobian (alpha, gamma, second time derivatives), a2ds_assemble_mat_type
 * (stiffness, geometric stiffness, mass), a2ds_assemble_all, a2ds_assemble_mat_combo — with the
 * atomic scatter, on natural-order or caller-supplied patterns — plus everything that works on
 * nodes and blocks (boundary conditions, halo exchange, matrix algebra, mat-vec); the coloured
 * scatter modes and the matrix-free product fail for it. */
int a2ds_set_mesh_order(a2ds_ctx *ctx, int order, int n_nodes, int n_owned, int n_elems,
                        const int *conn, const int *elem_comp);

/* TACSAssembler::setDependentNodes (src/TACSAssembler.cpp:716-775).  Dependent node d is the
 * weighted sum  sum_j dep_weights[j] * node dep_conn[j],  j in dep_ptr[d] .. dep_ptr[d + 1]
 * (local independent nodes), and appears in the connectivity as the entry -(d + 1), exactly as
 * in the reference's elementTacsNodes.  Like the reference (before initialize()) the call comes
 * BEFORE a2ds_set_mesh / a2ds_set_mesh_order, which apply it; n_dep = 0 withdraws it.  Replaces
 * the dependent-node branches of TACSBVec::endDistributeValues / beginSetValues
 * (src/bpmat/TACSBVec.cpp:930-975, :855-885: values of a dependent node gathered from, residual
 * rows distributed to its independent nodes), TACSAssembler::addMatValues ->
 * addWeightValues (src/TACSAssembler.h:469-510: W^T K_e W) and computeLocalNodeToNodeCSR
 * (src/TACSAssembler.cpp:1850-1935: the pattern couples all independent nodes behind an
 * element).  On the device a dependent node is one more row behind the local rows of X / u /
 * res, the element kernels are unchanged; node pairs with a dependent node go to scratch blocks
 * behind the matrix and are distributed by a fold kernel whose target list is built with the
 * matrix.  Vectors at the boundary keep n_nodes rows.  Atomic scatter only; the matrix-free
 * product and the streamed assembly are not available on such meshes. */
int a2ds_set_dependent_nodes(a2ds_ctx *ctx, int n_dep, const int *dep_ptr, const int *dep_conn,
                             const double *dep_weights);

/* TACSAssembler::setNodes (src/TACSAssembler.cpp:912): X[3 n + k], all local nodes */
int a2ds_set_nodes(a2ds_ctx *ctx, const double *X);

/* Per-component tables.  Replaces the virtual calls made per quadrature point:
 *   con->evalTangentStiffness -> Cs[22 c + .]   (src/constitutive/TACSShellConstitutive.h:34)
 *   con->evalThermalStrain(theta = 1) -> eth[9 c + .]  (TACSIsoShellConstitutive.cpp:438-456)
 *   element->temperature -> temperature[c]      (src/elements/shell/TACSShellElement.h:35)
 *   element class -> elem_class[c];  transform kind + reference axis (common to all) */
int a2ds_set_components(a2ds_ctx *ctx, int n_comp, const double *Cs, const double *eth,
                        const double *temperature, const int *elem_class, int transform,
                        const double *ref_axis);

/* Mass moments of every component, moments[3 c + k], as the host constitutive object
 * returns them (TACSShellConstitutive::evalMassMoments,
 * src/constitutive/TACSIsoShellConstitutive.cpp:120-129).  Only the mass path needs them;
 * may be called before or after a2ds_set_components (default: zero). */
int a2ds_set_mass_moments(a2ds_ctx *ctx, int n_comp, const double *moments);

/* TACSAssembler::setVariables (src/TACSAssembler.cpp:3825-3857): u[6 n + k] for all
 * local nodes when n_given == n_nodes, or for the owned nodes only when
 * n_given == n_owned (ghost values then come from a2ds_halo_forward).  The upload runs on a
 * copy stream and overlaps the zeroing of the next assembly's outputs; the next consumer waits
 * for it.  With a page-locked `u` the call returns before the copy has read it: leave the
 * buffer unchanged until a call that synchronises (an assemble call with a host result,
 * a2ds_synchronize).  Pageable memory is read before the call returns. */
int a2ds_set_state(a2ds_ctx *ctx, int n_given, const double *u);
int a2ds_set_state_dev(a2ds_ctx *ctx, int n_given, const double *u_dev);
/* the qdot / qddot arguments of TACSAssembler::setVariables.  Only uddot enters this element
 * class (inertial term M * uddot of the residual, TACSShellElement.h:410-447); udot is
 * accepted for signature parity.  uddot == NULL removes the inertial term again.  With a
 * halo and n_given == n_owned the ghost values are exchanged here. */
int a2ds_set_state_rates(a2ds_ctx *ctx, int n_given, const double *udot, const double *uddot);

/* TACSBcMap (src/bpmat/KSM.h:43-75): bc_nodes[b] local node, bc_vars[b] bit mask of
 * constrained DOFs (bits above the six shell DOFs are ignored), bc_vals[6 b + k] prescribed
 * values (NULL: all zero).  Nodes outside [0, n_nodes) are refused. */
int a2ds_set_bcs(a2ds_ctx *ctx, int n_bc, const int *bc_nodes, const int *bc_vars,
                 const double *bc_vals);

/* A2DS_SCATTER_* ; colouring is computed on the host when first needed */
int a2ds_set_scatter_mode(a2ds_ctx *ctx, int mode);

/* ---- matrices ---------------------------------------------------------------
 * A matrix is 1..4 BCSR blocks that share one device allocation, mirroring
 * TACSParallelMat {Aloc, Bext} (src/bpmat/TACSParallelMat.h:103) and TACSSchurMat
 * {B, E, F, C} (src/bpmat/TACSSchurMat.h:98).  The non-zero pattern is TAKEN from the
 * host matrices (BCSRMat::getArrays, src/bpmat/BCSRMat.h:82), never recomputed.
 * For block b: row_map[b][n] / col_map[b][n] give the block-local row / column of
 * local node n, or -1 when the node is not in that block's row / column set (NULL =
 * identity); this replaces the two-level search of TACSSchurMat::addValues
 * (src/bpmat/TACSSchurMat.cpp:453-531) + BCSRMat::addRowValues (BCSRMat.cpp:1778).
 * bc_ident[b] != 0: the block holds the diagonal — BC rows get 1.0 there
 * (TACSSchurMat::applyBCs, TACSSchurMat.cpp:662-703).
 * Returns a matrix id >= 0 in *mat. */
int a2ds_mat_create(a2ds_ctx *ctx, int n_blocks, const int *nrows, const int *const *rowp,
                    const int *const *cols, const int *const *row_map,
                    const int *const *col_map, const int *bc_ident, int *mat);
/* Same, with the pattern computed here from the element connectivity: one block,
 * one row per local node, columns = all nodes sharing an element with the row
 * node, sorted ascending.  This is the pattern TACSAssembler::createMat builds
 * (computeLocalNodeToNodeCSR, src/TACSAssembler.cpp:1839 + TacsSortAndUniquifyCSR,
 * src/utils/TacsUtilities.cpp:280) for natural ordering; tests check it is
 * identical to the reference's, bit for bit. */
int a2ds_mat_create_natural(a2ds_ctx *ctx, int *mat);
/* copy the pattern of a block to the host: rowp[nrows + 1], cols[nnz] (either may be NULL) */
int a2ds_mat_pattern(a2ds_ctx *ctx, int mat, int block, int *nrows, int *rowp, int *cols);
/* number of 6x6 blocks stored in BCSR block `block` */
int a2ds_mat_nnz(a2ds_ctx *ctx, int mat, int block, long long *nnz);
/* TACSMat::zeroEntries (src/bpmat/BCSRMat.cpp:1745) */
int a2ds_mat_zero(a2ds_ctx *ctx, int mat);
/* copy block values to the host array A[36 nnz] (the "hand K and G back to the
 * reference solver" path: BCSRMat::getArrays) / borrow the device pointer */
int a2ds_mat_download(a2ds_ctx *ctx, int mat, int block, double *A);
int a2ds_mat_values_dev(a2ds_ctx *ctx, int mat, int block, double **A_dev);
/* NOTE: the borrowed pointer is valid until the next a2ds_assemble_* call that zeroes this matrix:
 * matrices that take part in assembleJacobian / assembleMatType(G) / a2ds_assemble_all are double
 * buffered on the device (the element kernel zeroes the value array the NEXT assembly will add
 * into, and that assembly swaps the two arrays instead of zeroing: TACSMat::zeroEntries without a
 * pass over the matrix); ask again after each assembly.  Every other entry point (download, copy,
 * axpy, mult, apply_bcs) follows the swap.  A2DS_DOUBLE_BUFFER=0 switches it off; it is also
 * skipped for a matrix whose second array would take the last 15 % of the GPU's memory. */
/* the same switch at run time; on = 0 also frees the second value arrays (one matrix size each) */
int a2ds_set_double_buffer(a2ds_ctx *ctx, int on);
/* the blocks of the listed block rows only, row after row (36 doubles per block): spot checks
 * of matrices too large to copy back (a row loop over BCSRMat::getArrays, BCSRMat.cpp:2312) */
int a2ds_mat_download_rows(a2ds_ctx *ctx, int mat, int block, int n_rows, const int *rows,
                           double *A);

/* ---- matrix algebra on the device-resident values (the buckling flow's K/G handling) ----
 * TACSMat::copyValues (src/bpmat/BCSRMat.cpp:2375): dst <- src, same pattern required */
int a2ds_mat_copy(a2ds_ctx *ctx, int dst, int src);
/* TACSMat::axpy (src/bpmat/BCSRMat.cpp:2430): y <- y + alpha x, same pattern required
 * (aux = K + sigma G, src/TACSBuckling.cpp:269) */
int a2ds_mat_axpy(a2ds_ctx *ctx, double alpha, int x, int y);
/* TACSMat::applyBCs (src/bpmat/TACSSchurMat.cpp:662-703, BCSRMat::zeroRow) */
int a2ds_mat_apply_bcs(a2ds_ctx *ctx, int mat);
/* 6x6 BCSR mat-vec y = A x for one BCSR block (BCSRMatVecMult6,
 * src/bpmat/BCSRMatMult6.cpp:82): x has 6*ncols, y 6*nrows entries in the block's own
 * row/column numbering.  _dev: device pointers; the other takes host pointers. */
int a2ds_mat_mult_dev(a2ds_ctx *ctx, int mat, int block, const double *x_dev, double *y_dev);
int a2ds_mat_mult(a2ds_ctx *ctx, int mat, int block, int ncols, const double *x, double *y);
/* TACSParallelMat::mult (src/bpmat/TACSParallelMat.cpp:248) for the owned rows of a matrix
 * that every rank assembled over its local nodes (a2ds_mat_create_natural; interface rows
 * stay unassembled per rank): ghost entries of x are fetched from their owners, y = A x over
 * all local rows, ghost rows of y are added at their owners (the halo of a2ds_set_halo),
 * y = x on constrained DOFs.  x_dev, y_dev: 6 * n_nodes doubles on the device; the ghost
 * part of x is overwritten.  Without a halo this is a2ds_mat_mult_dev plus the BC rows. */
int a2ds_mat_mult_dist_dev(a2ds_ctx *ctx, int mat, double *x_dev, double *y_dev);

/* ---- assembly: the three reference entry points -------------------------------
 * TACSAssembler::assembleRes (src/TACSAssembler.cpp:4000-4063): res[6 n + k] for the
 * owned nodes, boundary rows r = u - ubar.  res may be NULL (result stays on device,
 * see a2ds_res_dev). */
int a2ds_assemble_res(a2ds_ctx *ctx, double *res);
/* TACSAssembler::assembleJacobian (src/TACSAssembler.cpp:4084-4174): zero res and A, add
 * the residual and alpha * dR/du + gamma * dR/d(uddot) (the mass matrix,
 * src/elements/shell/TACSShellElement.h:614-648), apply BCs to both.  beta is accepted and
 * has no effect: this element class has no velocity dependent term.  res may be NULL. */
int a2ds_assemble_jacobian(a2ds_ctx *ctx, double alpha, double beta, double gamma,
                           double *res, int mat);
/* TACSAssembler::assembleMatType (src/TACSAssembler.cpp:4186-4249), A2DS_*_MATRIX */
int a2ds_assemble_mat_type(a2ds_ctx *ctx, int mat_type, int mat);
/* TACSAssembler::assembleMatCombo (src/TACSAssembler.cpp:4264-4318):
 * A = sum_i scales[i] * matType(mat_types[i]), BCs applied once at the end */
int a2ds_assemble_mat_combo(a2ds_ctx *ctx, int n, const int *mat_types, const double *scales,
                            int mat);
/* residual + tangent + geometric stiffness in one pass over the elements (what the
 * buckling flow asks for in three calls, src/TACSBuckling.cpp:239-266) */
int a2ds_assemble_all(a2ds_ctx *ctx, double *res, int kmat, int gmat);

/* TACSAssembler::addJacobianVecProduct (src/TACSAssembler.cpp:4331-4391), matrix free:
 * y <- y + scale * (alpha K) x, BC rows of y zeroed; x, y hold 6 * n_nodes doubles (all
 * local nodes; ghost entries of x must be current).  Linear-strain elements: K x is the
 * residual kernel evaluated at x; nonlinear-strain elements: the element tangent about
 * the current state is formed on chip and multiplied with x instead of being scattered. */
int a2ds_add_jacobian_vec_product(a2ds_ctx *ctx, double scale, double alpha, const double *x,
                                  double *y);
int a2ds_add_jacobian_vec_product_dev(a2ds_ctx *ctx, double scale, double alpha,
                                      const double *x_dev, double *y_dev);

/* device-resident residual of the last assembly, 6 * n_nodes doubles */
int a2ds_res_dev(a2ds_ctx *ctx, double **res_dev);
/* device-resident state, 6 * n_nodes doubles */
int a2ds_state_dev(a2ds_ctx *ctx, double **u_dev);

/* ---- multi-GPU: ghost exchange ------------------------------------------------
 * TACSBVecDistribute (src/bpmat/TACSBVecDistribute.cpp:543-747).  Peers are the other
 * ranks of an NCCL communicator; the 128-byte unique id is created on rank 0 with
 * a2ds_comm_unique_id and broadcast by the caller (torch.distributed / MPI).
 * send_ptr/send_nodes: for peer p, the owned local nodes whose values that peer
 * reads as ghosts; recv_ptr/recv_nodes: the local ghost nodes owned by peer p. */
int a2ds_comm_unique_id(char id[128]);
int a2ds_comm_init(a2ds_ctx *ctx, int n_ranks, int rank, const char id[128]);
int a2ds_set_halo(a2ds_ctx *ctx, int n_peers, const int *peer_rank, const int *send_ptr,
                  const int *send_nodes, const int *recv_ptr, const int *recv_nodes);
/* The same lists from the reference's own plan: TACSBVecDistribute's slab lists of GLOBAL node
 * numbers (src/bpmat/TACSBVecDistribute.h:156-172: ext_proc/ext_ptr/ext_count over the sorted
 * ext_vars this rank reads, req_proc/req_ptr/req_count/req_vars the owned nodes other ranks
 * read) in the device path's local numbering (owned g - lo first, then the ghosts in ext_vars
 * order).  Outputs sized by the caller: peer_rank[n_ext_proc + n_req_proc], send_ptr / recv_ptr
 * one longer, send_nodes[sum req_count], recv_nodes[sum ext_count].  Host only. */
int a2ds_halo_from_distribute(int lo, int n_owned, int n_ext_proc, const int *ext_proc,
                              const int *ext_ptr, const int *ext_count, int n_req_proc,
                              const int *req_proc, const int *req_ptr, const int *req_count,
                              const int *req_vars, int *n_peers, int *peer_rank, int *send_ptr,
                              int *send_nodes, int *recv_ptr, int *recv_nodes);
/* forward: owner -> ghost copies of the state (beginForward/endForward) */
int a2ds_halo_forward(a2ds_ctx *ctx);

/* Matrix halo for the TACSParallelMat flavour (TACSMatDistribute::beginAssembly/endAssembly,
 * src/bpmat/TACSMatDistribute.cpp:1036-1176): after the element loop the blocks a rank
 * contributed to block rows it does not own are sent to the owners and added there, so
 * that every rank's OWNED rows are fully assembled.  The plan comes from the host, which
 * knows the global numbering: per peer, send_blocks = indices (into the concatenated block
 * array of this matrix) of my ghost-row blocks, recv_blocks = indices of the blocks the
 * peer's contributions are added to, both in one order agreed by the two sides (e.g.
 * sorted by (global row, global column)).  The owner's pattern must contain every
 * arriving block, i.e. its local node set includes the column nodes of those contributions
 * (the reference's external column map does the same).  Once set, every assemble call on
 * this matrix performs the exchange before the boundary conditions.  Matrices without a
 * plan keep the TACSSchurMat convention (interface rows unassembled per rank). */
int a2ds_mat_set_halo(a2ds_ctx *ctx, int mat, int n_peers, const int *peer_rank,
                      const int *send_ptr, const int *send_blocks, const int *recv_ptr,
                      const int *recv_blocks);
/* reverse: ghost residual contributions added to their owners, done inside
 * a2ds_assemble_* when a halo is set (beginReverse/endReverse, TACS_ADD_VALUES) */

/* ---- host-only helpers (no device needed) ---------------------------------------
 * The natural-order non-zero pattern used by a2ds_mat_create_natural: two-pass, call
 * with cols == NULL to fill rowp[n_nodes + 1] and obtain *nnz, then again with cols. */
int a2ds_host_pattern(int n_nodes, int n_elems, const int *conn, int *rowp, int *cols,
                      long long *nnz);
/* Greedy element colouring used by A2DS_SCATTER_COLORED: color[e] in [0, *n_colors);
 * elements sharing a node never share a colour. */
int a2ds_host_color_elements(int n_nodes, int n_elems, const int *conn, int *color,
                             int *n_colors);
/* The colouring rule the device uses (k_color_round, csrc/aux_kernels.cuh): greedy in the order of
 * a hashed element priority — a Jones-Plassmann sweep, O(log n) rounds on the device — stepped
 * sequentially on the host; bit-identical to a2ds_get_element_colors. */
int a2ds_host_color_elements_hashed(int n_nodes, int n_elems, const int *conn, int *color,
                                    int *n_colors);
/* Element colours of the context's mesh as A2DS_SCATTER_COLORED / _COLOR_ORDER use them, computed
 * on the device (4-node meshes; color may be NULL to obtain only *n_colors).  The reference has
 * no counterpart (its threaded element loop adds under a mutex, src/TACSAssembler.cpp:4561-4576). */
int a2ds_get_element_colors(a2ds_ctx *ctx, int *color, int *n_colors);

/* ---- partition (host only) ---------------------------------------------------------
 * One rank's part of an element-wise partitioned global mesh and its ghost-exchange plan,
 * derived without communication from the global connectivity (4 global nodes per element)
 * and the element -> rank array every rank holds.  Node ownership and numbering follow
 * TACSCreator::createTACS (src/TACSCreator.cpp:1156-1205): a node belongs to the rank of the
 * first element, in global element order, that refers to it; local numbering is the owned
 * nodes (ascending global number) followed by the ghosts (ascending).  The halo lists are
 * what TACSBVecDistribute is built from in TACSAssembler::initialize
 * (src/bpmat/TACSBVecDistribute.cpp:280-420): per peer, my nodes it reads as ghosts and my
 * ghosts it owns, both in ascending global order — so the two sides of a pair agree. */
typedef struct a2ds_partition a2ds_partition;
/* element -> rank by recursive coordinate bisection of the element centroids (X: 3 per global
 * node): balanced, deterministic, compact parts.  In place of the METIS call of
 * TACSCreator::partitionMesh (src/TACSCreator.cpp:923, METIS calls :1118-1125); any other element -> rank array
 * (slabs, METIS output) works with a2ds_partition_build just as well. */
int a2ds_partition_rcb(int n_nodes, int n_elems, const int *conn, const double *X, int n_ranks,
                       int *elem_rank);
int a2ds_partition_build(int n_nodes, int n_elems, const int *conn, const int *elem_rank,
                         int n_ranks, int rank, a2ds_partition **part);
void a2ds_partition_free(a2ds_partition *part);
int a2ds_partition_sizes(const a2ds_partition *part, int *n_local_nodes, int *n_owned,
                         int *n_local_elems, int *n_peers, int *n_send, int *n_recv);
/* borrowed pointers: elems[n_local_elems] global element ids; conn_local[4 n_local_elems];
 * glob[n_local_nodes] local -> global node; ghost_owner[n_local_nodes - n_owned] */
int a2ds_partition_mesh(const a2ds_partition *part, const int **elems, const int **conn_local,
                        const int **glob, const int **ghost_owner);
/* the arguments of a2ds_set_halo */
int a2ds_partition_halo(const a2ds_partition *part, const int **peers, const int **send_ptr,
                        const int **send_nodes, const int **recv_ptr, const int **recv_nodes);
/* a2ds_set_mesh + a2ds_set_halo for this rank; elem_comp: component per GLOBAL element or NULL */
int a2ds_partition_apply(a2ds_ctx *ctx, const a2ds_partition *part, const int *elem_comp);
/* The same partition prepared for the TACSParallelMat flavour (every rank's OWNED matrix rows
 * fully assembled, a2ds_mat_set_halo): the local node set additionally holds every node that
 * shares an element of ANY rank with an owned node (the reference's external column map,
 * TACSMatDistribute, src/bpmat/TACSMatDistribute.cpp:150-330), the vector halo covers them,
 * and a2ds_partition_matrix hands out the local pattern (full rows for the owned nodes, the
 * local elements' couplings for ghost rows; columns ascending) and, per peer in the order of
 * a2ds_partition_halo, the block indices to send / to add arriving blocks to, both sorted by
 * (global row, global column) so that the two sides of a pair agree. */
int a2ds_partition_build_matrix(int n_nodes, int n_elems, const int *conn, const int *elem_rank,
                                int n_ranks, int rank, a2ds_partition **part);
int a2ds_partition_matrix(const a2ds_partition *part, const int **rowp, const int **cols,
                          const int **send_ptr, const int **send_blocks, const int **recv_ptr,
                          const int **recv_blocks);
/* a2ds_mat_create with that pattern + a2ds_mat_set_halo */
int a2ds_partition_create_mat(a2ds_ctx *ctx, const a2ds_partition *part, int *mat);

/* ---- mesh input (host only) -------------------------------------------------------
 * The data format in front of the path: NASTRAN bulk-data decks as the reference's examples
 * ship them, and a flat binary container for meshes too large to parse at every start.
 *
 * a2ds_mesh_read_bdf stands in for TACSMeshLoader::scanBDFFile
 * (src/io/TACSMeshLoader.cpp:570-1096): GRID / GRID* / SPC / SPC* and the element keywords of
 * src/io/TACSMeshLoader.h:20-35 in small, large and comma-separated fields; nodes and elements
 * come out sorted by their file numbers (0-based), element nodes in the tensor-product order of
 * the element basis, components 0-based — array for array what getConnectivity / getBCs
 * (:1221-1278) hand out.  n_threads <= 0: one parser per host core (at most 16); the result
 * does not depend on it.  Failures return non-zero (missing file, empty line inside the bulk
 * data — where the reference stops too —, an element card without its numbers or with too few
 * nodes, references to undefined grid points); unknown cards are reported on stderr and
 * skipped, as the reference does. */
typedef struct a2ds_mesh a2ds_mesh;
int a2ds_mesh_read_bdf(const char *path, int n_threads, a2ds_mesh **mesh);
/* binary container written by a2ds_mesh_write_bin: no parsing, arrays read as they are */
int a2ds_mesh_read_bin(const char *path, a2ds_mesh **mesh);
int a2ds_mesh_write_bin(const a2ds_mesh *mesh, const char *path);
/* a container from arrays in memory (generated meshes); BC arrays may be NULL with n_bcs == 0 */
int a2ds_mesh_from_arrays(int n_nodes, int n_elems, const int *elem_ptr, const int *elem_conn,
                          const int *elem_comp, const double *X, int n_bcs, const int *bc_nodes,
                          const int *bc_ptr, const int *bc_vars, const double *bc_vals,
                          a2ds_mesh **mesh);
void a2ds_mesh_free(a2ds_mesh *mesh);
/* getNumNodes / getNumElements / getNumComponents (:1102, :1189, :537); conn_size and bc_size
 * are the lengths of elem_conn and bc_vars / bc_vals.  Any pointer may be NULL. */
int a2ds_mesh_sizes(const a2ds_mesh *mesh, int *n_nodes, int *n_elems, int *conn_size,
                    int *n_bcs, int *bc_size, int *n_comp);
/* TACSMeshLoader::getConnectivity (:1221-1247): borrowed pointers, valid until a2ds_mesh_free;
 * elem_ptr[n_elems + 1], elem_conn[conn_size], elem_comp[n_elems], X[3 n_nodes] */
int a2ds_mesh_connectivity(const a2ds_mesh *mesh, const int **elem_ptr, const int **elem_conn,
                           const int **elem_comp, const double **X);
/* TACSMeshLoader::getBCs (:1252-1278): one entry per SPC card; bc_vars are 0-based DOFs
 * (digits 1..8 of the card), bc_vals the card's value repeated per DOF */
int a2ds_mesh_bcs(const a2ds_mesh *mesh, const int **bc_nodes, const int **bc_ptr,
                  const int **bc_vars, const double **bc_vals);
/* the file's own node / element numbers (0-based, ascending): entry k belongs to node /
 * element k of the arrays above (what getAssemblerNodeNums searches, :1200-1215) */
int a2ds_mesh_file_numbers(const a2ds_mesh *mesh, const int **node_nums, const int **elem_nums);
/* getElementDescript / getComponentDescript (:552-566): keyword of the first element of the
 * component, and its name from an ICEM "$       Shell" comment ("" if there is none) */
int a2ds_mesh_component(const a2ds_mesh *mesh, int comp, const char **elem_descript,
                        const char **comp_descript);
/* the arrays a2ds_set_mesh / a2ds_set_bcs take, for a deck of 4-node shells: conn4[4 n_elems],
 * per SPC card a DOF bit mask and six prescribed values.  Fails if an element is not 4-noded. */
int a2ds_mesh_quad4(const a2ds_mesh *mesh, int *conn4, int *bc_masks, double *bc_vals6);

/* ---- instrumentation ------------------------------------------------------------
 * device time of the last assemble call in milliseconds (CUDA events on the context
 * stream) and the number of kernels it launched */
int a2ds_last_timing(a2ds_ctx *ctx, float *ms, int *launches);
/* device time of the element kernel(s) alone in the last assemble call (the memsets,
 * halo and BC kernels excluded) */
int a2ds_last_kernel_ms(a2ds_ctx *ctx, float *ms);
/* bracket an arbitrary region of calls with CUDA events on the context stream;
 * a2ds_region_end waits for the region to finish and returns its device time */
int a2ds_region_begin(a2ds_ctx *ctx);
int a2ds_region_end(a2ds_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif
