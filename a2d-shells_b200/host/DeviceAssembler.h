// DeviceAssembler.h — C++ side-car over the C ABI of include/a2ds.h for drivers that own
// their loop (route 2 of INTEGRATION.md).  Method names, argument order and meaning follow
// TACSAssembler (src/TACSAssembler.h:213-220); the error convention does not — the
// reference prints and carries on, this throws std::runtime_error with a2ds_last_error().
// Header only; link with -la2ds_b200.
#ifndef A2DS_DEVICE_ASSEMBLER_H
#define A2DS_DEVICE_ASSEMBLER_H

#include <stdexcept>
#include <string>
#include <vector>

#include "a2ds.h"

namespace a2ds {

class DeviceAssembler {
 public:
  explicit DeviceAssembler(int device = 0) : ctx_(nullptr), n_nodes_(0), n_owned_(0), part_(nullptr) {
    check(a2ds_create(device, &ctx_), "a2ds_create");
  }
  ~DeviceAssembler() {
    if (part_) a2ds_partition_free(part_);
    if (ctx_) a2ds_destroy(ctx_);
  }
  DeviceAssembler(const DeviceAssembler &) = delete;
  DeviceAssembler &operator=(const DeviceAssembler &) = delete;

  // TACSAssembler::setElementConnectivity / setElements / setNodes / addBCs
  void setMesh(int n_nodes, int n_owned, int n_elems, const int *conn, const int *elem_comp) {
    check(a2ds_set_mesh(ctx_, n_nodes, n_owned, n_elems, conn, elem_comp), "a2ds_set_mesh");
    n_nodes_ = n_nodes; n_owned_ = n_owned;
  }
  // TACSAssembler::setDependentNodes — before setMesh, as the reference wants it before
  // initialize(); the connectivity then refers to dependent node d as -(d + 1)
  void setDependentNodes(int n_dep, const int *dep_ptr, const int *dep_conn, const double *dep_weights) {
    check(a2ds_set_dependent_nodes(ctx_, n_dep, dep_ptr, dep_conn, dep_weights), "a2ds_set_dependent_nodes");
  }
  void setNodes(const double *X) { check(a2ds_set_nodes(ctx_, X), "a2ds_set_nodes"); }
  void setComponents(int n_comp, const double *Cs, const double *eth, const double *temperature,
                     const int *elem_class, int transform, const double *ref_axis) {
    check(a2ds_set_components(ctx_, n_comp, Cs, eth, temperature, elem_class, transform, ref_axis),
          "a2ds_set_components");
  }
  void addBCs(int n_bc, const int *nodes, const int *var_masks, const double *vals) {
    check(a2ds_set_bcs(ctx_, n_bc, nodes, var_masks, vals), "a2ds_set_bcs");
  }
  // TACSAssembler::setVariables (all local nodes, or the owned ones followed by haloForward)
  void setVariables(const double *u, bool owned_only = false) {
    check(a2ds_set_state(ctx_, owned_only ? n_owned_ : n_nodes_, u), "a2ds_set_state");
  }
  // qdot / qddot of TACSAssembler::setVariables (only qddot enters: inertial term)
  void setVariableRates(const double *udot, const double *uddot, bool owned_only = false) {
    check(a2ds_set_state_rates(ctx_, owned_only ? n_owned_ : n_nodes_, udot, uddot),
          "a2ds_set_state_rates");
  }
  void setMassMoments(int n_comp, const double *moments) {
    check(a2ds_set_mass_moments(ctx_, n_comp, moments), "a2ds_set_mass_moments");
  }
  void haloForward() { check(a2ds_halo_forward(ctx_), "a2ds_halo_forward"); }
  // matrices double buffered on the device (default on: no zeroEntries pass per assembly, one more
  // value array per matrix); off frees the second arrays
  void setDoubleBuffer(bool on) { check(a2ds_set_double_buffer(ctx_, on ? 1 : 0), "a2ds_set_double_buffer"); }
  // multi-GPU, one process per GPU: this rank's part of an element-wise partitioned global
  // mesh (node ownership and numbering as TACSCreator::createTACS) and its ghost-exchange
  // plan, then a2ds_set_mesh + a2ds_set_halo.  Returns local -> global node numbers.
  // matrix_halo: TACSParallelMat flavour — createPartitionedMat() then gives matrices whose
  // owned rows are fully assembled (ghost-row blocks travel inside every assemble call)
  std::vector<int> setPartitionedMesh(int n_nodes, int n_elems, const int *conn,
                                      const int *elem_rank, const int *elem_comp, int n_ranks,
                                      int rank, bool matrix_halo = false) {
    a2ds_partition *part = nullptr;
    check(matrix_halo
              ? a2ds_partition_build_matrix(n_nodes, n_elems, conn, elem_rank, n_ranks, rank, &part)
              : a2ds_partition_build(n_nodes, n_elems, conn, elem_rank, n_ranks, rank, &part),
          "a2ds_partition_build");
    const int rc = a2ds_partition_apply(ctx_, part, elem_comp);
    const int *glob = nullptr;
    a2ds_partition_sizes(part, &n_nodes_, &n_owned_, nullptr, nullptr, nullptr, nullptr);
    a2ds_partition_mesh(part, nullptr, nullptr, &glob, nullptr);
    std::vector<int> out(glob, glob + n_nodes_);
    if (part_) a2ds_partition_free(part_);
    part_ = matrix_halo ? part : nullptr;
    if (!matrix_halo) a2ds_partition_free(part);
    check(rc, "a2ds_partition_apply");
    return out;
  }
  int createPartitionedMat() {
    if (!part_) throw std::runtime_error("createPartitionedMat: setPartitionedMesh(..., matrix_halo = true) first");
    int m = -1;
    check(a2ds_partition_create_mat(ctx_, part_, &m), "a2ds_partition_create_mat");
    return m;
  }
  // NCCL communicator: id from a2ds_comm_unique_id on rank 0, broadcast by the launcher
  void initComm(int n_ranks, int rank, const char id[128]) {
    check(a2ds_comm_init(ctx_, n_ranks, rank, id), "a2ds_comm_init");
  }

  // TACSAssembler::createMat (natural order) / matrices whose pattern comes from a host object
  int createMat() {
    int m = -1;
    check(a2ds_mat_create_natural(ctx_, &m), "a2ds_mat_create_natural");
    return m;
  }
  int createMatFromPattern(int n_blocks, const int *nrows, const int *const *rowp,
                           const int *const *cols, const int *const *row_map,
                           const int *const *col_map, const int *bc_ident) {
    int m = -1;
    check(a2ds_mat_create(ctx_, n_blocks, nrows, rowp, cols, row_map, col_map, bc_ident, &m),
          "a2ds_mat_create");
    return m;
  }

  // the three reference entry points; residual: 6 * n_owned doubles or NULL
  void assembleRes(double *residual) { check(a2ds_assemble_res(ctx_, residual), "assembleRes"); }
  void assembleJacobian(double alpha, double beta, double gamma, double *residual, int mat) {
    check(a2ds_assemble_jacobian(ctx_, alpha, beta, gamma, residual, mat), "assembleJacobian");
  }
  void assembleMatType(int mat_type, int mat) {
    check(a2ds_assemble_mat_type(ctx_, mat_type, mat), "assembleMatType");
  }
  // residual + K + G in one pass
  void assembleMatCombo(int n, const int *mat_types, const double *scales, int mat) {
    check(a2ds_assemble_mat_combo(ctx_, n, mat_types, scales, mat), "assembleMatCombo");
  }
  void addJacobianVecProduct(double scale, double alpha, const double *x, double *y) {
    check(a2ds_add_jacobian_vec_product(ctx_, scale, alpha, x, y), "addJacobianVecProduct");
  }
  void assembleAll(double *residual, int kmat, int gmat) {
    check(a2ds_assemble_all(ctx_, residual, kmat, gmat), "assembleAll");
  }

  // TACSMat::copyValues / axpy / applyBCs / mult on the device-resident values
  void copyValues(int dst, int src) { check(a2ds_mat_copy(ctx_, dst, src), "copyValues"); }
  void axpy(double alpha, int x, int y) { check(a2ds_mat_axpy(ctx_, alpha, x, y), "axpy"); }
  void applyBCs(int mat) { check(a2ds_mat_apply_bcs(ctx_, mat), "applyBCs"); }
  void mult(int mat, int block, int ncols, const double *x, double *y) {
    check(a2ds_mat_mult(ctx_, mat, block, ncols, x, y), "mult");
  }
  // BCSRMat::getArrays analogue: values of one BCSR block, A[36 k + 6 r + c]
  std::vector<double> getValues(int mat, int block = 0) {
    long long nnz = 0;
    check(a2ds_mat_nnz(ctx_, mat, block, &nnz), "a2ds_mat_nnz");
    std::vector<double> A(36 * (size_t)nnz);
    check(a2ds_mat_download(ctx_, mat, block, A.data()), "a2ds_mat_download");
    return A;
  }
  a2ds_ctx *context() { return ctx_; }

 private:
  static void check(int rc, const char *what) {
    if (rc) throw std::runtime_error(std::string(what) + ": " + a2ds_last_error());
  }
  a2ds_ctx *ctx_;
  int n_nodes_, n_owned_;
  a2ds_partition *part_;  // kept only in matrix-halo mode (createPartitionedMat needs it)
};

}  // namespace a2ds
#endif
