// a2ds.cu — host side and C ABI of liba2ds_b200.so (include/a2ds.h): contexts, meshes,
// components, matrices (patterns taken from host matrices or built in natural order, element ->
// block offset tables), element lists and colouring, the assemble entry points with their
// zeroing / halo / boundary-condition sequence, matrix algebra, NCCL vector and matrix halos.
// The kernels live in assemble_kernels.cuh (k_assemble, k_mass) and aux_kernels.cuh (offsets,
// BCs, halo pack / unpack, copy / axpy, 6x6 BCSR mat-vec); the per-lane element math in
// mitc4_math.h; mesh input and the partition planner (host only) in mesh_io.cpp, partition.cpp.
#include <cuda_runtime.h>
#include <memory>
#include <cub/cub.cuh>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/a2ds.h"
#include "mitc4_math.h"

using namespace a2ds;

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const std::string &m) {
  g_err = m;
  return 1;
}
// A2DS_VERBOSE=1: wall time of the host-side set-up stages on stderr
struct StageTimer {
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  const bool on = getenv("A2DS_VERBOSE") != nullptr;
  void lap(const char *what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[a2ds] %-28s %.3f s\n", what, std::chrono::duration<double>(n - t).count());
    t = n;
  }
};

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                 \
  } while (0)
#define NC(call)                                                                       \
  do {                                                                                 \
    ncclResult_t r_ = (call);                                                          \
    if (r_ != ncclSuccess) return fail(std::string(#call) + ": " + ncclGetErrorString(r_)); \
  } while (0)

// used by the other translation units of the library (mesh_io.cpp)
int a2ds_set_error_(const char *msg) { return fail(msg); }

// no C++ exception leaves the C ABI (std::bad_alloc of a host-side table and the like):
// every multi-line entry point runs between these two
#define A2DS_TRY try {
#define A2DS_CATCH(fn)                                              \
  }                                                                 \
  catch (const std::exception &e_) {                                \
    return fail(std::string(#fn ": ") + e_.what());                 \
  }

extern "C" const char *a2ds_last_error(void) { return g_err.c_str(); }
extern "C" const char *a2ds_version(void) { return "a2ds-b200 0.1 (sm_100a)"; }

#include "assemble_kernels.cuh"
#include "assemble9_kernels.cuh"
#include "aux_kernels.cuh"

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct MatrixRec {
  bool dead = false;   // its mesh was replaced: device memory freed, id kept invalid
  int n_blocks = 0;
  std::vector<long long> nnz, base;
  std::vector<int> nrows;
  long long total = 0;
  double *A = nullptr;
  // double buffering (A2DS_ZWAIT == 2): a second value array; the element kernel that adds into
  // A zeroes it on the side for the next assembly, which swaps the two instead of zeroing A
  double *spare = nullptr;
  bool spare_clean = false, spare_refused = false;
  int *off = nullptr;          // 16 per element
  BlockDev *blk_dev = nullptr;
  std::vector<void *> owned;   // device allocations to free
  // host copy of the patterns (shared between the natural-order matrices of one mesh)
  std::vector<std::shared_ptr<const std::vector<int>>> h_rowp, h_cols;
  std::vector<const int *> d_rowp, d_cols;       // device copy, per block
  unsigned long long pattern_hash = 0;           // fingerprint of (rowp, cols), 0 = not computed
  unsigned long long *shared_hash = nullptr;     // natural-order matrices of one mesh share it
  // dependent nodes: n_fold scratch blocks behind the `total` blocks of the matrix take the
  // element blocks of node pairs with a dependent node; k_dep_fold distributes them (W^T K_e W)
  int n_fold = 0;
  int *fold_ptr = nullptr, *fold_tgt = nullptr;
  double *fold_w = nullptr;
  // matrix halo (ParallelMat flavour): blocks of ghost rows sent to / received from peers
  bool has_halo = false;
  std::vector<int> halo_peers, halo_send_ptr, halo_recv_ptr;
  int *halo_send_blk = nullptr, *halo_recv_blk = nullptr;
  double *halo_send_buf = nullptr, *halo_recv_buf = nullptr;
};

// element ranges and residual row chunks of the streamed assembly (host, cached per mesh)
struct StreamPlan {
  bool ready = false;
  int C = 0;
  std::vector<int> e0;        // C + 1 element boundaries
  std::vector<int> max_node;  // highest node an element range refers to
  std::vector<int> order;     // launch order of the ranges
  std::vector<int> row_lo;    // R + 1 boundaries of the owned node rows (R = ROWS_PER_CHUNK * C)
  std::vector<int> row_pos;   // launch position after which a row chunk is final (C: after the reverse halo)
};

struct a2ds_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // host -> device state uploads run on their own stream so that they overlap the zeroing of
  // the matrices at the start of the next assembly (both take ~1 ms at 1 M elements)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_state = nullptr, ev_used = nullptr;
  bool state_pending = false;
  bool halo_pending = false;  // forward ghost exchange deferred until the upload is consumed
  // streamed assembly (large meshes with host I/O in the step): the state goes up in row chunks,
  // the element kernel runs per element range as soon as its rows have arrived, and the
  // residual rows go back (on d2h_stream) as soon as the last range that adds to them is done
  static const int MAX_CHUNKS = 16, ROWS_PER_CHUNK = 4;   // residual row chunks are finer than the ranges
  cudaStream_t d2h_stream = nullptr;
  // ... and the matrices are zeroed row range by row range on zero_stream, ahead of the element
  // range that needs them, next to the kernel of the range before (natural-order matrices)
  cudaStream_t zero_stream = nullptr;
  cudaEvent_t ev_zero_go = nullptr, ev_z[MAX_CHUNKS + 1] = {};
  bool stream_zero = true;       // A2DS_STREAM_ZERO=0: zero the matrices in front of the first range
  bool stream_resident = false;  // A2DS_STREAM_RESIDENT=1: steps without host I/O take the element ranges too (for the zeroing)
  cudaEvent_t ev_up[MAX_CHUNKS] = {}, ev_row[MAX_CHUNKS * ROWS_PER_CHUNK] = {};
  int stream_chunks = 8;         // A2DS_STREAM_CHUNKS (1: off)
  int stream_min_elems = 200000; // A2DS_STREAM_MIN_ELEMS: smaller meshes are not worth the launch tails
  int up_chunks = 0, up_rows = 0;   // chunks / node rows of the upload in flight
  int up_hi[MAX_CHUNKS] = {};       // node rows [0, up_hi[k]) are on the device once ev_up[k] has fired
  struct StreamPlan splan[2];       // [ghost-touching ranges last]
  std::vector<int> h_send_nodes;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr, evr0 = nullptr, evr1 = nullptr;
  int n_sm = 0;
  int n_nodes = 0, n_owned = 0, n_elems = 0, n_comp = 0, n_bc = 0;
  int npe = 4;   // nodes per element: 4 (MITC4, every entry point) or 9 (MITC9, see a2ds_set_mesh_order)
  bool mesh_set = false;
  // dependent nodes (a2ds_set_dependent_nodes; TACSAssembler::setDependentNodes): the spec is
  // kept on the host and applied by the next a2ds_set_mesh; dependent node d of the current mesh
  // is row n_nodes + d of X / u / res / udd, and h_conn refers to it by that row
  std::vector<int> h_dep_ptr, h_dep_conn;
  std::vector<double> h_dep_w;
  int n_dep = 0;
  int *dep_ptr = nullptr, *dep_conn = nullptr;
  double *dep_w = nullptr;
  std::vector<int> h_dep_slots;   // e * npe^2 + slot of every element node pair with a dependent node
  size_t n_ext() const { return (size_t)n_nodes + (size_t)n_dep; }
  int *conn = nullptr, *elem_comp = nullptr;
  std::vector<int> h_conn, h_elem_comp, h_class;
  // natural-order pattern of the current mesh, built by the first a2ds_mat_create_natural
  std::shared_ptr<std::vector<int>> nat_rowp, nat_cols;   // host copy
  int *nat_d_rowp = nullptr, *nat_d_cols = nullptr;       // device copy, shared by those matrices
  unsigned long long nat_hash = 0;
  bool nat_ready = false;
  std::vector<double> h_mom;          // mass moments per component (a2ds_set_mass_moments)
  std::vector<CompData> h_comps;
  double *X = nullptr, *u = nullptr, *res = nullptr;
  // persistent device scratch of the host-pointer variants (mat-vec, matrix-free product):
  // two vectors, grown on demand, so that a Krylov loop does not allocate per call
  double *scratch_x = nullptr, *scratch_y = nullptr;
  size_t scratch_nx = 0, scratch_ny = 0;
  void *shape9_dev = nullptr;    // shape function tables of the 9-node element (k_mass9)
  int *work_counter = nullptr;   // [0] batch counter, [1..] zero_done rounds; zeroed before every k_assemble launch
  // matrices the next k_assemble_t launch has to zero itself (in-kernel zeroing)
  double *pz_K = nullptr, *pz_G = nullptr;
  // spare arrays of the request being assembled (whole arrays; the streamed assembly hands
  // them to its element ranges piece by piece through pz_*)
  double *sp_K = nullptr, *sp_G = nullptr;
  long long sp_nK = 0, sp_nG = 0, sp_doneK = 0, sp_doneG = 0;
  bool double_buffer = true;     // A2DS_DOUBLE_BUFFER=0: zero the matrices in front of the kernel
  struct ZeroPlan *zplan_dev = nullptr;
  long long pz_nK = 0, pz_nG = 0;
  double *udd = nullptr;  // second time derivative of the state (null until set)
  CompData *comps = nullptr;
  int *bc_nodes = nullptr, *bc_vars = nullptr;
  double *bc_vals = nullptr;
  std::vector<MatrixRec> mats;
  int scatter_mode = A2DS_SCATTER_ATOMIC;
  int warps_per_block_forced = 0;  // A2DS_WARPS_PER_BLOCK: development override
  // element lists: [class][colour]; colour list 0 of the atomic mode holds everything
  bool lists_ready = false;
  int n_colors = 0;
  // class = strain model (0 linear, 1 nonlinear) + 2 * (section with membrane-bending coupling)
  std::vector<int *> list_dev[4];
  std::vector<int> list_len[4];
  // halo
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;
  std::vector<int> peers, send_ptr, recv_ptr;
  int *send_nodes = nullptr, *recv_nodes = nullptr;
  double *send_buf = nullptr, *recv_buf = nullptr;
  bool has_halo = false;
  // instrumentation
  float last_ms = 0.f;
  int last_launches = 0;
};

static int halo_exchange(a2ds_ctx *c, double *vec, bool reverse);

// boundary k of C pieces of the streamed assembly (upload chunks, element ranges): a smoothstep,
// 3.5 %, 12.6 %, 26 % ... 96.5 % for C = 9, so that the first and the last pieces are short
static double stream_fraction(int k, int C) {
  const double x = (double)k / C;
  return C > 2 ? x * x * (3.0 - 2.0 * x) : x;
}


// make the main stream wait for a state upload still in flight on the copy stream
static int state_wait(a2ds_ctx *c) {
  if (c->state_pending) {
    CU(cudaStreamWaitEvent(c->stream, c->ev_state, 0));
    c->state_pending = false;
  }
  if (c->halo_pending) {  // a2ds_halo_forward was called while the upload was in flight
    c->halo_pending = false;
    if (halo_exchange(c, c->u, false)) return 1;
  }
  return 0;
}

template <class T>
static int upload(T **dst, const T *src, size_t n, cudaStream_t st) {
  if (*dst) cudaFree(*dst);
  *dst = nullptr;
  if (n == 0) return 0;
  CU(cudaMalloc((void **)dst, n * sizeof(T)));
  CU(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int a2ds_create(int device, a2ds_ctx **out) {
  A2DS_TRY
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail("a2ds_create: no CUDA device available (there is no CPU path)");
  if (device < 0 || device >= n || device >= MAX_DEVICES)
    return fail("a2ds_create: bad device index");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail("a2ds_create: kernels are built for sm_100a only, device is sm_" +
                std::to_string(prop.major) + std::to_string(prop.minor));
  a2ds_ctx *c = new a2ds_ctx();
  c->device = device;
  c->n_sm = prop.multiProcessorCount;
  if (const char *env = getenv("A2DS_WARPS_PER_BLOCK")) {
    const int v = atoi(env);
    if (v >= 1 && v <= MAX_WARPS_PER_BLOCK) c->warps_per_block_forced = v;
  }
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->zero_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->ev_zero_go, cudaEventDisableTiming));
  for (cudaEvent_t &e : c->ev_z) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  if (const char *env = getenv("A2DS_STREAM_ZERO")) c->stream_zero = atoi(env) != 0;
  if (const char *env = getenv("A2DS_DOUBLE_BUFFER")) c->double_buffer = atoi(env) != 0;
  if (const char *env = getenv("A2DS_STREAM_RESIDENT")) c->stream_resident = atoi(env) != 0;
  for (int k = 0; k < a2ds_ctx::MAX_CHUNKS; k++) CU(cudaEventCreateWithFlags(&c->ev_up[k], cudaEventDisableTiming));
  for (cudaEvent_t &e : c->ev_row) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  if (const char *env = getenv("A2DS_STREAM_CHUNKS"))
    c->stream_chunks = std::max(1, std::min((int)a2ds_ctx::MAX_CHUNKS, atoi(env)));
  if (const char *env = getenv("A2DS_STREAM_MIN_ELEMS")) c->stream_min_elems = std::max(1, atoi(env));
  CU(cudaEventCreateWithFlags(&c->ev_state, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_used, cudaEventDisableTiming));
  CU(cudaEventCreate(&c->ev0));
  CU(cudaEventCreate(&c->ev1));
  CU(cudaEventCreate(&c->evk0));
  CU(cudaEventCreate(&c->evk1));
  CU(cudaEventCreate(&c->evr0));
  CU(cudaEventCreate(&c->evr1));
  *out = c;
  return 0;
  A2DS_CATCH(a2ds_create)
}

static void free_lists(a2ds_ctx *c) {
  for (int k = 0; k < 4; k++) {
    for (int *p : c->list_dev[k])
      if (p) cudaFree(p);
    c->list_dev[k].clear();
    c->list_len[k].clear();
  }
  c->lists_ready = false;
  c->splan[0].ready = c->splan[1].ready = false;
}

extern "C" int a2ds_destroy(a2ds_ctx *c) {
  A2DS_TRY
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->copy_stream);
  cudaStreamSynchronize(c->d2h_stream);
  cudaStreamSynchronize(c->zero_stream);
  cudaStreamSynchronize(c->stream);
  for (auto &m : c->mats)
    for (void *p : m.owned) cudaFree(p);
  free_lists(c);
  cudaFree(c->conn); cudaFree(c->elem_comp); cudaFree(c->X); cudaFree(c->u); cudaFree(c->res);
  cudaFree(c->udd);
  cudaFree(c->scratch_x); cudaFree(c->scratch_y); cudaFree(c->zplan_dev); cudaFree(c->shape9_dev);
  cudaFree(c->nat_d_rowp); cudaFree(c->nat_d_cols);
  cudaFree(c->comps); cudaFree(c->bc_nodes); cudaFree(c->bc_vars); cudaFree(c->bc_vals);
  cudaFree(c->send_nodes); cudaFree(c->recv_nodes); cudaFree(c->send_buf); cudaFree(c->recv_buf);
  if (c->comm) ncclCommDestroy(c->comm);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->evk0);
  cudaEventDestroy(c->evk1); cudaEventDestroy(c->evr0); cudaEventDestroy(c->evr1);
  cudaEventDestroy(c->ev_state); cudaEventDestroy(c->ev_used);
  for (cudaEvent_t e : c->ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_row) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_z) cudaEventDestroy(e);
  cudaEventDestroy(c->ev_zero_go);
  cudaStreamDestroy(c->zero_stream);
  cudaStreamDestroy(c->d2h_stream);
  cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
  A2DS_CATCH(a2ds_destroy)
}

extern "C" int a2ds_synchronize(a2ds_ctx *c) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->copy_stream));
  CU(cudaStreamSynchronize(c->d2h_stream));
  CU(cudaStreamSynchronize(c->zero_stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
  A2DS_CATCH(a2ds_synchronize)
}

// Dependent nodes (TACSAssembler::setDependentNodes, src/TACSAssembler.cpp:716-775): the spec is
// stored and applied by the following a2ds_set_mesh calls (n_dep = 0 withdraws it), as the
// reference wants it set before initialize().
extern "C" int a2ds_set_dependent_nodes(a2ds_ctx *c, int n_dep, const int *dep_ptr, const int *dep_conn,
                                        const double *dep_weights) {
  A2DS_TRY
  if (n_dep < 0) return fail("a2ds_set_dependent_nodes: negative count");
  if (n_dep == 0) {
    c->h_dep_ptr.clear(); c->h_dep_conn.clear(); c->h_dep_w.clear();
    return 0;
  }
  if (!dep_ptr || !dep_conn || !dep_weights || dep_ptr[0] != 0)
    return fail("a2ds_set_dependent_nodes: dep_ptr must start at 0 and all three arrays be given");
  for (int d = 0; d < n_dep; d++)
    if (dep_ptr[d + 1] < dep_ptr[d]) return fail("a2ds_set_dependent_nodes: dep_ptr is not monotone");
  for (int j = 0; j < dep_ptr[n_dep]; j++)
    if (dep_conn[j] < 0)
      return fail("a2ds_set_dependent_nodes: a dependent node must refer to independent nodes only");
  c->h_dep_ptr.assign(dep_ptr, dep_ptr + n_dep + 1);
  c->h_dep_conn.assign(dep_conn, dep_conn + dep_ptr[n_dep]);
  c->h_dep_w.assign(dep_weights, dep_weights + dep_ptr[n_dep]);
  return 0;
  A2DS_CATCH(a2ds_set_dependent_nodes)
}

// rows of the dependent nodes of a node vector with nc values per node (see k_dep_gather)
static int dep_gather(a2ds_ctx *c, double *v, int nc) {
  if (!c->n_dep) return 0;
  k_dep_gather<<<(c->n_dep * nc + 127) / 128, 128, 0, c->stream>>>(c->n_dep, nc, c->n_nodes, c->dep_ptr,
                                                                   c->dep_conn, c->dep_w, v);
  c->last_launches++;
  return 0;
}

extern "C" int a2ds_set_mesh(a2ds_ctx *c, int n_nodes, int n_owned, int n_elems, const int *conn,
                             const int *elem_comp) {
  return a2ds_set_mesh_order(c, 2, n_nodes, n_owned, n_elems, conn, elem_comp);
}

extern "C" int a2ds_set_mesh_order(a2ds_ctx *c, int order, int n_nodes, int n_owned, int n_elems,
                                   const int *conn, const int *elem_comp) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (order != 2 && order != 3) return fail("a2ds_set_mesh_order: order must be 2 (4 nodes) or 3 (9 nodes)");
  if (n_owned > n_nodes || n_nodes < 0 || n_elems < 0) return fail("a2ds_set_mesh: bad sizes");
  const int npe = order * order;
  const int n_dep = c->h_dep_ptr.empty() ? 0 : (int)c->h_dep_ptr.size() - 1;
  for (size_t i = 0; i < npe * (size_t)n_elems; i++)
    if (conn[i] < -n_dep || conn[i] >= n_nodes)
      return fail(n_dep ? "a2ds_set_mesh: connectivity refers to a node outside [0, n_nodes) or to a "
                          "dependent node that was not declared"
                        : "a2ds_set_mesh: connectivity refers to a node outside [0, n_nodes) "
                          "(negative entries need a2ds_set_dependent_nodes first)");
  for (int v : c->h_dep_conn)
    if (v >= n_nodes) return fail("a2ds_set_mesh: a dependent node refers to a node outside [0, n_nodes)");
  if (c->mesh_set) {
    // a second mesh on the same context: nothing sized for the old one may survive — the
    // matrices (offset tables of 16 * old n_elems), the halo lists and the BC arrays
    CU(cudaStreamSynchronize(c->stream));
    for (auto &m : c->mats) {
      for (void *p : m.owned) cudaFree(p);
      m.owned.clear();
      m.A = nullptr; m.spare = nullptr; m.spare_clean = false; m.off = nullptr; m.blk_dev = nullptr; m.has_halo = false;
      m.dead = true;
    }
    c->has_halo = false;
    c->peers.clear(); c->send_ptr.clear(); c->recv_ptr.clear();
    c->n_bc = 0;
  }
  c->n_nodes = n_nodes; c->n_owned = n_owned; c->n_elems = n_elems; c->npe = npe;
  c->h_conn.assign(conn, conn + npe * (size_t)n_elems);
  c->n_dep = n_dep;
  c->h_dep_slots.clear();
  if (n_dep) {
    // dependent node d = connectivity entry -(d + 1) -> row n_nodes + d behind the local rows
    for (size_t e = 0; e < (size_t)n_elems; e++) {
      int *nd = &c->h_conn[npe * e];
      bool any = false;
      for (int i = 0; i < npe; i++)
        if (nd[i] < 0) { nd[i] = n_nodes + (-nd[i] - 1); any = true; }
      if (!any) continue;
      for (int slot = 0; slot < npe * npe; slot++)
        if (nd[slot / npe] >= n_nodes || nd[slot % npe] >= n_nodes)
          c->h_dep_slots.push_back((int)(e * npe * npe) + slot);
    }
    if (upload(&c->dep_ptr, c->h_dep_ptr.data(), c->h_dep_ptr.size(), c->stream)) return 1;
    if (upload(&c->dep_conn, c->h_dep_conn.data(), c->h_dep_conn.size(), c->stream)) return 1;
    if (upload(&c->dep_w, c->h_dep_w.data(), c->h_dep_w.size(), c->stream)) return 1;
  }
  c->nat_ready = false;
  cudaFree(c->nat_d_rowp); cudaFree(c->nat_d_cols);
  c->nat_d_rowp = c->nat_d_cols = nullptr; c->nat_hash = 0;
  if (elem_comp) c->h_elem_comp.assign(elem_comp, elem_comp + n_elems);
  else c->h_elem_comp.assign(n_elems, 0);
  if (upload(&c->conn, c->h_conn.data(), npe * (size_t)n_elems, c->stream)) return 1;
  if (upload(&c->elem_comp, c->h_elem_comp.data(), (size_t)n_elems, c->stream)) return 1;
  CU(cudaStreamSynchronize(c->copy_stream));
  c->state_pending = false;
  c->halo_pending = false;
  cudaFree(c->X); cudaFree(c->u); cudaFree(c->res); cudaFree(c->udd);
  c->udd = nullptr;
  c->X = c->u = c->res = nullptr;
  const size_t n_ext = std::max<size_t>(c->n_ext(), 1);   // dependent rows behind the local ones
  CU(cudaMalloc((void **)&c->X, 3 * n_ext * sizeof(double)));
  CU(cudaMalloc((void **)&c->u, 6 * n_ext * sizeof(double)));
  CU(cudaMalloc((void **)&c->res, 6 * n_ext * sizeof(double)));
  CU(cudaMemsetAsync(c->X, 0, 3 * n_ext * sizeof(double), c->stream));
  CU(cudaMemsetAsync(c->u, 0, 6 * n_ext * sizeof(double), c->stream));
  CU(cudaMemsetAsync(c->res, 0, 6 * n_ext * sizeof(double), c->stream));
  free_lists(c);
  c->mesh_set = true;
  return 0;
  A2DS_CATCH(a2ds_set_mesh)
}

extern "C" int a2ds_set_nodes(a2ds_ctx *c, const double *X) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (!c->X) return fail("a2ds_set_nodes: call a2ds_set_mesh first");
  CU(cudaMemcpyAsync(c->X, X, 3 * (size_t)c->n_nodes * sizeof(double), cudaMemcpyHostToDevice,
                     c->stream));
  // locations of the dependent nodes (TACSAssembler::setNodes distributes xptVec the same way)
  if (dep_gather(c, c->X, 3)) return 1;
  CU(cudaStreamSynchronize(c->stream));
  return 0;
  A2DS_CATCH(a2ds_set_nodes)
}

extern "C" int a2ds_set_components(a2ds_ctx *c, int n_comp, const double *Cs, const double *eth,
                                   const double *temperature, const int *elem_class,
                                   int transform, const double *ref_axis) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (n_comp <= 0) return fail("a2ds_set_components: need at least one component");
  if (transform != A2DS_TRANSFORM_NATURAL && transform != A2DS_TRANSFORM_REF_AXIS)
    return fail("a2ds_set_components: unknown transform kind");
  std::vector<CompData> h(n_comp);
  c->h_class.resize(n_comp);
  for (int i = 0; i < n_comp; i++) {
    memcpy(h[i].Cs, &Cs[22 * i], 22 * sizeof(double));
    memcpy(h[i].eth, &eth[9 * i], 9 * sizeof(double));
    h[i].temperature = temperature ? temperature[i] : 0.0;
    h[i].mom[0] = h[i].mom[1] = h[i].mom[2] = 0.0;
    if (i < (int)c->h_mom.size() / 3) memcpy(h[i].mom, &c->h_mom[3 * i], 3 * sizeof(double));
    h[i].model = elem_class ? elem_class[i] : 0;
    if (h[i].model != A2DS_QUAD4_SHELL && h[i].model != A2DS_QUAD4_NONLINEAR_SHELL)
      return fail("a2ds_set_components: unsupported element class (only TACSQuad4Shell and "
                  "TACSQuad4NonlinearShell are implemented)");
    h[i].coupled = 0;
    for (int k = 6; k < 12; k++)
      if (h[i].Cs[k] != 0.0) h[i].coupled = 1;
    c->h_class[i] = h[i].model + 2 * h[i].coupled;
    h[i].pad_ = 0;
    h[i].transform = transform;
    h[i].axis[0] = h[i].axis[1] = h[i].axis[2] = 0.0;
    if (transform == A2DS_TRANSFORM_REF_AXIS) {
      if (!ref_axis) return fail("a2ds_set_components: reference axis missing");
      // normalised as TACSShellRefAxisTransform's constructor does (Transform.h:97-108)
      double nrm = sqrt(ref_axis[0] * ref_axis[0] + ref_axis[1] * ref_axis[1] +
                        ref_axis[2] * ref_axis[2]);
      double inv = nrm != 0.0 ? 1.0 / nrm : 0.0;
      for (int k = 0; k < 3; k++) h[i].axis[k] = ref_axis[k] * inv;
    }
  }
  c->n_comp = n_comp;
  c->h_comps = h;
  if (upload(&c->comps, h.data(), (size_t)n_comp, c->stream)) return 1;
  free_lists(c);
  return 0;
  A2DS_CATCH(a2ds_set_components)
}

// mass moments per component, as TACSShellConstitutive::evalMassMoments returns them
// (TACSIsoShellConstitutive.cpp:120-129); may be called before or after
// a2ds_set_components
extern "C" int a2ds_set_mass_moments(a2ds_ctx *c, int n_comp, const double *moments) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (n_comp <= 0 || !moments) return fail("a2ds_set_mass_moments: bad arguments");
  c->h_mom.assign(moments, moments + 3 * (size_t)n_comp);
  if (!c->h_comps.empty()) {
    if ((int)c->h_comps.size() != n_comp)
      return fail("a2ds_set_mass_moments: component count differs from a2ds_set_components");
    for (int i = 0; i < n_comp; i++) memcpy(c->h_comps[i].mom, &moments[3 * i], 3 * sizeof(double));
    CU(cudaStreamSynchronize(c->stream));
    if (upload(&c->comps, c->h_comps.data(), (size_t)n_comp, c->stream)) return 1;
  }
  return 0;
  A2DS_CATCH(a2ds_set_mass_moments)
}

// time derivatives of the state (TACSAssembler::setVariables(q, qdot, qddot),
// src/TACSAssembler.cpp:3825-3857).  Only the second derivative enters this element class
// (inertial term); udot is accepted for signature parity and ignored.  uddot == NULL
// removes the inertial term again.
extern "C" int a2ds_set_state_rates(a2ds_ctx *c, int n_given, const double *udot,
                                    const double *uddot) {
  A2DS_TRY
  (void)udot;
  CU(cudaSetDevice(c->device));
  if (!c->mesh_set) return fail("a2ds_set_state_rates: mesh not set");
  if (!uddot) {
    if (c->udd) { CU(cudaStreamSynchronize(c->stream)); cudaFree(c->udd); c->udd = nullptr; }
    return 0;
  }
  if (n_given != c->n_nodes && n_given != c->n_owned)
    return fail("a2ds_set_state_rates: n_given must be n_nodes or n_owned");
  if (!c->udd) {
    CU(cudaMalloc((void **)&c->udd, std::max<size_t>(6 * c->n_ext() * sizeof(double), 8)));
    CU(cudaMemsetAsync(c->udd, 0, 6 * c->n_ext() * sizeof(double), c->stream));
  }
  CU(cudaMemcpyAsync(c->udd, uddot, 6 * (size_t)n_given * sizeof(double), cudaMemcpyHostToDevice,
                     c->stream));
  if (c->has_halo && n_given == c->n_owned && halo_exchange(c, c->udd, false)) return 1;
  return 0;
  A2DS_CATCH(a2ds_set_state_rates)
}

extern "C" int a2ds_set_state_dev(a2ds_ctx *c, int n_given, const double *u_dev) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (n_given != c->n_nodes && n_given != c->n_owned)
    return fail("a2ds_set_state: n_given must be n_nodes or n_owned");
  if (state_wait(c)) return 1;
  CU(cudaMemcpyAsync(c->u, u_dev, 6 * (size_t)n_given * sizeof(double), cudaMemcpyDeviceToDevice,
                     c->stream));
  return 0;
  A2DS_CATCH(a2ds_set_state_dev)
}

extern "C" int a2ds_set_state(a2ds_ctx *c, int n_given, const double *u) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (n_given != c->n_nodes && n_given != c->n_owned)
    return fail("a2ds_set_state: n_given must be n_nodes or n_owned");
  // after everything queued so far that reads the old state ...
  CU(cudaEventRecord(c->ev_used, c->stream));
  CU(cudaStreamWaitEvent(c->copy_stream, c->ev_used, 0));
  // in row chunks when the mesh is large enough for the streamed assembly to use them
  const int C = (c->n_elems >= c->stream_min_elems && n_given >= 64 * c->stream_chunks) ? c->stream_chunks : 1;
  for (int k = 0; k < C; k++) {
    const size_t lo = k == 0 ? 0 : (size_t)(stream_fraction(k, C) * n_given);
    const size_t hi = k == C - 1 ? (size_t)n_given : (size_t)(stream_fraction(k + 1, C) * n_given);
    CU(cudaMemcpyAsync(c->u + 6 * lo, u + 6 * lo, 6 * (hi - lo) * sizeof(double), cudaMemcpyHostToDevice,
                       c->copy_stream));
    CU(cudaEventRecord(c->ev_up[k], c->copy_stream));
    c->up_hi[k] = (int)hi;
  }
  c->up_chunks = C; c->up_rows = n_given;
  // ... and before the first consumer of the new one (state_wait)
  CU(cudaEventRecord(c->ev_state, c->copy_stream));
  c->state_pending = true;
  return 0;
  A2DS_CATCH(a2ds_set_state)
}

extern "C" int a2ds_set_bcs(a2ds_ctx *c, int n_bc, const int *nodes, const int *vars,
                            const double *vals) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (!c->mesh_set) return fail("a2ds_set_bcs: call a2ds_set_mesh first");
  if (n_bc < 0 || (n_bc > 0 && (!nodes || !vars))) return fail("a2ds_set_bcs: bad arguments");
  for (int b = 0; b < n_bc; b++) {
    if (nodes[b] < 0 || nodes[b] >= c->n_nodes)
      return fail("a2ds_set_bcs: boundary-condition node " + std::to_string(nodes[b]) +
                  " is outside [0, n_nodes)");
    // bits above the six shell DOFs are ignored (TACSBcMap carries them for larger blocks)
  }
  const std::vector<double> zeros(vals ? 0 : 6 * (size_t)n_bc, 0.0);  // NULL: homogeneous
  c->n_bc = n_bc;
  if (upload(&c->bc_nodes, nodes, (size_t)n_bc, c->stream)) return 1;
  if (upload(&c->bc_vars, vars, (size_t)n_bc, c->stream)) return 1;
  if (upload(&c->bc_vals, vals ? vals : zeros.data(), 6 * (size_t)n_bc, c->stream)) return 1;
  return 0;
  A2DS_CATCH(a2ds_set_bcs)
}

extern "C" int a2ds_set_double_buffer(a2ds_ctx *c, int on) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  c->double_buffer = on != 0;
  if (!on) {
    // give the spare value arrays back (after whatever still zeroes them)
    CU(cudaStreamSynchronize(c->stream));
    for (auto &m : c->mats) {
      if (!m.spare) continue;
      auto it = std::find(m.owned.begin(), m.owned.end(), (void *)m.spare);
      if (it != m.owned.end()) m.owned.erase(it);
      CU(cudaFree(m.spare));
      m.spare = nullptr; m.spare_clean = false;
    }
  }
  for (auto &m : c->mats) m.spare_refused = false;
  return 0;
  A2DS_CATCH(a2ds_set_double_buffer)
}

extern "C" int a2ds_set_scatter_mode(a2ds_ctx *c, int mode) {
  A2DS_TRY
  if (mode != A2DS_SCATTER_ATOMIC && mode != A2DS_SCATTER_COLORED &&
      mode != A2DS_SCATTER_ATOMIC_COLOR_ORDER)
    return fail("a2ds_set_scatter_mode: unknown mode");
  if (mode != c->scatter_mode) free_lists(c);
  c->scatter_mode = mode;
  return 0;
  A2DS_CATCH(a2ds_set_scatter_mode)
}

// Greedy element colouring: two elements that share a node get different colours, so
// within a colour every BCSR block and every residual row receives at most one
// contribution -> the sum order is fixed by the colour order.
static void color_elements(int nn, int ne, const int *conn, std::vector<int> &color,
                           int &n_colors) {
  std::vector<int> ptr(nn + 1, 0);
  for (size_t i = 0; i < 4 * (size_t)ne; i++) ptr[conn[i] + 1]++;
  for (int i = 0; i < nn; i++) ptr[i + 1] += ptr[i];
  std::vector<int> adj(ptr[nn]), fill(ptr.begin(), ptr.end() - 1);
  for (int e = 0; e < ne; e++)
    for (int i = 0; i < 4; i++) adj[fill[conn[4 * e + i]]++] = e;
  color.assign(ne, -1);
  n_colors = 0;
  // mark[col] == e: colour col is taken by a neighbour of element e (any number of colours:
  // high-valence fan / junction nodes can need more than a machine word of them)
  std::vector<int> mark;
  for (int e = 0; e < ne; e++) {
    for (int i = 0; i < 4; i++) {
      const int n = conn[4 * e + i];
      for (int k = ptr[n]; k < ptr[n + 1]; k++) {
        const int o = adj[k];
        if (color[o] >= 0) {
          if (color[o] >= (int)mark.size()) mark.resize(color[o] + 1, -1);
          mark[color[o]] = e;
        }
      }
    }
    int col = 0;
    while (col < (int)mark.size() && mark[col] == e) col++;
    color[e] = col;
    n_colors = std::max(n_colors, col + 1);
  }
}

// The colouring rule of the device (aux_kernels.cuh, k_color_round) stepped sequentially:
// greedy in decreasing order of color_key.  Bit-identical to the device's result.
static void color_elements_hashed(int nn, int ne, const int *conn, std::vector<int> &color,
                                  int &n_colors) {
  std::vector<int> ptr(nn + 1, 0);
  for (size_t i = 0; i < 4 * (size_t)ne; i++) ptr[conn[i] + 1]++;
  for (int i = 0; i < nn; i++) ptr[i + 1] += ptr[i];
  std::vector<int> adj(ptr[nn]), fill(ptr.begin(), ptr.end() - 1);
  for (int e = 0; e < ne; e++)
    for (int i = 0; i < 4; i++) adj[fill[conn[4 * e + i]]++] = e;
  std::vector<unsigned long long> order(ne);
  for (int e = 0; e < ne; e++) order[e] = color_key(e);
  std::sort(order.begin(), order.end(), std::greater<unsigned long long>());
  color.assign(ne, -1);
  n_colors = 0;
  std::vector<int> mark;
  for (int s = 0; s < ne; s++) {
    const int e = color_key_elem(order[s]);
    for (int i = 0; i < 4; i++) {
      const int n = conn[4 * e + i];
      for (int k = ptr[n]; k < ptr[n + 1]; k++) {
        const int o = adj[k];
        if (color[o] >= 0) {
          if (color[o] >= (int)mark.size()) mark.resize(color[o] + 1, -1);
          mark[color[o]] = e;
        }
      }
    }
    int col = 0;
    while (col < (int)mark.size() && mark[col] == e) col++;
    color[e] = col;
    n_colors = std::max(n_colors, col + 1);
  }
}

// Element colours of the context's mesh, computed on the device (k_pat_count / k_pat_fill give
// the elements around every node, k_color_round sweeps until nothing is left); the host copy is
// one download.  The reference has no counterpart: its element loop is sequential per thread
// and adds under a mutex (src/TACSAssembler.cpp:4561-4576).
static int color_elements_device(a2ds_ctx *c, std::vector<int> &color, int &n_colors) {
  const int nn = c->n_nodes, ne = c->n_elems;
  const size_t n4 = 4 * (size_t)ne;
  color.assign(ne, 0);
  n_colors = ne ? 1 : 0;
  if (!ne) return 0;
  int *deg = nullptr, *ptr = nullptr, *adj = nullptr, *col = nullptr, *left = nullptr;
  void *tmp = nullptr;
  auto cleanup = [&]() { cudaFree(deg); cudaFree(ptr); cudaFree(adj); cudaFree(col); cudaFree(left); cudaFree(tmp); };
  CU(cudaMalloc((void **)&deg, ((size_t)nn + 1) * sizeof(int)));
  CU(cudaMalloc((void **)&ptr, ((size_t)nn + 1) * sizeof(int)));
  CU(cudaMalloc((void **)&adj, n4 * sizeof(int)));
  CU(cudaMalloc((void **)&col, (size_t)ne * sizeof(int)));
  CU(cudaMalloc((void **)&left, sizeof(int)));
  CU(cudaMemsetAsync(deg, 0, ((size_t)nn + 1) * sizeof(int), c->stream));
  CU(cudaMemsetAsync(col, 0xff, (size_t)ne * sizeof(int), c->stream));
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, deg, ptr, nn + 1, c->stream);
  CU(cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)));
  const unsigned g4 = (unsigned)((n4 + 255) / 256), ge = (unsigned)((ne + 255) / 256);
  k_pat_count<<<g4, 256, 0, c->stream>>>(n4, c->conn, deg);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, deg, ptr, nn + 1, c->stream);
  CU(cudaMemsetAsync(deg, 0, ((size_t)nn + 1) * sizeof(int), c->stream));   // reused as the fill cursor
  k_pat_fill<<<g4, 256, 0, c->stream>>>(n4, c->conn, ptr, deg, adj);
  int remaining = ne, rounds = 0;
  while (remaining > 0) {
    CU(cudaMemsetAsync(left, 0, sizeof(int), c->stream));
    k_color_round<<<ge, 256, 0, c->stream>>>(ne, c->conn, ptr, adj, col, left);
    int now = 0;
    CU(cudaMemcpyAsync(&now, left, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (now >= remaining && ++rounds > 4) {   // cannot happen: the highest key always colours itself
      cleanup();
      return fail("element colouring on the device made no progress");
    }
    if (now < remaining) rounds = 0;
    remaining = now;
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(color.data(), col, (size_t)ne * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  cleanup();
  for (int e = 0; e < ne; e++) n_colors = std::max(n_colors, color[e] + 1);
  return 0;
}

// colours of the context's mesh, whichever way they are made (A2DS_HOST_COLORING=1: the host
// pass of the same rule; =2: the natural-order greedy pass of a2ds_host_color_elements)
static int mesh_colors(a2ds_ctx *c, std::vector<int> &color, int &n_colors) {
  static const int host_rule = getenv("A2DS_HOST_COLORING") ? atoi(getenv("A2DS_HOST_COLORING")) : 0;
  if (c->npe != 4) return fail("element colouring: 4-node meshes only");
  if (host_rule == 2) {   // natural-order greedy: sequential by nature, fewest colours on structured meshes
    color_elements(c->n_nodes, c->n_elems, c->h_conn.data(), color, n_colors);
    return 0;
  }
  if (host_rule) {
    color_elements_hashed(c->n_nodes, c->n_elems, c->h_conn.data(), color, n_colors);
    return 0;
  }
  return color_elements_device(c, color, n_colors);
}

static int build_lists(a2ds_ctx *c) {
  if (c->lists_ready) return 0;
  if (c->n_comp == 0) return fail("assemble: a2ds_set_components has not been called");
  if (c->n_dep && c->scatter_mode != A2DS_SCATTER_ATOMIC)
    return fail("assemble: meshes with dependent nodes are assembled with the atomic scatter only");
  for (int e = 0; e < c->n_elems; e++)
    if (c->h_elem_comp[e] < 0 || c->h_elem_comp[e] >= c->n_comp)
      return fail("assemble: element component index out of range");
  std::vector<int> color;
  int ncol = 1;
  if (c->scatter_mode != A2DS_SCATTER_ATOMIC) {
    StageTimer tm;
    if (mesh_colors(c, color, ncol)) return 1;
    tm.lap("element colouring");
  }
  // colour order in ONE launch: the lists of the colours are concatenated, so that elements
  // in flight together rarely share a node (fewer RED collisions in L2), without the launch
  // boundaries (and the reproducibility) of the coloured mode
  const bool one_launch = c->scatter_mode == A2DS_SCATTER_ATOMIC_COLOR_ORDER;
  const int ncol_color = ncol;
  if (one_launch) ncol = 1;
  c->n_colors = ncol;
  for (int k = 0; k < 4; k++) {
    std::vector<std::vector<int>> lists(ncol);
    if (one_launch) {
      for (int col = 0; col < ncol_color; col++)
        for (int e = 0; e < c->n_elems; e++)
          if (color[e] == col && c->h_class[c->h_elem_comp[e]] == k) lists[0].push_back(e);
    } else {
      for (int e = 0; e < c->n_elems; e++)
        if (c->h_class[c->h_elem_comp[e]] == k)
          lists[c->scatter_mode == A2DS_SCATTER_COLORED ? color[e] : 0].push_back(e);
    }
    c->list_dev[k].assign(ncol, nullptr);
    c->list_len[k].assign(ncol, 0);
    for (int col = 0; col < ncol; col++) {
      c->list_len[k][col] = (int)lists[col].size();
      if (lists[col].empty()) continue;
      // the identity list needs no indirection
      if ((int)lists[col].size() == c->n_elems && !one_launch) continue;
      if (upload(&c->list_dev[k][col], lists[col].data(), lists[col].size(), c->stream)) return 1;
    }
  }
  c->lists_ready = true;
  return 0;
}

extern "C" int a2ds_get_element_colors(a2ds_ctx *c, int *color, int *n_colors) {
  A2DS_TRY
  if (!c->mesh_set) return fail("a2ds_get_element_colors: call a2ds_set_mesh first");
  CU(cudaSetDevice(c->device));
  std::vector<int> col;
  int nc = 0;
  if (mesh_colors(c, col, nc)) return 1;
  if (color && !col.empty()) memcpy(color, col.data(), col.size() * sizeof(int));
  if (n_colors) *n_colors = nc;
  return 0;
  A2DS_CATCH(a2ds_get_element_colors)
}

// ---- dependent nodes: the fold plan of a matrix -------------------------------------------
// One BCSR block of a matrix as the host sees it (mirror of BlockDev)
struct HostBlock {
  const int *rowp, *cols, *row_map, *col_map;
  long long base;
  int nrows;
};
// block offset of node pair (rn, cn): the search of k_build_offsets on the host
static long long host_find_block(const std::vector<HostBlock> &hb, int rn, int cn) {
  for (const HostBlock &b : hb) {
    const int rr = b.row_map ? b.row_map[rn] : (rn < b.nrows ? rn : -1);
    const int cc = b.col_map ? b.col_map[cn] : cn;
    if (rr < 0 || rr >= b.nrows || cc < 0) continue;
    const int *lo = b.cols + b.rowp[rr], *hi = b.cols + b.rowp[rr + 1];
    const int *it = std::lower_bound(lo, hi, cc);
    if (it != hi && *it == cc) return b.base + (it - b.cols);
  }
  return -1;
}
// Every element node pair with a dependent node gets a scratch block behind the matrix
// (m.total + s, through the offset table) and the list of (independent row, independent column)
// blocks it is distributed to with the product of the weights: the varp / vars / weights
// expansion of TACSAssembler::addMatValues (src/TACSAssembler.h:485-510) done once per matrix.
static int build_dep_fold(a2ds_ctx *c, const std::vector<HostBlock> &hb, MatrixRec &m, const char *who) {
  const int ns = (int)c->h_dep_slots.size();
  if (!ns) return 0;
  const int npe = c->npe, n2 = npe * npe, nn = c->n_nodes;
  std::vector<int> fptr(1, 0), ftgt;
  std::vector<double> fw;
  auto expand = [&](int node, std::vector<std::pair<int, double>> &out) {
    out.clear();
    if (node < nn) { out.emplace_back(node, 1.0); return; }
    const int d = node - nn;
    for (int j = c->h_dep_ptr[d]; j < c->h_dep_ptr[d + 1]; j++) out.emplace_back(c->h_dep_conn[j], c->h_dep_w[j]);
  };
  std::vector<std::pair<int, double>> ri, cj;
  long long missing = 0;
  for (int s = 0; s < ns; s++) {
    const int pos = c->h_dep_slots[s], e = pos / n2, slot = pos - e * n2;
    expand(c->h_conn[npe * (size_t)e + slot / npe], ri);
    expand(c->h_conn[npe * (size_t)e + slot % npe], cj);
    for (auto &a : ri)
      for (auto &b : cj) {
        const long long k = host_find_block(hb, a.first, b.first);
        if (k < 0) { missing++; continue; }
        ftgt.push_back((int)k); fw.push_back(a.second * b.second);
      }
    fptr.push_back((int)ftgt.size());
  }
  if (missing)
    return fail(std::string(who) + ": the pattern misses " + std::to_string(missing) +
                " blocks between the independent nodes of dependent nodes");
  int *d_pos = nullptr;
  if (upload(&d_pos, c->h_dep_slots.data(), (size_t)ns, c->stream)) return 1;
  if (upload(&m.fold_ptr, fptr.data(), fptr.size(), c->stream)) return 1;
  if (upload(&m.fold_tgt, ftgt.data(), ftgt.size(), c->stream)) return 1;
  if (upload(&m.fold_w, fw.data(), fw.size(), c->stream)) return 1;
  m.owned.push_back(m.fold_ptr);
  if (m.fold_tgt) m.owned.push_back(m.fold_tgt);
  if (m.fold_w) m.owned.push_back(m.fold_w);
  k_dep_patch_offsets<<<(ns + 255) / 256, 256, 0, c->stream>>>(ns, d_pos, (int)m.total, m.off);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(d_pos);
  m.n_fold = ns;
  return 0;
}
// distribute and clear the scratch blocks of a matrix an element pass has added to
static int dep_fold(a2ds_ctx *c, MatrixRec &m) {
  if (!m.n_fold) return 0;
  const size_t nt = 36 * (size_t)m.n_fold;
  k_dep_fold<<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(m.n_fold, m.total, m.fold_ptr, m.fold_tgt,
                                                                 m.fold_w, m.A);
  c->last_launches++;
  return 0;
}

extern "C" int a2ds_mat_create(a2ds_ctx *c, int n_blocks, const int *nrows,
                               const int *const *rowp, const int *const *cols,
                               const int *const *row_map, const int *const *col_map,
                               const int *bc_ident, int *mat) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (n_blocks < 1 || n_blocks > 4) return fail("a2ds_mat_create: 1..4 BCSR blocks");
  if (!c->mesh_set) return fail("a2ds_mat_create: call a2ds_set_mesh first");
  StageTimer tm;
  MatrixRec m;
  m.n_blocks = n_blocks;
  std::vector<BlockDev> hb(n_blocks);
  long long total = 0;
  // the patterns come from the caller: every index the kernels will follow is checked here
  for (int b = 0; b < n_blocks; b++) {
    if (nrows[b] < 0) return fail("a2ds_mat_create: negative row count");
    if (nrows[b] > 0) {
      if (!rowp[b] || rowp[b][0] != 0) return fail("a2ds_mat_create: rowp must start at 0");
      for (int r = 0; r < nrows[b]; r++)
        if (rowp[b][r + 1] < rowp[b][r]) return fail("a2ds_mat_create: rowp is not monotone");
      if (rowp[b][nrows[b]] > 0 && !cols[b]) return fail("a2ds_mat_create: cols missing");
      for (int k = 0; k < rowp[b][nrows[b]]; k++)
        if (cols[b][k] < 0) return fail("a2ds_mat_create: negative column index");
    }
    if (row_map && row_map[b])
      for (int n = 0; n < c->n_nodes; n++)
        if (row_map[b][n] >= nrows[b]) return fail("a2ds_mat_create: row_map entry beyond the block's rows");
  }
  for (int b = 0; b < n_blocks; b++) {
    const long long nnz = nrows[b] > 0 ? rowp[b][nrows[b]] : 0;
    m.nnz.push_back(nnz); m.base.push_back(total); m.nrows.push_back(nrows[b]);
    m.h_rowp.push_back(std::make_shared<const std::vector<int>>(
        nrows[b] > 0 ? std::vector<int>(rowp[b], rowp[b] + nrows[b] + 1) : std::vector<int>(1, 0)));
    m.h_cols.push_back(std::make_shared<const std::vector<int>>(cols[b], cols[b] + nnz));
    int *d_rowp = nullptr, *d_cols = nullptr, *d_rm = nullptr, *d_cm = nullptr;
    std::vector<int> zero_rowp(1, 0);
    if (upload(&d_rowp, nrows[b] > 0 ? rowp[b] : zero_rowp.data(), (size_t)nrows[b] + 1, c->stream)) return 1;
    if (upload(&d_cols, cols[b], (size_t)nnz, c->stream)) return 1;
    if (row_map && row_map[b] && upload(&d_rm, row_map[b], (size_t)c->n_nodes, c->stream)) return 1;
    if (col_map && col_map[b] && upload(&d_cm, col_map[b], (size_t)c->n_nodes, c->stream)) return 1;
    m.owned.push_back(d_rowp); m.owned.push_back(d_cols);
    if (d_rm) m.owned.push_back(d_rm);
    if (d_cm) m.owned.push_back(d_cm);
    hb[b].rowp = d_rowp; hb[b].cols = d_cols; hb[b].row_map = d_rm; hb[b].col_map = d_cm;
    m.d_rowp.push_back(d_rowp); m.d_cols.push_back(d_cols);
    hb[b].base = total; hb[b].nrows = nrows[b];
    hb[b].ident = bc_ident ? bc_ident[b] : (b == 0);
    total += nnz;
  }
  if (total >= (1ll << 31)) return fail("a2ds_mat_create: more than 2^31 blocks on one GPU");
  tm.lap("mat_create: pattern upload");
  m.total = total;
  const long long n_scratch = (long long)c->h_dep_slots.size();   // fold scratch of the dependent nodes
  if (total + n_scratch >= (1ll << 31)) return fail("a2ds_mat_create: more than 2^31 blocks on one GPU");
  CU(cudaMalloc((void **)&m.A, std::max<long long>(total + n_scratch, 1) * 36 * sizeof(double)));
  m.owned.push_back(m.A);
  CU(cudaMemsetAsync(m.A, 0, (total + n_scratch) * 36 * sizeof(double), c->stream));
  CU(cudaMalloc((void **)&m.blk_dev, n_blocks * sizeof(BlockDev)));
  m.owned.push_back(m.blk_dev);
  CU(cudaMemcpyAsync(m.blk_dev, hb.data(), n_blocks * sizeof(BlockDev), cudaMemcpyHostToDevice,
                     c->stream));
  CU(cudaMalloc((void **)&m.off, std::max<size_t>(c->npe * c->npe * (size_t)c->n_elems, 1) * sizeof(int)));
  m.owned.push_back(m.off);
  int *d_missing = nullptr;
  CU(cudaMalloc((void **)&d_missing, sizeof(int)));
  CU(cudaMemsetAsync(d_missing, 0, sizeof(int), c->stream));
  if (c->n_elems > 0) {
    const size_t nt = c->npe * c->npe * (size_t)c->n_elems;
    k_build_offsets<<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(
        c->n_elems, c->npe, c->conn, n_blocks, m.blk_dev, m.off, d_missing, c->n_nodes);
    CU(cudaGetLastError());
  }
  int missing = 0;
  CU(cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  tm.lap("mat_create: alloc+zero+offsets");
  cudaFree(d_missing);
  if (missing) {
    for (void *p : m.owned) cudaFree(p);
    return fail("a2ds_mat_create: " + std::to_string(missing) +
                " element blocks have no entry in the supplied non-zero pattern");
  }
  if (n_scratch) {
    std::vector<HostBlock> hh(n_blocks);
    for (int b = 0; b < n_blocks; b++)
      hh[b] = HostBlock{m.h_rowp[b]->data(), m.h_cols[b]->data(), row_map ? row_map[b] : nullptr,
                        col_map ? col_map[b] : nullptr, m.base[b], nrows[b]};
    if (build_dep_fold(c, hh, m, "a2ds_mat_create")) {
      for (void *p : m.owned) cudaFree(p);
      return 1;
    }
  }
  c->mats.push_back(std::move(m));
  tm.lap("mat_create: record");
  *mat = (int)c->mats.size() - 1;
  return 0;
  A2DS_CATCH(a2ds_mat_create)
}

static int scratch_vectors(a2ds_ctx *c, size_t nx, size_t ny, double **dx, double **dy) {
  if (nx > c->scratch_nx) {
    cudaFree(c->scratch_x); c->scratch_x = nullptr; c->scratch_nx = 0;
    CU(cudaMalloc((void **)&c->scratch_x, std::max<size_t>(nx, 1) * sizeof(double)));
    c->scratch_nx = nx;
  }
  if (ny > c->scratch_ny) {
    cudaFree(c->scratch_y); c->scratch_y = nullptr; c->scratch_ny = 0;
    CU(cudaMalloc((void **)&c->scratch_y, std::max<size_t>(ny, 1) * sizeof(double)));
    c->scratch_ny = ny;
  }
  *dx = c->scratch_x; *dy = c->scratch_y;
  return 0;
}

static int check_mat(a2ds_ctx *c, int mat, int block = 0) {
  if (mat < 0 || mat >= (int)c->mats.size()) return fail("bad matrix id");
  if (c->mats[mat].dead) return fail("matrix id belongs to a mesh that has been replaced (a2ds_set_mesh)");
  if (block < 0 || block >= c->mats[mat].n_blocks) return fail("bad BCSR block index");
  return 0;
}

// node -> nodes of its elements (4 per incidence), then sort + unique per row.  The rows are
// independent once the incidences are bucketed: the sort / unique pass runs on all host cores
// (two sweeps: count the distinct columns of every row, then write them at their offsets).
static void sort_unique_rows(int nn, const std::vector<int> &ptr, std::vector<int> &tmp,
                             std::vector<int> &rowp, std::vector<int> &cols);
static int natural_pattern(int nn, int ne, const int *conn, std::vector<int> &rowp,
                           std::vector<int> &cols, int npe = 4) {
  std::vector<int> ptr(nn + 1, 0);
  for (size_t i = 0; i < npe * (size_t)ne; i++) {
    if (conn[i] < 0 || conn[i] >= nn) return fail("pattern: node index out of range");
    ptr[conn[i] + 1] += npe;
  }
  for (int i = 0; i < nn; i++) {
    if ((long long)ptr[i] + ptr[i + 1] > 0x7fffffffll) return fail("pattern too large");
    ptr[i + 1] += ptr[i];
  }
  std::vector<int> tmp(ptr[nn]), fill(ptr.begin(), ptr.end() - 1);
  for (int e = 0; e < ne; e++)
    for (int i = 0; i < npe; i++) {
      const int r = conn[npe * (size_t)e + i];
      for (int j = 0; j < npe; j++) tmp[fill[r]++] = conn[npe * (size_t)e + j];
    }
  fill.clear(); fill.shrink_to_fit();
  sort_unique_rows(nn, ptr, tmp, rowp, cols);
  return 0;
}

// rows of candidate columns (bucket r = tmp[ptr[r] .. ptr[r + 1])) -> sorted, distinct CSR
// (TacsSortAndUniquifyCSR, src/utils/TacsUtilities.cpp:280)
static void sort_unique_rows(int nn, const std::vector<int> &ptr, std::vector<int> &tmp,
                             std::vector<int> &rowp, std::vector<int> &cols) {
  rowp.assign(nn + 1, 0);
  const int nt = (int)std::max(1u, std::min(16u, nn < (1 << 16) ? 1u : std::thread::hardware_concurrency()));
  auto sweep = [&](bool write) {
    auto work = [&](int t) {
      const int r0 = (int)((long long)nn * t / nt), r1 = (int)((long long)nn * (t + 1) / nt);
      for (int r = r0; r < r1; r++) {
        int *b = tmp.data() + ptr[r], *e = tmp.data() + ptr[r + 1];
        if (!write) {
          std::sort(b, e);
          rowp[r + 1] = (int)(std::unique(b, e) - b);   // distinct columns now lead the bucket
        } else {
          std::copy(b, b + (rowp[r + 1] - rowp[r]), cols.begin() + rowp[r]);
        }
      }
    };
    if (nt == 1) { work(0); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
  };
  sweep(false);
  for (int r = 0; r < nn; r++) rowp[r + 1] += rowp[r];
  cols.assign((size_t)rowp[nn], 0);
  sweep(true);
}

// the same with dependent nodes: every independent node behind an element couples to every
// other one (TACSAssembler::computeLocalNodeToNodeCSR, src/TACSAssembler.cpp:1850-1935)
static int natural_pattern_dep(a2ds_ctx *c, std::vector<int> &rowp, std::vector<int> &cols) {
  const int nn = c->n_nodes, ne = c->n_elems, npe = c->npe;
  std::vector<int> eptr(ne + 1, 0), vars;
  for (int e = 0; e < ne; e++) {
    for (int i = 0; i < npe; i++) {
      const int v = c->h_conn[npe * (size_t)e + i];
      if (v < nn) vars.push_back(v);
      else vars.insert(vars.end(), c->h_dep_conn.begin() + c->h_dep_ptr[v - nn],
                       c->h_dep_conn.begin() + c->h_dep_ptr[v - nn + 1]);
    }
    if (vars.size() > 0x7fffffffull) return fail("pattern too large");
    eptr[e + 1] = (int)vars.size();
  }
  std::vector<long long> cnt(nn + 1, 0);
  for (int e = 0; e < ne; e++)
    for (int a = eptr[e]; a < eptr[e + 1]; a++) cnt[vars[a] + 1] += eptr[e + 1] - eptr[e];
  std::vector<int> ptr(nn + 1, 0);
  for (int i = 0; i < nn; i++) {
    cnt[i + 1] += cnt[i];
    if (cnt[i + 1] > 0x7fffffffll) return fail("pattern too large");
    ptr[i + 1] = (int)cnt[i + 1];
  }
  std::vector<int> tmp(ptr[nn]), fill(ptr.begin(), ptr.end() - 1);
  for (int e = 0; e < ne; e++)
    for (int a = eptr[e]; a < eptr[e + 1]; a++)
      for (int b = eptr[e]; b < eptr[e + 1]; b++) tmp[fill[vars[a]]++] = vars[b];
  sort_unique_rows(nn, ptr, tmp, rowp, cols);
  return 0;
}

extern "C" int a2ds_host_pattern(int n_nodes, int n_elems, const int *conn, int *rowp, int *cols,
                                 long long *nnz) {
  A2DS_TRY
  std::vector<int> rp, cl;
  if (natural_pattern(n_nodes, n_elems, conn, rp, cl)) return 1;
  if (rowp) memcpy(rowp, rp.data(), rp.size() * sizeof(int));
  if (cols) memcpy(cols, cl.data(), cl.size() * sizeof(int));
  if (nnz) *nnz = (long long)cl.size();
  return 0;
  A2DS_CATCH(a2ds_host_pattern)
}

extern "C" int a2ds_host_color_elements(int n_nodes, int n_elems, const int *conn, int *color,
                                        int *n_colors) {
  A2DS_TRY
  std::vector<int> col;
  int nc = 0;
  color_elements(n_nodes, n_elems, conn, col, nc);
  memcpy(color, col.data(), col.size() * sizeof(int));
  *n_colors = nc;
  return 0;
  A2DS_CATCH(a2ds_host_color_elements)
}

extern "C" int a2ds_host_color_elements_hashed(int n_nodes, int n_elems, const int *conn, int *color,
                                               int *n_colors) {
  A2DS_TRY
  std::vector<int> col;
  int nc = 0;
  color_elements_hashed(n_nodes, n_elems, conn, col, nc);
  if (!col.empty()) memcpy(color, col.data(), col.size() * sizeof(int));
  *n_colors = nc;
  return 0;
  A2DS_CATCH(a2ds_host_color_elements_hashed)
}

// natural-order pattern built on the device (aux_kernels.cuh, k_pat_*); the host copy is one
// download.  Returns 1 on error, 2 when a node has too many elements around it (host path).
static int natural_pattern_device(a2ds_ctx *c) {
  const int nn = c->n_nodes;
  const size_t n4 = 4 * (size_t)c->n_elems;
  int *deg = nullptr, *ptr = nullptr, *adj = nullptr, *cnt = nullptr, *ovf = nullptr;
  void *tmp = nullptr;
  auto cleanup = [&]() { cudaFree(deg); cudaFree(ptr); cudaFree(adj); cudaFree(cnt); cudaFree(ovf); cudaFree(tmp); };
  CU(cudaMalloc((void **)&deg, ((size_t)nn + 1) * sizeof(int)));
  CU(cudaMalloc((void **)&ptr, ((size_t)nn + 1) * sizeof(int)));
  CU(cudaMalloc((void **)&cnt, ((size_t)nn + 1) * sizeof(int)));
  CU(cudaMalloc((void **)&adj, std::max<size_t>(n4, 1) * sizeof(int)));
  CU(cudaMalloc((void **)&ovf, sizeof(int)));
  CU(cudaMemsetAsync(deg, 0, ((size_t)nn + 1) * sizeof(int), c->stream));
  CU(cudaMemsetAsync(cnt, 0, ((size_t)nn + 1) * sizeof(int), c->stream));
  CU(cudaMemsetAsync(ovf, 0, sizeof(int), c->stream));
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, deg, ptr, nn + 1, c->stream);
  CU(cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)));
  const unsigned g4 = (unsigned)((n4 + 255) / 256), gn = (unsigned)((nn + 127) / 128);
  if (n4) k_pat_count<<<g4, 256, 0, c->stream>>>(n4, c->conn, deg);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, deg, ptr, nn + 1, c->stream);
  CU(cudaMemsetAsync(deg, 0, ((size_t)nn + 1) * sizeof(int), c->stream));   // reused as the fill cursor
  if (n4) k_pat_fill<<<g4, 256, 0, c->stream>>>(n4, c->conn, ptr, deg, adj);
  if (nn) k_pat_rows<false><<<gn, 128, 0, c->stream>>>(nn, c->conn, ptr, adj, cnt, nullptr, nullptr, ovf);
  int overflow = 0;
  CU(cudaMemcpyAsync(&overflow, ovf, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  int *d_rowp = nullptr, *d_cols = nullptr;
  CU(cudaMalloc((void **)&d_rowp, ((size_t)nn + 1) * sizeof(int)));
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, d_rowp, nn + 1, c->stream);
  int nnz = 0;
  CU(cudaMemcpyAsync(&nnz, d_rowp + nn, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (overflow) { cleanup(); cudaFree(d_rowp); return 2; }
  CU(cudaMalloc((void **)&d_cols, std::max<size_t>((size_t)nnz, 1) * sizeof(int)));
  if (nn) k_pat_rows<true><<<gn, 128, 0, c->stream>>>(nn, c->conn, ptr, adj, nullptr, d_rowp, d_cols, ovf);
  CU(cudaGetLastError());
  auto hr = std::make_shared<std::vector<int>>((size_t)nn + 1);
  auto hc = std::make_shared<std::vector<int>>((size_t)nnz);
  CU(cudaMemcpyAsync(hr->data(), d_rowp, ((size_t)nn + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (nnz) CU(cudaMemcpyAsync(hc->data(), d_cols, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  cleanup();
  c->nat_rowp = hr; c->nat_cols = hc; c->nat_d_rowp = d_rowp; c->nat_d_cols = d_cols;
  return 0;
}

extern "C" int a2ds_mat_create_natural(a2ds_ctx *c, int *mat) {
  A2DS_TRY
  if (!c->mesh_set) return fail("a2ds_mat_create_natural: call a2ds_set_mesh first");
  CU(cudaSetDevice(c->device));
  StageTimer tm;
  if (!c->nat_ready) {
    static const bool host_only = getenv("A2DS_HOST_PATTERN") && atoi(getenv("A2DS_HOST_PATTERN")) != 0;
    // 9-node meshes and meshes with dependent nodes: host sweep
    int rc = (host_only || c->npe != 4 || c->n_dep) ? 2 : natural_pattern_device(c);
    if (rc == 1) return 1;
    if (rc == 2) {   // very high valence somewhere (or asked for): host sort / unique sweep
      auto hr = std::make_shared<std::vector<int>>(), hc = std::make_shared<std::vector<int>>();
      if (c->n_dep ? natural_pattern_dep(c, *hr, *hc)
                   : natural_pattern(c->n_nodes, c->n_elems, c->h_conn.data(), *hr, *hc, c->npe)) return 1;
      c->nat_rowp = hr; c->nat_cols = hc;
      if (upload(&c->nat_d_rowp, hr->data(), hr->size(), c->stream)) return 1;
      if (upload(&c->nat_d_cols, hc->data(), hc->size(), c->stream)) return 1;
    }
    c->nat_ready = true;
    tm.lap(rc == 2 ? "natural pattern (host)" : "natural pattern (device)");
  }
  // the matrices of one mesh share the pattern arrays, on the host and on the device
  MatrixRec m;
  m.n_blocks = 1;
  const long long nnz = (long long)c->nat_cols->size();
  m.nnz.push_back(nnz); m.base.push_back(0); m.nrows.push_back(c->n_nodes);
  m.h_rowp.push_back(c->nat_rowp); m.h_cols.push_back(c->nat_cols);
  m.d_rowp.push_back(c->nat_d_rowp); m.d_cols.push_back(c->nat_d_cols);
  m.total = nnz;
  BlockDev hb;
  hb.rowp = c->nat_d_rowp; hb.cols = c->nat_d_cols; hb.row_map = nullptr; hb.col_map = nullptr;
  hb.base = 0; hb.nrows = c->n_nodes; hb.ident = 1;
  const long long n_scratch = (long long)c->h_dep_slots.size();   // fold scratch of the dependent nodes
  if (nnz + n_scratch >= (1ll << 31)) return fail("a2ds_mat_create_natural: more than 2^31 blocks on one GPU");
  CU(cudaMalloc((void **)&m.A, std::max<long long>(nnz + n_scratch, 1) * 36 * sizeof(double)));
  m.owned.push_back(m.A);
  CU(cudaMemsetAsync(m.A, 0, (nnz + n_scratch) * 36 * sizeof(double), c->stream));
  CU(cudaMalloc((void **)&m.blk_dev, sizeof(BlockDev)));
  m.owned.push_back(m.blk_dev);
  CU(cudaMemcpyAsync(m.blk_dev, &hb, sizeof(BlockDev), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMalloc((void **)&m.off, std::max<size_t>(c->npe * c->npe * (size_t)c->n_elems, 1) * sizeof(int)));
  m.owned.push_back(m.off);
  int *d_missing = nullptr;
  CU(cudaMalloc((void **)&d_missing, sizeof(int)));
  CU(cudaMemsetAsync(d_missing, 0, sizeof(int), c->stream));
  if (c->n_elems > 0) {
    const size_t nt = c->npe * c->npe * (size_t)c->n_elems;
    k_build_offsets<<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(c->n_elems, c->npe, c->conn, 1,
                                                                        m.blk_dev, m.off, d_missing, c->n_nodes);
    CU(cudaGetLastError());
  }
  int missing = 0;
  CU(cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(d_missing);
  tm.lap("mat_create_natural: alloc+zero+offsets");
  if (missing) {
    for (void *p : m.owned) cudaFree(p);
    return fail("a2ds_mat_create_natural: the pattern misses " + std::to_string(missing) + " element blocks");
  }
  if (n_scratch) {
    std::vector<HostBlock> hh(1, HostBlock{c->nat_rowp->data(), c->nat_cols->data(), nullptr, nullptr, 0, c->n_nodes});
    if (build_dep_fold(c, hh, m, "a2ds_mat_create_natural")) {
      for (void *p : m.owned) cudaFree(p);
      return 1;
    }
  }
  m.shared_hash = &c->nat_hash;
  c->mats.push_back(std::move(m));
  *mat = (int)c->mats.size() - 1;
  return 0;
  A2DS_CATCH(a2ds_mat_create_natural)
}

extern "C" int a2ds_mat_pattern(a2ds_ctx *c, int mat, int block, int *nrows, int *rowp, int *cols) {
  A2DS_TRY
  if (check_mat(c, mat, block)) return 1;
  MatrixRec &m = c->mats[mat];
  if (nrows) *nrows = m.nrows[block];
  if (rowp) memcpy(rowp, m.h_rowp[block]->data(), m.h_rowp[block]->size() * sizeof(int));
  if (cols) memcpy(cols, m.h_cols[block]->data(), m.h_cols[block]->size() * sizeof(int));
  return 0;
  A2DS_CATCH(a2ds_mat_pattern)
}

extern "C" int a2ds_mat_nnz(a2ds_ctx *c, int mat, int block, long long *nnz) {
  A2DS_TRY
  if (check_mat(c, mat, block)) return 1;
  *nnz = c->mats[mat].nnz[block];
  return 0;
  A2DS_CATCH(a2ds_mat_nnz)
}

extern "C" int a2ds_mat_zero(a2ds_ctx *c, int mat) {
  A2DS_TRY
  if (check_mat(c, mat)) return 1;
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(c->mats[mat].A, 0, c->mats[mat].total * 36 * sizeof(double), c->stream));
  return 0;
  A2DS_CATCH(a2ds_mat_zero)
}

extern "C" int a2ds_mat_download(a2ds_ctx *c, int mat, int block, double *A) {
  A2DS_TRY
  if (check_mat(c, mat, block)) return 1;
  CU(cudaSetDevice(c->device));
  MatrixRec &m = c->mats[mat];
  CU(cudaMemcpyAsync(A, m.A + 36 * m.base[block], m.nnz[block] * 36 * sizeof(double),
                     cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
  A2DS_CATCH(a2ds_mat_download)
}

// selected block rows of one BCSR block to the host: the blocks of rows[0], rows[1], ... one
// after the other in A (36 doubles each, sum over the rows of rowp[r + 1] - rowp[r] blocks);
// what BCSRMat::getArrays + a row loop gives on the host side — for spot checks of matrices
// that are too large to copy back whole
extern "C" int a2ds_mat_download_rows(a2ds_ctx *c, int mat, int block, int n_rows, const int *rows,
                                      double *A) {
  A2DS_TRY
  if (check_mat(c, mat, block)) return 1;
  CU(cudaSetDevice(c->device));
  MatrixRec &m = c->mats[mat];
  const std::vector<int> &rowp = *m.h_rowp[block];
  size_t out = 0;
  for (int i = 0; i < n_rows; i++) {
    const int r = rows[i];
    if (r < 0 || r >= m.nrows[block]) return fail("a2ds_mat_download_rows: row out of range");
    const size_t nb = (size_t)(rowp[r + 1] - rowp[r]);
    if (nb)
      CU(cudaMemcpyAsync(A + 36 * out, m.A + 36 * ((size_t)m.base[block] + rowp[r]),
                         nb * 36 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    out += nb;
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
  A2DS_CATCH(a2ds_mat_download_rows)
}

extern "C" int a2ds_mat_values_dev(a2ds_ctx *c, int mat, int block, double **A_dev) {
  A2DS_TRY
  if (check_mat(c, mat, block)) return 1;
  *A_dev = c->mats[mat].A + 36 * c->mats[mat].base[block];
  return 0;
  A2DS_CATCH(a2ds_mat_values_dev)
}

// 64-bit FNV-1a over the block patterns: computed once per matrix so that copyValues / axpy
// do not compare megabytes of indices on the host at every call
static unsigned long long pattern_fingerprint(MatrixRec &m) {
  if (m.pattern_hash) return m.pattern_hash;
  if (m.shared_hash && *m.shared_hash) return m.pattern_hash = *m.shared_hash;
  unsigned long long h = 1469598103934665603ull;
  auto mix = [&](const std::vector<int> &v) {
    for (int x : v) { h ^= (unsigned)x; h *= 1099511628211ull; }
    h ^= 0xffu; h *= 1099511628211ull;
  };
  for (const auto &v : m.h_rowp) mix(*v);
  for (const auto &v : m.h_cols) mix(*v);
  m.pattern_hash = h ? h : 1;
  if (m.shared_hash) *m.shared_hash = m.pattern_hash;
  return m.pattern_hash;
}

static int same_pattern(a2ds_ctx *c, int a, int b) {
  if (check_mat(c, a) || check_mat(c, b)) return 1;
  MatrixRec &A = c->mats[a], &B = c->mats[b];
  if (A.n_blocks != B.n_blocks || A.total != B.total ||
      pattern_fingerprint(A) != pattern_fingerprint(B))
    return fail("matrix operation: the two matrices do not share one non-zero pattern");
  return 0;
}

extern "C" int a2ds_mat_copy(a2ds_ctx *c, int dst, int src) {
  A2DS_TRY
  if (same_pattern(c, dst, src)) return 1;
  CU(cudaSetDevice(c->device));
  const size_t n2 = (size_t)c->mats[src].total * 18;
  if (n2)
    k_axpy<true><<<(unsigned)((n2 + 1023) / 1024), 256, 0, c->stream>>>(
        n2, 1.0, reinterpret_cast<const double2 *>(c->mats[src].A),
        reinterpret_cast<double2 *>(c->mats[dst].A));
  CU(cudaGetLastError());
  return 0;
  A2DS_CATCH(a2ds_mat_copy)
}

extern "C" int a2ds_mat_axpy(a2ds_ctx *c, double alpha, int x, int y) {
  A2DS_TRY
  if (same_pattern(c, x, y)) return 1;
  CU(cudaSetDevice(c->device));
  const size_t n2 = (size_t)c->mats[x].total * 18;
  if (n2)
    k_axpy<false><<<(unsigned)((n2 + 1023) / 1024), 256, 0, c->stream>>>(
        n2, alpha, reinterpret_cast<const double2 *>(c->mats[x].A),
        reinterpret_cast<double2 *>(c->mats[y].A));
  CU(cudaGetLastError());
  return 0;
  A2DS_CATCH(a2ds_mat_axpy)
}

static int apply_mat_bcs(a2ds_ctx *c, int mat0, int mat1, int mat2);
extern "C" int a2ds_mat_apply_bcs(a2ds_ctx *c, int mat) {
  A2DS_TRY
  if (check_mat(c, mat)) return 1;
  CU(cudaSetDevice(c->device));
  if (apply_mat_bcs(c, mat, -1, -1)) return 1;
  CU(cudaGetLastError());
  return 0;
  A2DS_CATCH(a2ds_mat_apply_bcs)
}

extern "C" int a2ds_mat_mult_dev(a2ds_ctx *c, int mat, int block, const double *x_dev,
                                 double *y_dev) {
  A2DS_TRY
  if (check_mat(c, mat, block)) return 1;
  CU(cudaSetDevice(c->device));
  MatrixRec &m = c->mats[mat];
  const int nrows = m.nrows[block];
  if (nrows == 0) return 0;
  const size_t nt = 6 * (size_t)nrows;
  k_spmv6<<<(unsigned)((nt + 191) / 192), 192, 0, c->stream>>>(
      nrows, m.d_rowp[block], m.d_cols[block], m.A + 36 * m.base[block], x_dev, y_dev);
  CU(cudaGetLastError());
  return 0;
  A2DS_CATCH(a2ds_mat_mult_dev)
}

// Distributed mat-vec for a matrix assembled per rank over its local nodes (interface rows
// unassembled: the sum over the ranks of the local matrices is the global matrix; BCs applied
// per rank).  Owned rows of y on return equal TACSParallelMat::mult
// (src/bpmat/TACSParallelMat.cpp:248) without any exchange of matrix rows:
//   ghost entries of x <- owners, y = A_local x over ALL local rows, ghost rows of y -> owners
//   (add), y[bc] = x[bc].
extern "C" int a2ds_mat_mult_dist_dev(a2ds_ctx *c, int mat, double *x_dev, double *y_dev) {
  A2DS_TRY
  if (check_mat(c, mat, 0)) return 1;
  CU(cudaSetDevice(c->device));
  MatrixRec &m = c->mats[mat];
  if (m.n_blocks != 1 || m.nrows[0] != c->n_nodes)
    return fail("a2ds_mat_mult_dist: needs a natural-order matrix over the local nodes "
                "(a2ds_mat_create_natural)");
  if (m.has_halo)
    return fail("a2ds_mat_mult_dist: this matrix has a matrix halo (TACSParallelMat flavour): its "
                "owned rows are already complete and its ghost rows still hold the local shares, so "
                "the reverse vector exchange would count the interface couplings twice; use a "
                "matrix without a2ds_mat_set_halo (interface rows unassembled) for this product");
  if (halo_exchange(c, x_dev, false)) return 1;
  if (a2ds_mat_mult_dev(c, mat, 0, x_dev, y_dev)) return 1;
  if (halo_exchange(c, y_dev, true)) return 1;
  if (c->n_bc) {
    k_vec_copy_bcs<<<(6 * c->n_bc + 255) / 256, 256, 0, c->stream>>>(c->n_bc, c->bc_nodes, c->bc_vars,
                                                                     x_dev, y_dev, c->n_owned);
    CU(cudaGetLastError());
  }
  return 0;
  A2DS_CATCH(a2ds_mat_mult_dist_dev)
}

extern "C" int a2ds_mat_mult(a2ds_ctx *c, int mat, int block, int ncols, const double *x,
                             double *y) {
  A2DS_TRY
  if (check_mat(c, mat, block)) return 1;
  CU(cudaSetDevice(c->device));
  const int nrows = c->mats[mat].nrows[block];
  const std::vector<int> &hc = *c->mats[mat].h_cols[block];
  const int max_col = hc.empty() ? -1 : *std::max_element(hc.begin(), hc.end());
  if (ncols <= max_col || !x || !y)
    return fail("a2ds_mat_mult: x has " + std::to_string(ncols) + " block columns, the matrix "
                "block refers to column " + std::to_string(max_col));
  double *dx = nullptr, *dy = nullptr;
  if (scratch_vectors(c, 6 * (size_t)ncols, 6 * (size_t)nrows, &dx, &dy)) return 1;
  CU(cudaMemcpyAsync(dx, x, 6 * (size_t)ncols * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  int rc = a2ds_mat_mult_dev(c, mat, block, dx, dy);
  if (!rc) {
    CU(cudaMemcpyAsync(y, dy, 6 * (size_t)nrows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return rc;
  A2DS_CATCH(a2ds_mat_mult)
}

extern "C" int a2ds_res_dev(a2ds_ctx *c, double **r) { *r = c->res; return 0; }
extern "C" int a2ds_state_dev(a2ds_ctx *c, double **u) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (state_wait(c)) return 1;
  *u = c->u;
  return 0;
  A2DS_CATCH(a2ds_state_dev)
}

// ---- halo -------------------------------------------------------------------
extern "C" int a2ds_comm_unique_id(char id[128]) {
  A2DS_TRY
  ncclUniqueId uid;
  NC(ncclGetUniqueId(&uid));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id, &uid, 128);
  return 0;
  A2DS_CATCH(a2ds_comm_unique_id)
}

extern "C" int a2ds_comm_init(a2ds_ctx *c, int n_ranks, int rank, const char id[128]) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  NC(ncclCommInitRank(&c->comm, n_ranks, uid, rank));
  c->n_ranks = n_ranks; c->rank = rank;
  return 0;
  A2DS_CATCH(a2ds_comm_init)
}

extern "C" int a2ds_set_halo(a2ds_ctx *c, int n_peers, const int *peer_rank, const int *send_ptr,
                             const int *send_nodes, const int *recv_ptr, const int *recv_nodes) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (!c->mesh_set) return fail("a2ds_set_halo: call a2ds_set_mesh first");
  if (n_peers < 0 || !send_ptr || !recv_ptr || send_ptr[0] != 0 || recv_ptr[0] != 0)
    return fail("a2ds_set_halo: bad arguments");
  for (int p = 0; p < n_peers; p++) {
    if (peer_rank[p] < 0 || (c->comm && (peer_rank[p] >= c->n_ranks || peer_rank[p] == c->rank)))
      return fail("a2ds_set_halo: peer rank " + std::to_string(peer_rank[p]) + " is not another rank");
    if (send_ptr[p + 1] < send_ptr[p] || recv_ptr[p + 1] < recv_ptr[p])
      return fail("a2ds_set_halo: send_ptr / recv_ptr must not decrease");
  }
  // owners send owned nodes, ghosts are received: anything else would overwrite owned values
  for (int i = 0; i < send_ptr[n_peers]; i++)
    if (send_nodes[i] < 0 || send_nodes[i] >= c->n_owned)
      return fail("a2ds_set_halo: send list holds a node that is not an owned node");
  for (int i = 0; i < recv_ptr[n_peers]; i++)
    if (recv_nodes[i] < c->n_owned || recv_nodes[i] >= c->n_nodes)
      return fail("a2ds_set_halo: receive list holds a node that is not a ghost node");
  c->peers.assign(peer_rank, peer_rank + n_peers);
  c->send_ptr.assign(send_ptr, send_ptr + n_peers + 1);
  c->recv_ptr.assign(recv_ptr, recv_ptr + n_peers + 1);
  const size_t ns = send_ptr[n_peers], nr = recv_ptr[n_peers];
  c->h_send_nodes.assign(send_nodes, send_nodes + ns);
  c->splan[0].ready = c->splan[1].ready = false;
  if (upload(&c->send_nodes, send_nodes, ns, c->stream)) return 1;
  if (upload(&c->recv_nodes, recv_nodes, nr, c->stream)) return 1;
  cudaFree(c->send_buf); cudaFree(c->recv_buf);
  c->send_buf = c->recv_buf = nullptr;
  // the same buffers serve both directions: size for the larger side
  const size_t nb = std::max(ns, nr);
  CU(cudaMalloc((void **)&c->send_buf, std::max<size_t>(nb, 1) * 6 * sizeof(double)));
  CU(cudaMalloc((void **)&c->recv_buf, std::max<size_t>(nb, 1) * 6 * sizeof(double)));
  c->has_halo = n_peers > 0;
  return 0;
  A2DS_CATCH(a2ds_set_halo)
}

// forward: owners send the values of send_nodes, ghosts receive into recv_nodes
static int halo_exchange(a2ds_ctx *c, double *vec, bool reverse) {
  if (!c->has_halo) return 0;
  if (!c->comm) return fail("halo exchange: a2ds_comm_init has not been called");
  const int np = (int)c->peers.size();
  // forward packs the owned side, reverse packs the ghost side
  const std::vector<int> &pk_ptr = reverse ? c->recv_ptr : c->send_ptr;
  const std::vector<int> &up_ptr = reverse ? c->send_ptr : c->recv_ptr;
  const int *pk_nodes = reverse ? c->recv_nodes : c->send_nodes;
  const int *up_nodes = reverse ? c->send_nodes : c->recv_nodes;
  const int n_pk = pk_ptr[np], n_up = up_ptr[np];
  if (n_pk) k_pack6<<<(6 * n_pk + 255) / 256, 256, 0, c->stream>>>(n_pk, pk_nodes, vec, c->send_buf);
  NC(ncclGroupStart());
  for (int p = 0; p < np; p++) {
    const int ns = pk_ptr[p + 1] - pk_ptr[p], nr = up_ptr[p + 1] - up_ptr[p];
    if (ns) NC(ncclSend(c->send_buf + 6 * (size_t)pk_ptr[p], 6 * (size_t)ns, ncclDouble, c->peers[p], c->comm, c->stream));
    if (nr) NC(ncclRecv(c->recv_buf + 6 * (size_t)up_ptr[p], 6 * (size_t)nr, ncclDouble, c->peers[p], c->comm, c->stream));
  }
  NC(ncclGroupEnd());
  if (!reverse) {
    if (n_up)
      k_unpack6<<<(6 * n_up + 255) / 256, 256, 0, c->stream>>>(n_up, up_nodes, c->recv_buf, vec, 0);
    c->last_launches += (n_pk ? 1 : 0) + (n_up ? 1 : 0);
  } else {
    // reverse: a node shared by three or more ranks receives from several peers, so the
    // contributions are added one peer after the other, in the fixed peer order
    // (race free and deterministic; a node occurs at most once in one peer's list)
    for (int p = 0; p < np; p++) {
      const int nr = up_ptr[p + 1] - up_ptr[p];
      if (!nr) continue;
      k_unpack6<<<(6 * nr + 255) / 256, 256, 0, c->stream>>>(
          nr, up_nodes + up_ptr[p], c->recv_buf + 6 * (size_t)up_ptr[p], vec, 1);
      c->last_launches++;
    }
    c->last_launches += (n_pk ? 1 : 0);
  }
  CU(cudaGetLastError());
  return 0;
}

static int mat_halo_reverse(a2ds_ctx *c, MatrixRec &m) {
  if (!m.has_halo) return 0;
  if (!c->comm) return fail("matrix halo: a2ds_comm_init has not been called");
  const int np = (int)m.halo_peers.size();
  const int n_send = m.halo_send_ptr[np];
  if (n_send)
    k_pack36<<<(unsigned)((18 * (size_t)n_send + 255) / 256), 256, 0, c->stream>>>(
        n_send, m.halo_send_blk, reinterpret_cast<const double2 *>(m.A),
        reinterpret_cast<double2 *>(m.halo_send_buf));
  NC(ncclGroupStart());
  for (int p = 0; p < np; p++) {
    const int ns = m.halo_send_ptr[p + 1] - m.halo_send_ptr[p];
    const int nr = m.halo_recv_ptr[p + 1] - m.halo_recv_ptr[p];
    if (ns) NC(ncclSend(m.halo_send_buf + 36 * (size_t)m.halo_send_ptr[p], 36 * (size_t)ns, ncclDouble,
                        m.halo_peers[p], c->comm, c->stream));
    if (nr) NC(ncclRecv(m.halo_recv_buf + 36 * (size_t)m.halo_recv_ptr[p], 36 * (size_t)nr, ncclDouble,
                        m.halo_peers[p], c->comm, c->stream));
  }
  NC(ncclGroupEnd());
  for (int p = 0; p < np; p++) {  // one peer after the other: race free, fixed order
    const int nr = m.halo_recv_ptr[p + 1] - m.halo_recv_ptr[p];
    if (!nr) continue;
    k_unpack36_add<<<(unsigned)((18 * (size_t)nr + 255) / 256), 256, 0, c->stream>>>(
        nr, m.halo_recv_blk + m.halo_recv_ptr[p],
        reinterpret_cast<const double2 *>(m.halo_recv_buf + 36 * (size_t)m.halo_recv_ptr[p]),
        reinterpret_cast<double2 *>(m.A));
    c->last_launches++;
  }
  c->last_launches += n_send ? 1 : 0;
  CU(cudaGetLastError());
  return 0;
}

extern "C" int a2ds_mat_set_halo(a2ds_ctx *c, int mat, int n_peers, const int *peer_rank,
                                 const int *send_ptr, const int *send_blocks,
                                 const int *recv_ptr, const int *recv_blocks) {
  A2DS_TRY
  if (check_mat(c, mat)) return 1;
  CU(cudaSetDevice(c->device));
  MatrixRec &m = c->mats[mat];
  if (n_peers < 0) return fail("a2ds_mat_set_halo: bad peer count");
  const long long ns = n_peers ? send_ptr[n_peers] : 0, nr = n_peers ? recv_ptr[n_peers] : 0;
  for (long long i = 0; i < ns; i++)
    if (send_blocks[i] < 0 || send_blocks[i] >= m.total) return fail("a2ds_mat_set_halo: bad send block");
  for (long long i = 0; i < nr; i++)
    if (recv_blocks[i] < 0 || recv_blocks[i] >= m.total) return fail("a2ds_mat_set_halo: bad receive block");
  m.halo_peers.assign(peer_rank, peer_rank + n_peers);
  m.halo_send_ptr.assign(send_ptr, send_ptr + n_peers + 1);
  m.halo_recv_ptr.assign(recv_ptr, recv_ptr + n_peers + 1);
  if (n_peers == 0) { m.halo_send_ptr.assign(1, 0); m.halo_recv_ptr.assign(1, 0); }
  int *sb = nullptr, *rb = nullptr;
  if (upload(&sb, send_blocks, (size_t)ns, c->stream)) return 1;
  if (upload(&rb, recv_blocks, (size_t)nr, c->stream)) return 1;
  double *sbuf = nullptr, *rbuf = nullptr;
  CU(cudaMalloc((void **)&sbuf, std::max<size_t>(36 * (size_t)ns, 2) * sizeof(double)));
  CU(cudaMalloc((void **)&rbuf, std::max<size_t>(36 * (size_t)nr, 2) * sizeof(double)));
  m.halo_send_blk = sb; m.halo_recv_blk = rb; m.halo_send_buf = sbuf; m.halo_recv_buf = rbuf;
  if (sb) m.owned.push_back(sb);
  if (rb) m.owned.push_back(rb);
  m.owned.push_back(sbuf); m.owned.push_back(rbuf);
  m.has_halo = n_peers > 0;
  return 0;
  A2DS_CATCH(a2ds_mat_set_halo)
}

extern "C" int a2ds_halo_forward(a2ds_ctx *c) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  // With a host upload still in flight the exchange is queued behind it at the first use of
  // the state (every rank does the same, so the sends and receives still pair up): the
  // zeroing at the start of the next assembly then overlaps the upload on all ranks.
  if (c->state_pending) { c->halo_pending = true; return 0; }
  return halo_exchange(c, c->u, false);
  A2DS_CATCH(a2ds_halo_forward)
}

// ---- assembly ------------------------------------------------------------------
static const int MAX_ZERO_ROUNDS = 4096;   // rounds of the in-kernel zeroing (see ZeroPlan)
// double buffering: the share of the spare value arrays (c->pz_*) the launch about to be made
// zeroes on the side, about one round per trip of a warp; by value in the kernel parameters
static int spare_zero_plan(a2ds_ctx *c, KParams &p, int n_gw, int n_groups) {
  p.zval.rounds = 0;
  if (A2DS_ZWAIT != 2 || !(c->pz_K || c->pz_G)) return 0;
  n_gw = std::max(1, n_gw);
  int rounds = std::max(1, std::min(MAX_ZERO_ROUNDS, n_groups / n_gw));
  // a warp's chunk of one round stays below 256 MB (32-bit byte counts in the kernel): a short
  // launch in front of a large matrix takes more rounds (done at the end of the kernel), and
  // beyond MAX_ZERO_ROUNDS the array is zeroed by a memset after all
  const long long max_blocks = (256ll << 20) / 288;
  const long long need = (std::max(c->pz_K ? c->pz_nK : 0, c->pz_G ? c->pz_nG : 0) + max_blocks * n_gw - 1) /
                         (max_blocks * n_gw);
  if (need > MAX_ZERO_ROUNDS) {
    if (c->pz_K) CU(cudaMemsetAsync(c->pz_K, 0, c->pz_nK * 36 * sizeof(double), c->stream));
    if (c->pz_G) CU(cudaMemsetAsync(c->pz_G, 0, c->pz_nG * 36 * sizeof(double), c->stream));
    c->pz_K = c->pz_G = nullptr;
    return 0;
  }
  rounds = std::max(rounds, (int)need);
  ZeroPlan zp;
  memset(&zp, 0, sizeof(zp));
  zp.zK = (double2 *)c->pz_K; zp.nK = 18ll * c->pz_nK;
  zp.zG = (double2 *)c->pz_G; zp.nG = 18ll * c->pz_nG;
  const long long per_round = (long long)rounds * n_gw;
  zp.cK = 18 * (int)std::max<long long>(1, (c->pz_nK + per_round - 1) / per_round);   // whole blocks (18 double2)
  zp.cG = 18 * (int)std::max<long long>(1, (c->pz_nG + per_round - 1) / per_round);
  zp.rounds = rounds;
  p.zval = zp;
  c->pz_K = c->pz_G = nullptr;
  return 0;
}

template <bool RES, bool KMAT, bool GMAT, bool NL>
static int launch_one(a2ds_ctx *c, KParams &p) {
  const size_t raw = (GMAT || NL) ? sizeof(WarpScratch) : offsetof(WarpScratch, E2);
  const size_t per_warp = (raw + 15) & ~size_t(15);
  p.scratch_bytes = (int)per_warp;
  auto kern = k_assemble<RES, KMAT, GMAT, NL>;
  // pick the block size (1..4 warps) that keeps the most warps resident per SM: shared
  // memory is allocated per block, so smaller blocks pack better when it is the limiter
  // per instantiation AND per device: function attributes are device state
  static int best_wpb_dev[MAX_DEVICES] = {0}, best_per_sm_dev[MAX_DEVICES] = {0};
  int &best_wpb = best_wpb_dev[c->device], &best_per_sm = best_per_sm_dev[c->device];
  if (best_wpb == 0 || c->warps_per_block_forced) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(per_warp * MAX_WARPS_PER_BLOCK)));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    int best = 0;
    for (int wv = MAX_WARPS_PER_BLOCK; wv >= 1; wv--) {
      if (c->warps_per_block_forced && wv != c->warps_per_block_forced) continue;
      int per_sm = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wv * 32, per_warp * wv));
      if (per_sm * wv > best) { best = per_sm * wv; best_wpb = wv; best_per_sm = per_sm; }
    }
    if (best == 0) return fail("k_assemble does not fit on an SM");
    if (getenv("A2DS_VERBOSE"))
      fprintf(stderr, "[a2ds] k_assemble<%d,%d,%d,%d>: %zu B scratch/warp, %d warps/block x %d blocks/SM\n",
              (int)RES, (int)KMAT, (int)GMAT, (int)NL, per_warp, best_wpb, best_per_sm);
  }
  const int wpb = best_wpb, per_sm = best_per_sm;
  const size_t smem = per_warp * (size_t)wpb;
  if (!c->work_counter) CU(cudaMalloc((void **)&c->work_counter, (1 + MAX_ZERO_ROUNDS) * sizeof(int)));
  p.work_counter = c->work_counter;
  const int n_groups = (p.n_list + NB - 1) / NB;
  const int want = (n_groups + wpb - 1) / wpb;
  const int grid = std::max(1, std::min(want, c->n_sm * per_sm));
  p.zplan = nullptr;
  if (spare_zero_plan(c, p, grid * wpb, n_groups)) return 1;
  CU(cudaMemsetAsync(c->work_counter, 0, sizeof(int), c->stream));
  kern<<<grid, wpb * 32, smem, c->stream>>>(p);
  CU(cudaGetLastError());
  c->last_launches++;
  return 0;
}

// the tying-level kernel (components without membrane-bending coupling)
template <bool RES, bool KMAT, bool GMAT, bool NL>
static int launch_one_t(a2ds_ctx *c, KParams &p) {
  const size_t raw = sizeof(WarpScratchT);
  const size_t per_warp = (raw + 15) & ~size_t(15);
  p.scratch_bytes = (int)per_warp;
  auto kern = k_assemble_t<RES, KMAT, GMAT, NL>;
  static int best_wpb_dev[MAX_DEVICES] = {0}, best_per_sm_dev[MAX_DEVICES] = {0};
  int &best_wpb = best_wpb_dev[c->device], &best_per_sm = best_per_sm_dev[c->device];
  if (best_wpb == 0 || c->warps_per_block_forced) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(per_warp * MAX_WARPS_PER_BLOCK + BLOCK_SHARED_T)));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    int best = 0;
    for (int wv = MAX_WARPS_PER_BLOCK; wv >= 1; wv--) {
      if (c->warps_per_block_forced && wv != c->warps_per_block_forced) continue;
      int per_sm = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wv * 32, per_warp * wv + BLOCK_SHARED_T));
      if (per_sm * wv > best) { best = per_sm * wv; best_wpb = wv; best_per_sm = per_sm; }
    }
    if (best == 0) return fail("k_assemble_t does not fit on an SM");
    if (getenv("A2DS_VERBOSE"))
      fprintf(stderr, "[a2ds] k_assemble_t<%d,%d,%d,%d>: %zu B scratch/warp, %d warps/block x %d blocks/SM\n",
              (int)RES, (int)KMAT, (int)GMAT, (int)NL, per_warp, best_wpb, best_per_sm);
  }
  const int wpb = best_wpb, per_sm = best_per_sm;
  const size_t smem = per_warp * (size_t)wpb + BLOCK_SHARED_T;
  if (!c->work_counter) CU(cudaMalloc((void **)&c->work_counter, (1 + MAX_ZERO_ROUNDS) * sizeof(int)));
  p.work_counter = c->work_counter;
  const int n_groups = (p.n_list + NB - 1) / NB;
  const int want = (n_groups + wpb - 1) / wpb;
  const int grid = std::max(1, std::min(want, c->n_sm * per_sm));
  p.zplan = nullptr;
  if (spare_zero_plan(c, p, grid * wpb, n_groups)) return 1;
  int rounds = 0;
  if (A2DS_ZWAIT == 1 && (c->pz_K || c->pz_G)) {
    // in-kernel zeroing by rounds (see ZeroPlan): about one round per trip of a warp
    static const int ahead = getenv("A2DS_ZERO_AHEAD") ? atoi(getenv("A2DS_ZERO_AHEAD")) : 3;
    const int n_gw = grid * wpb;
    rounds = std::max(1, std::min(MAX_ZERO_ROUNDS, n_groups / n_gw));
    ZeroPlan zp;
    zp.zK = (double2 *)c->pz_K; zp.nK = 18ll * c->pz_nK;
    zp.zG = (double2 *)c->pz_G; zp.nG = 18ll * c->pz_nG;
    const long long per_round = (long long)rounds * n_gw;
    // whole blocks per chunk (18 double2 each)
    zp.cK = 18 * (int)std::max<long long>(1, (c->pz_nK + per_round - 1) / per_round);
    zp.cG = 18 * (int)std::max<long long>(1, (c->pz_nG + per_round - 1) / per_round);
    // the kernel takes ONE round index from the highest offset of both matrices: use the
    // smaller chunk (round boundaries in blocks)
    const double bpr = (double)n_gw * (double)(std::min(zp.zK ? zp.cK : zp.cG, zp.zG ? zp.cG : zp.cK) / 18);
    zp.inv_round = (float)(1.0 / bpr) * (1.0f + 1e-6f);
    zp.rounds = rounds; zp.ahead = std::max(1, ahead);
    zp.done = c->work_counter + 1;
    if (!c->zplan_dev) CU(cudaMalloc((void **)&c->zplan_dev, sizeof(ZeroPlan)));
    CU(cudaMemcpyAsync(c->zplan_dev, &zp, sizeof(ZeroPlan), cudaMemcpyHostToDevice, c->stream));
    p.zplan = c->zplan_dev;
    c->pz_K = c->pz_G = nullptr;
  }
  CU(cudaMemsetAsync(c->work_counter, 0, (1 + rounds) * sizeof(int), c->stream));
  if (p.zplan && A2DS_ZWAIT == 1) {
    // the zeroing protocol waits on every warp of the grid: all blocks must be co-resident
    void *args[] = {(void *)&p};
    CU(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(wpb * 32), args, smem, c->stream));
  } else {
    kern<<<grid, wpb * 32, smem, c->stream>>>(p);
  }
  CU(cudaGetLastError());
  c->last_launches++;
  return 0;
}

// element kernel of a class: tying-level formulation unless the section couples membrane and
// bending (or A2DS_FORMULATION=0 asks for the first formulation everywhere: A/B measurements)
template <bool RES, bool KMAT, bool GMAT, bool NL>
static int launch_elem(a2ds_ctx *c, KParams &p, bool coupled) {
  static const bool force_first = getenv("A2DS_FORMULATION") && atoi(getenv("A2DS_FORMULATION")) == 0;
  if (coupled || force_first) return launch_one<RES, KMAT, GMAT, NL>(c, p);
  return launch_one_t<RES, KMAT, GMAT, NL>(c, p);
}

// mass kernel of the 9-node shells (assemble9_kernels.cuh): one warp per element
template <bool RES, bool MAT>
static int launch_mass9(a2ds_ctx *c, KParams &p) {
  if (!c->shape9_dev) {
    CU(cudaMalloc((void **)&c->shape9_dev, sizeof(a2ds::Shape9)));
    k_shape9_tables<<<1, 64, 0, c->stream>>>((a2ds::Shape9 *)c->shape9_dev);
    CU(cudaGetLastError());
  }
  auto kern = k_mass9<RES, MAT>;
  const int wpb = 4;
  const size_t smem = wpb * sizeof(Mass9Warp);
  static bool attr_set[MAX_DEVICES] = {false};
  if (!attr_set[c->device]) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[c->device] = true;
  }
  if (!c->work_counter) CU(cudaMalloc((void **)&c->work_counter, (1 + MAX_ZERO_ROUNDS) * sizeof(int)));
  p.work_counter = c->work_counter;
  CU(cudaMemsetAsync(c->work_counter, 0, sizeof(int), c->stream));
  const int grid = std::max(1, std::min((p.n_list + wpb - 1) / wpb, c->n_sm * 4));
  kern<<<grid, wpb * 32, smem, c->stream>>>(p, (const a2ds::Shape9 *)c->shape9_dev);
  CU(cudaGetLastError());
  c->last_launches++;
  return 0;
}

// launch the mass kernel over one element list
template <bool RES, bool MAT>
static int launch_mass(a2ds_ctx *c, KParams &p) {
  if (c->npe == 9) return launch_mass9<RES, MAT>(c, p);
  const size_t per_warp = (offsetof(WarpScratch, E2) + 15) & ~size_t(15);
  p.scratch_bytes = (int)per_warp;
  auto kern = k_mass<RES, MAT>;
  static int per_sm_dev[MAX_DEVICES] = {0};  // function attributes are device state
  int &per_sm = per_sm_dev[c->device];
  const int wpb = MAX_WARPS_PER_BLOCK;
  if (per_sm == 0) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, per_warp * wpb));
    if (per_sm == 0) return fail("k_mass does not fit on an SM");
  }
  const int want = ((p.n_list + NB - 1) / NB + wpb - 1) / wpb;
  const int grid = std::max(1, std::min(want, c->n_sm * per_sm));
  if (!c->work_counter) CU(cudaMalloc((void **)&c->work_counter, (1 + MAX_ZERO_ROUNDS) * sizeof(int)));
  p.work_counter = c->work_counter;
  CU(cudaMemsetAsync(c->work_counter, 0, sizeof(int), c->stream));
  kern<<<grid, wpb * 32, per_warp * wpb, c->stream>>>(p);
  CU(cudaGetLastError());
  c->last_launches++;
  return 0;
}

// 9-node elements: one thread block per element (assemble9_kernels.cuh)
template <bool RES, bool KMAT, bool GMAT, bool NL>
static int launch_nine(a2ds_ctx *c, KParams &p) {
  auto kern = k_assemble9<RES, KMAT, GMAT, NL>;
  static int per_sm_dev[MAX_DEVICES] = {0};   // function attributes are device state
  int &per_sm = per_sm_dev[c->device];
  const size_t smem = (sizeof(Elem9Block) + 15) & ~size_t(15);
  if (per_sm == 0) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Q9_THREADS, smem));
    if (per_sm == 0) return fail("k_assemble9 does not fit on an SM");
    if (getenv("A2DS_VERBOSE"))
      fprintf(stderr, "[a2ds] k_assemble9<%d,%d,%d,%d>: %zu B shared per block, %d blocks/SM\n", (int)RES,
              (int)KMAT, (int)GMAT, (int)NL, smem, per_sm);
  }
  if (!c->work_counter) CU(cudaMalloc((void **)&c->work_counter, (1 + MAX_ZERO_ROUNDS) * sizeof(int)));
  p.work_counter = c->work_counter;
  CU(cudaMemsetAsync(c->work_counter, 0, sizeof(int), c->stream));
  const int grid = std::max(1, std::min(p.n_list, c->n_sm * per_sm));
  kern<<<grid, Q9_THREADS, smem, c->stream>>>(p);
  CU(cudaGetLastError());
  c->last_launches++;
  return 0;
}

// element kernels of one class list (p.elem_list / p.n_list set) for the outputs in `what`
static int launch_class(a2ds_ctx *c, KParams &p, int cls, int what) {
  if (c->npe == 9) {
    // TACSQuad9Shell / TACSQuad9NonlinearShell (any 22-entry section).  The geometric stiffness
    // is the linear-in-state term for both classes, as for the 4-node elements below.
    int rc = 0;
    if (cls & 1) {
      if (what & 1 || what & 2)
        rc = (what & 3) == 1 ? launch_nine<true, false, false, true>(c, p)
           : (what & 3) == 2 ? launch_nine<false, true, false, true>(c, p)
                             : launch_nine<true, true, false, true>(c, p);
    } else {
      if (what & 1 || what & 2)
        rc = (what & 3) == 1 ? launch_nine<true, false, false, false>(c, p)
           : (what & 3) == 2 ? launch_nine<false, true, false, false>(c, p)
                             : launch_nine<true, true, false, false>(c, p);
    }
    if (!rc && (what & 4)) rc = launch_nine<false, false, true, false>(c, p);
    return rc;
  }
  const bool cpl = cls >= 2;
  int rc = 0;
  if ((cls & 1) == 0) {
    switch (what) {
      case 0: break;
      case 1: rc = launch_elem<true, false, false, false>(c, p, cpl); break;
      case 2: rc = launch_elem<false, true, false, false>(c, p, cpl); break;
      case 3: rc = launch_elem<true, true, false, false>(c, p, cpl); break;
      case 4: rc = launch_elem<false, false, true, false>(c, p, cpl); break;
      case 7: rc = launch_elem<true, true, true, false>(c, p, cpl); break;
      default: return fail("assemble: unsupported output combination");
    }
  } else {
    // nonlinear strain model: residual and tangent about the current state.
    // Its geometric stiffness (a finite difference about the ZERO state in the
    // reference, TACSShellElement.h:705-751) is the same linear-in-state term as
    // for the linear model and is evaluated with the linear-model kernel.
    switch (what) {
      case 0: break;
      case 1: rc = launch_elem<true, false, false, true>(c, p, cpl); break;
      case 2: rc = launch_elem<false, true, false, true>(c, p, cpl); break;
      case 3: rc = launch_elem<true, true, false, true>(c, p, cpl); break;
      case 4: rc = launch_elem<false, false, true, false>(c, p, cpl); break;
      case 7:
        rc = launch_elem<true, true, false, true>(c, p, cpl);
        if (!rc) rc = launch_elem<false, false, true, false>(c, p, cpl);
        break;
      default: return fail("assemble: unsupported output combination");
    }
  }
  return rc;
}

// Streamed assembly plan: C element ranges in natural order (ranges that refer to ghost nodes
// last when a forward halo exchange is pending: it needs the whole upload), C chunks of the
// owned residual rows and, for each, the launch position after which nothing adds to it any
// more (C: only after the reverse halo exchange).
static void build_stream_plan(a2ds_ctx *c, bool ghost_last, StreamPlan &pl) {
  const int C = c->stream_chunks, ne = c->n_elems, no = c->n_owned;
  pl.C = C;
  pl.e0.assign(C + 1, 0);
  // short ranges at both ends (stream_fraction): what precedes the first range (its share of the
  // upload) and what follows the last one (its residual rows going home) is exposed, the rest is not
  for (int k = 1; k < C; k++) pl.e0[k] = (int)(stream_fraction(k, C) * ne) & ~63;   // whole batches, 16-byte aligned tables
  pl.e0[C] = ne;
  pl.max_node.assign(C, -1);
  const int *conn = c->h_conn.data();
  for (int k = 0; k < C; k++) {
    int mx = -1;
    for (size_t i = 4 * (size_t)pl.e0[k]; i < 4 * (size_t)pl.e0[k + 1]; i++) mx = std::max(mx, conn[i]);
    pl.max_node[k] = mx;
  }
  pl.order.clear();
  for (int k = 0; k < C; k++)
    if (!(ghost_last && pl.max_node[k] >= no)) pl.order.push_back(k);
  for (int k = 0; k < C; k++)
    if (ghost_last && pl.max_node[k] >= no) pl.order.push_back(k);
  std::vector<int> pos(C);
  for (int i = 0; i < C; i++) pos[pl.order[i]] = i;
  std::vector<signed char> node_pos((size_t)std::max(no, 1), -1);
  for (int k = 0; k < C; k++)
    for (size_t i = 4 * (size_t)pl.e0[k]; i < 4 * (size_t)pl.e0[k + 1]; i++) {
      const int n = conn[i];
      if (n < no && node_pos[n] < pos[k]) node_pos[n] = (signed char)pos[k];
    }
  if (c->has_halo)   // owned rows that receive ghost contributions from peers
    for (int n : c->h_send_nodes) node_pos[n] = (signed char)C;
  const int R = C * a2ds_ctx::ROWS_PER_CHUNK;
  pl.row_lo.assign(R + 1, 0);
  for (int j = 0; j <= R; j++) pl.row_lo[j] = (int)((long long)no * j / R);
  pl.row_pos.assign(R, 0);
  for (int j = 0; j < R; j++) {
    int mx = 0;
    for (int n = pl.row_lo[j]; n < pl.row_lo[j + 1]; n++) mx = std::max(mx, (int)node_pos[n]);
    pl.row_pos[j] = mx;
  }
  pl.ready = true;
}

// double buffering: the share of the spare arrays the next element launch zeroes, in proportion
// to the elements it processes (`upto_elems` of `all_elems` done once it has run)
static void spare_share(a2ds_ctx *c, long long upto_elems, long long all_elems) {
  const bool last = upto_elems >= all_elems;
  auto share = [&](double *sp, long long n, long long &done, double *&pz, long long &pzn) {
    if (!sp) return;
    const long long upto = last ? n : std::min(n, (long long)((double)n * (double)upto_elems / (double)all_elems));
    if (upto > done) { pz = sp + 36 * done; pzn = upto - done; done = upto; }
  };
  share(c->sp_K, c->sp_nK, c->sp_doneK, c->pz_K, c->pz_nK);
  share(c->sp_G, c->sp_nG, c->sp_doneG, c->pz_G, c->pz_nG);
}

struct AsmReq;
// the element launches, the residual halo, its boundary conditions and its way back of a
// streamed assembly (see a2ds_ctx::d2h_stream); the outputs are zeroed already
static int run_streamed(a2ds_ctx *c, const AsmReq &rq, KParams p, int cls, int what, bool zero_mats,
                        bool zero_K, bool zero_G);

// One assembly request.  what: bit 0 residual, bit 1 tangent, bit 2 geometric stiffness,
// bit 3 mass matrix (into mmat, which may be the tangent matrix: gamma term of the Jacobian).
struct AsmReq {
  int what = 0;
  double alpha = 1.0, gscale = 1.0, mscale = 1.0;
  int kmat = -1, gmat = -1, mmat = -1;
  bool zero = true;    // zero the outputs first
  bool finish = true;  // halo reverse exchange of the residual and boundary conditions
  double *res_host = nullptr;
};

// boundary conditions of up to three matrices in one launch (-1: none)
static int apply_mat_bcs(a2ds_ctx *c, int mat0, int mat1, int mat2) {
  if (!c->n_bc) return 0;
  BcMats M;
  int nm = 0, max_blocks = 0;
  for (int mat : {mat0, mat1, mat2}) {
    if (mat < 0) continue;
    MatrixRec &m = c->mats[mat];
    M.blk[nm] = m.blk_dev; M.A[nm] = m.A; M.n_blocks[nm] = m.n_blocks;
    max_blocks = std::max(max_blocks, m.n_blocks);
    nm++;
  }
  if (!nm) return 0;
  const long long nt = 32ll * c->n_bc * max_blocks;
  k_mat_bcs<<<dim3((unsigned)((nt + 127) / 128), nm), 128, 0, c->stream>>>(c->n_bc, c->bc_nodes, c->bc_vars, M);
  c->last_launches++;
  return 0;
}


static int run_streamed(a2ds_ctx *c, const AsmReq &rq, KParams p, int cls, int what, bool zero_mats,
                        bool zero_K, bool zero_G) {
  const bool RES = what & 1;
  const bool ghost_last = c->halo_pending;
  StreamPlan &pl = c->splan[ghost_last ? 1 : 0];
  if (!pl.ready || pl.C != c->stream_chunks) build_stream_plan(c, ghost_last, pl);
  const int C = pl.C;
  const bool pending = c->state_pending;
  int waited = -1;           // upload chunks [0, waited] are ordered before the main stream's next work
  bool halo_done = !c->halo_pending;
  // finish the row chunks that are final after launch position `at` and send them home
  const int R = (int)pl.row_pos.size();
  auto finish_rows = [&](int at) -> int {
    if (!RES) return 0;
    for (int j = 0; j < R; j++) {
      if (pl.row_pos[j] != at) continue;
      const int lo = pl.row_lo[j];
      while (j + 1 < R && pl.row_pos[j + 1] == at) j++;   // neighbours that are final together
      const int hi = pl.row_lo[j + 1];
      if (hi <= lo) continue;
      if (c->n_bc) {
        k_res_bcs<<<(6 * c->n_bc + 255) / 256, 256, 0, c->stream>>>(c->n_bc, c->bc_nodes, c->bc_vars,
                                                                    c->bc_vals, c->u, c->res, lo, hi);
        c->last_launches++;
      }
      if (rq.res_host) {
        CU(cudaEventRecord(c->ev_row[j], c->stream));
        CU(cudaStreamWaitEvent(c->d2h_stream, c->ev_row[j], 0));
        CU(cudaMemcpyAsync(rq.res_host + 6 * (size_t)lo, c->res + 6 * (size_t)lo,
                           6 * (size_t)(hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, c->d2h_stream));
      }
    }
    return 0;
  };
  const KParams base = p;
  long long elems_done = 0;
  if (zero_mats) {
    // all zeroing is queued now, in the order the element ranges need it: block rows up to the
    // highest node of range `at` are clean once ev_z[at] has fired; it runs next to the kernels
    // of the first ranges instead of in front of them
    CU(cudaEventRecord(c->ev_zero_go, c->stream));   // after everything that still reads the matrices
    CU(cudaStreamWaitEvent(c->zero_stream, c->ev_zero_go, 0));
    int zeroed = 0;
    auto zero_rows = [&](int hi) -> int {
      if (hi <= zeroed) return 0;
      for (int mid : {zero_K ? rq.kmat : -1, zero_G ? rq.gmat : -1}) {
        if (mid < 0) continue;
        MatrixRec &m = c->mats[mid];
        const int *rowp = m.h_rowp[0]->data();
        const size_t b0 = (size_t)rowp[zeroed], b1 = (size_t)rowp[hi];
        if (b1 > b0) CU(cudaMemsetAsync(m.A + 36 * b0, 0, (b1 - b0) * 36 * sizeof(double), c->zero_stream));
      }
      zeroed = hi;
      return 0;
    };
    for (int at = 0; at < C; at++) {
      if (zero_rows(std::min(c->n_nodes, pl.max_node[pl.order[at]] + 1))) return 1;
      CU(cudaEventRecord(c->ev_z[at], c->zero_stream));
    }
    if (zero_rows(c->n_nodes)) return 1;   // rows no element adds to
    CU(cudaEventRecord(c->ev_z[C], c->zero_stream));
  }
  for (int at = 0; at < C; at++) {
    const int k = pl.order[at];
    const int e0 = pl.e0[k], e1 = pl.e0[k + 1];
    const bool ghost = pl.max_node[k] >= c->n_owned;
    if (zero_mats) CU(cudaStreamWaitEvent(c->stream, c->ev_z[at], 0));
    if (pending) {
      // rows this range reads: up to its highest node, or everything that is coming
      const int need = (ghost && !halo_done) ? c->up_rows : std::min(pl.max_node[k] + 1, c->up_rows);
      int w = 0;
      while (w < c->up_chunks - 1 && c->up_hi[w] < need) w++;
      if (w > waited) { CU(cudaStreamWaitEvent(c->stream, c->ev_up[w], 0)); waited = w; }
    }
    if (ghost && !halo_done) {   // the deferred forward exchange: packs owned rows of the whole upload
      if (halo_exchange(c, c->u, false)) return 1;
      halo_done = true;
    }
    if (e1 > e0) {
      p = base;
      p.elem_list = nullptr; p.n_list = e1 - e0;
      p.conn = base.conn + 4 * (size_t)e0; p.elem_comp = base.elem_comp + e0;
      if (base.Koff) p.Koff = base.Koff + 16 * (size_t)e0;
      if (base.Goff) p.Goff = base.Goff + 16 * (size_t)e0;
      // this range's share of the spare arrays (double buffering), in proportion to its elements
      elems_done += e1 - e0;
      spare_share(c, elems_done, c->n_elems);
      if (launch_class(c, p, cls, what)) return 1;
    }
    if (finish_rows(at)) return 1;
  }
  if (zero_mats) CU(cudaStreamWaitEvent(c->stream, c->ev_z[C], 0));
  if (pending && waited < c->up_chunks - 1) CU(cudaStreamWaitEvent(c->stream, c->ev_state, 0));
  c->state_pending = false;
  if (!halo_done && halo_exchange(c, c->u, false)) return 1;
  c->halo_pending = false;
  CU(cudaEventRecord(c->evk1, c->stream));
  if (RES && halo_exchange(c, c->res, true)) return 1;
  return finish_rows(C);
}

static int run_assembly(a2ds_ctx *c, const AsmReq &rq) {
  CU(cudaSetDevice(c->device));
  if (!c->mesh_set) return fail("assemble: mesh or nodes not set");
  if (c->npe != 4 && c->scatter_mode != A2DS_SCATTER_ATOMIC)
    return fail("assemble: 9-node elements are assembled with the atomic scatter only");
  if (c->n_dep && c->scatter_mode != A2DS_SCATTER_ATOMIC)
    return fail("assemble: meshes with dependent nodes are assembled with the atomic scatter only");
  if (build_lists(c)) return 1;
  const int what = rq.what & 7;
  const bool RES = rq.what & 1, KM = (rq.what & 2) != 0, GM = (rq.what & 4) != 0,
             MM = (rq.what & 8) != 0;
  const int kmat = rq.kmat, gmat = rq.gmat, mmat = rq.mmat;
  // inertial residual M * uddot (TACSShellElement.h:410-447) once second derivatives are set
  const bool MRES = RES && c->udd != nullptr;
  if (KM && check_mat(c, kmat)) return 1;
  if (GM && check_mat(c, gmat)) return 1;
  if (MM && check_mat(c, mmat)) return 1;
  if (KM && GM && kmat == gmat && what == 7)
    return fail("assemble: tangent and geometric matrices must differ");
  c->pz_K = c->pz_G = nullptr;
  // streamed: one class, natural element order, atomic scatter, host I/O in the step
  int only_cls = -1, n_nonempty = 0;
  for (int cls = 0; cls < 4; cls++)
    if (c->list_len[cls][0] > 0) { n_nonempty++; only_cls = cls; }
  // ... and then natural-order matrices (block rows in node order, host row pointer at hand)
  // are zeroed range by range next to the kernels instead of in front of them; with
  // stream_resident a step without host I/O is split as well when it has such matrices to zero
  auto natural = [&](int mat) { return c->mats[mat].n_blocks == 1 && c->mats[mat].shared_hash != nullptr; };
  const bool zero_natural = c->stream_zero && (KM || GM) && (!KM || natural(kmat)) && (!GM || natural(gmat));
  const bool streamed = c->npe == 4 && A2DS_ZWAIT != 1 && c->stream_chunks > 1 && rq.zero && rq.finish &&
                        c->n_colors == 1 && n_nonempty == 1 && c->n_dep == 0 &&
                        c->list_dev[only_cls][0] == nullptr && c->n_elems >= c->stream_min_elems &&
                        !MM && !MRES && what != 0 &&
                        ((c->state_pending && c->up_chunks > 1) || (rq.res_host && RES) ||
                         (c->stream_resident && zero_natural));
  bool zero_streamed = streamed && zero_natural;
  c->sp_K = c->sp_G = nullptr;
  bool clean_K = false, clean_G = false;   // the matrix was swapped with its zeroed spare array
  if (rq.zero) {
    c->last_launches = 0;
    CU(cudaEventRecord(c->ev0, c->stream));
    if (RES) CU(cudaMemsetAsync(c->res, 0, 6 * c->n_ext() * sizeof(double), c->stream));
    // the tangent / geometric matrices are zeroed by the first element kernel itself when
    // that is k_assemble_t in a single launch per class (atomic scatter); otherwise here
    static const bool ikz_env = !(getenv("A2DS_INKERNEL_ZERO") && atoi(getenv("A2DS_INKERNEL_ZERO")) == 0);
    static const bool first_form = getenv("A2DS_FORMULATION") && atoi(getenv("A2DS_FORMULATION")) == 0;
    int first_cls = -1;
    for (int cls = 0; cls < 4 && first_cls < 0; cls++)
      if (c->list_len[cls][0] > 0) first_cls = cls;
    const bool ikz_built = A2DS_ZWAIT == 1;   // zero blocks compiled into k_assemble_t
    const bool ikz = ikz_built && ikz_env && !first_form && c->n_colors == 1 && (KM || GM) && first_cls >= 0 &&
                     first_cls < 2 && (what == 2 || what == 3 || what == 4 || what == 7);
    // double buffering: the first element kernel (k_assemble_t, one launch per class or per
    // element range) zeroes the spare array of every matrix it adds into; a matrix whose spare
    // array is clean is swapped with it instead of being zeroed.  Not for the tangent alone
    // (what == 2): that kernel is already bound by the read-modify-write traffic of its REDs and
    // loses more to the extra writes than the memset costs (measured: 2.59 -> 2.74 ms at 1 M
    // elements; residual + tangent 3.04 -> 2.85, geometric 4.27 -> 4.06, all three 5.76 -> 5.34)
    // (one launch per colour: only with the geometric stiffness in the pass — residual + tangent
    // in colour-sized launches loses to the extra writes like the tangent alone, 3.76 -> 4.00 ms)
    const bool dbuf = A2DS_ZWAIT == 2 && c->double_buffer && !first_form && c->npe == 4 &&
                      (c->n_colors == 1 || (what & 4)) &&
                      (KM || GM) && !MM && c->n_dep == 0 && first_cls >= 0 &&
                      !(KM && GM && kmat == gmat) && (what == 3 || what == 4 || what == 7);
    if (dbuf) {
      auto prepare = [&](int mat, double *&sp, long long &spn, bool &clean) -> int {
        MatrixRec &m = c->mats[mat];
        const size_t bytes = std::max<long long>(m.total, 1) * 36 * sizeof(double);
        if (!m.spare && !m.spare_refused) {
          size_t free_b = 0, total_b = 0;
          CU(cudaMemGetInfo(&free_b, &total_b));
          // leave room for everything else: no second array when it would take the last 15 %
          if (free_b < bytes + total_b * 15 / 100 || cudaMalloc((void **)&m.spare, bytes) != cudaSuccess) {
            (void)cudaGetLastError();
            m.spare = nullptr; m.spare_refused = true;
          } else {
            m.owned.push_back(m.spare);
            CU(cudaMemsetAsync(m.spare, 0, bytes, c->stream));
            m.spare_clean = true;
          }
        }
        if (!m.spare) return 0;
        if (m.spare_clean) { std::swap(m.A, m.spare); clean = true; }
        m.spare_clean = true;   // once the kernels queued below have run
        sp = m.spare; spn = m.total;
        return 0;
      };
      if (KM && prepare(kmat, c->sp_K, c->sp_nK, clean_K)) return 1;
      if (GM && prepare(gmat, c->sp_G, c->sp_nG, clean_G)) return 1;
      c->sp_doneK = c->sp_doneG = 0;
    }
    const bool set_K = KM && !clean_K, set_G = GM && !clean_G;
    if (zero_streamed && !set_K && !set_G) zero_streamed = false;
    if (ikz) {
      if (KM) { c->pz_K = c->mats[kmat].A; c->pz_nK = c->mats[kmat].total; }
      if (GM) { c->pz_G = c->mats[gmat].A; c->pz_nG = c->mats[gmat].total; }
    } else if (!zero_streamed) {
      if (set_K) CU(cudaMemsetAsync(c->mats[kmat].A, 0, c->mats[kmat].total * 36 * sizeof(double), c->stream));
      if (set_G) CU(cudaMemsetAsync(c->mats[gmat].A, 0, c->mats[gmat].total * 36 * sizeof(double), c->stream));
    }
    if (MM && !(KM && mmat == kmat))
      CU(cudaMemsetAsync(c->mats[mmat].A, 0, c->mats[mmat].total * 36 * sizeof(double), c->stream));
    c->last_launches += (RES ? 1 : 0) + (ikz ? 0 : (set_K ? 1 : 0) + (set_G ? 1 : 0)) + (MM && !(KM && mmat == kmat) ? 1 : 0);
  }

  KParams p;
  memset(&p, 0, sizeof(p));
  p.conn = c->conn; p.elem_comp = c->elem_comp; p.comps = c->comps;
  p.X = c->X; p.u = c->u; p.res = c->res; p.alpha = rq.alpha; p.gscale = rq.gscale;
  p.res_scale = 1.0; p.thermal = 1.0;
  if (KM) { p.Kval = c->mats[kmat].A; p.Koff = c->mats[kmat].off; }
  if (GM) { p.Gval = c->mats[gmat].A; p.Goff = c->mats[gmat].off; }

  CU(cudaEventRecord(c->evk0, c->stream));
  if (streamed) {
    if (run_streamed(c, rq, p, only_cls, what, zero_streamed, KM && !clean_K, GM && !clean_G)) return 1;
  } else {
    if (state_wait(c)) return 1;  // the upload overlapped the zeroing above
    // state (and accelerations) of the dependent nodes: TACSBVec::endDistributeValues
    if (c->n_dep && (dep_gather(c, c->u, 6) || (c->udd && dep_gather(c, c->udd, 6)))) return 1;
    long long elems_done = 0;
    for (int col = 0; col < c->n_colors; col++) {
      for (int cls = 0; cls < 4; cls++) {
        p.n_list = c->list_len[cls][col];
        p.elem_list = c->list_dev[cls][col];
        if (p.n_list == 0) continue;
        elems_done += p.n_list;
        spare_share(c, elems_done, c->n_elems);   // double buffering: every class launch its share
        if (launch_class(c, p, cls, what)) return 1;
        if (MM || MRES) {
          // mass path: same element lists (and colours), both element classes alike
          KParams pm = p;
          pm.u = c->udd; pm.alpha = rq.mscale;
          int rc = 0;
          if (MM) { pm.Kval = c->mats[mmat].A; pm.Koff = c->mats[mmat].off; }
          if (MM && MRES) rc = launch_mass<true, true>(c, pm);
          else if (MM) rc = launch_mass<false, true>(c, pm);
          else rc = launch_mass<true, false>(c, pm);
          if (rc) return rc;
        }
      }
    }
  }
  // double buffering: whatever part of a spare array no launch took (should not happen: the
  // element lists cover the mesh) is zeroed here, never left stale
  if (c->sp_K && c->sp_doneK < c->sp_nK)
    CU(cudaMemsetAsync(c->sp_K + 36 * c->sp_doneK, 0, (c->sp_nK - c->sp_doneK) * 36 * sizeof(double), c->stream));
  if (c->sp_G && c->sp_doneG < c->sp_nG)
    CU(cudaMemsetAsync(c->sp_G + 36 * c->sp_doneG, 0, (c->sp_nG - c->sp_doneG) * 36 * sizeof(double), c->stream));
  c->sp_K = c->sp_G = nullptr;
  if (!streamed) CU(cudaEventRecord(c->evk1, c->stream));
  if (c->n_dep) {
    // what the elements added to dependent rows / node pairs goes to the independent nodes with
    // the weights (TACSBVec::beginSetValues, TACSAssembler::addMatValues); the scratch is cleared
    if (RES) {
      k_dep_scatter<<<(c->n_dep * 6 + 127) / 128, 128, 0, c->stream>>>(c->n_dep, 6, c->n_nodes, c->dep_ptr,
                                                                      c->dep_conn, c->dep_w, c->res);
      c->last_launches++;
    }
    if (KM && dep_fold(c, c->mats[kmat])) return 1;
    if (GM && !(KM && gmat == kmat) && dep_fold(c, c->mats[gmat])) return 1;
    if (MM && !(KM && mmat == kmat) && !(GM && mmat == gmat) && dep_fold(c, c->mats[mmat])) return 1;
  }
  if (!rq.finish) return 0;
  // ghost residual contributions -> owners (TACSBVec::beginSetValues/endSetValues, ADD)
  if (!streamed) {
    if (RES && halo_exchange(c, c->res, true)) return 1;
    if (RES && c->n_bc) {
      k_res_bcs<<<(6 * c->n_bc + 255) / 256, 256, 0, c->stream>>>(c->n_bc, c->bc_nodes, c->bc_vars,
                                                                  c->bc_vals, c->u, c->res, 0, c->n_owned);
      c->last_launches++;
    }
  }
  // ghost block rows -> owners (only for matrices with a halo plan), then the BCs
  if (KM && mat_halo_reverse(c, c->mats[kmat])) return 1;
  if (GM && !(KM && gmat == kmat) && mat_halo_reverse(c, c->mats[gmat])) return 1;
  if (MM && !(KM && mmat == kmat) && !(GM && mmat == gmat) && mat_halo_reverse(c, c->mats[mmat]))
    return 1;
  if (apply_mat_bcs(c, KM ? kmat : -1, (GM && !(KM && gmat == kmat)) ? gmat : -1,
                    (MM && !(KM && mmat == kmat) && !(GM && mmat == gmat)) ? mmat : -1)) return 1;
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev1, c->stream));
  if (rq.res_host) {
    if (streamed && RES) {   // the rows went back chunk by chunk
      CU(cudaStreamSynchronize(c->d2h_stream));
    } else {
      CU(cudaMemcpyAsync(rq.res_host, c->res, 6 * (size_t)c->n_owned * sizeof(double),
                         cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

static int run_assembly(a2ds_ctx *c, int what, double alpha, int kmat, int gmat, double *res_host) {
  AsmReq rq;
  rq.what = what; rq.alpha = alpha; rq.kmat = kmat; rq.gmat = gmat; rq.res_host = res_host;
  return run_assembly(c, rq);
}

extern "C" int a2ds_assemble_res(a2ds_ctx *c, double *res) {
  A2DS_TRY
  return run_assembly(c, 1, 1.0, -1, -1, res);
  A2DS_CATCH(a2ds_assemble_res)
}

// beta multiplies dR/d(udot): this element class has no velocity dependent term (the
// reference passes beta on to the director, where TACSLinearizedRotation ignores it,
// TACSDirector.h:369-486), so any beta is accepted and has no effect.
extern "C" int a2ds_assemble_jacobian(a2ds_ctx *c, double alpha, double beta, double gamma,
                                      double *res, int mat) {
  A2DS_TRY
  (void)beta;
  AsmReq rq;
  rq.what = 3 | (gamma != 0.0 ? 8 : 0);
  rq.alpha = alpha; rq.mscale = gamma; rq.kmat = mat; rq.mmat = mat; rq.res_host = res;
  return run_assembly(c, rq);
  A2DS_CATCH(a2ds_assemble_jacobian)
}

extern "C" int a2ds_assemble_mat_type(a2ds_ctx *c, int mat_type, int mat) {
  A2DS_TRY
  if (mat_type == A2DS_STIFFNESS_MATRIX) return run_assembly(c, 2, 1.0, mat, -1, nullptr);
  if (mat_type == A2DS_GEOMETRIC_STIFFNESS_MATRIX) return run_assembly(c, 4, 1.0, -1, mat, nullptr);
  if (mat_type == A2DS_MASS_MATRIX) {
    AsmReq rq;
    rq.what = 8; rq.mmat = mat;
    return run_assembly(c, rq);
  }
  return fail("a2ds_assemble_mat_type: unknown matrix type");
  A2DS_CATCH(a2ds_assemble_mat_type)
}

// TACSAssembler::assembleMatCombo (src/TACSAssembler.cpp:4264-4318): A = sum_i scale[i] *
// matType[i], boundary conditions applied once at the end.
extern "C" int a2ds_assemble_mat_combo(a2ds_ctx *c, int n, const int *mat_types,
                                       const double *scales, int mat) {
  A2DS_TRY
  if (n <= 0) return fail("a2ds_assemble_mat_combo: need at least one matrix type");
  if (check_mat(c, mat)) return 1;
  for (int i = 0; i < n; i++) {
    AsmReq rq;
    rq.zero = (i == 0); rq.finish = (i == n - 1);
    if (mat_types[i] == A2DS_STIFFNESS_MATRIX) { rq.what = 2; rq.kmat = mat; rq.alpha = scales[i]; }
    else if (mat_types[i] == A2DS_GEOMETRIC_STIFFNESS_MATRIX) { rq.what = 4; rq.gmat = mat; rq.gscale = scales[i]; }
    else if (mat_types[i] == A2DS_MASS_MATRIX) { rq.what = 8; rq.mmat = mat; rq.mscale = scales[i]; }
    else return fail("a2ds_assemble_mat_combo: unknown matrix type");
    if (run_assembly(c, rq)) return 1;
  }
  return 0;
  A2DS_CATCH(a2ds_assemble_mat_combo)
}

extern "C" int a2ds_assemble_all(a2ds_ctx *c, double *res, int kmat, int gmat) {
  A2DS_TRY
  return run_assembly(c, 7, 1.0, kmat, gmat, res);
  A2DS_CATCH(a2ds_assemble_all)
}

// TACSAssembler::addJacobianVecProduct (src/TACSAssembler.cpp:4331-4391), matrix free:
// y <- y + scale * (alpha K) x, then the BC rows of y are zeroed.  For the linear strain
// model K does not depend on the state and K x is the element residual evaluated at x
// without thermal strain: r(x) = sum w B^T C B x — the residual kernel with u := x.
extern "C" int a2ds_add_jacobian_vec_product_dev(a2ds_ctx *c, double scale, double alpha,
                                                 const double *x_dev, double *y_dev) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  if (!c->mesh_set) return fail("addJacobianVecProduct: mesh or nodes not set");
  if (c->npe != 4) return fail("addJacobianVecProduct: 4-node elements only");
  if (c->n_dep) return fail("addJacobianVecProduct: not available on meshes with dependent nodes "
                            "(assemble the tangent and use a2ds_mat_mult)");
  if (((uintptr_t)x_dev & 15) || ((uintptr_t)y_dev & 7))
    return fail("addJacobianVecProduct: x must be 16-byte aligned (rows are fetched with 16-byte cp.async)");
  if (build_lists(c)) return 1;
  if (state_wait(c)) return 1;
  c->last_launches = 0;
  CU(cudaEventRecord(c->ev0, c->stream));
  KParams p;
  memset(&p, 0, sizeof(p));
  p.conn = c->conn; p.elem_comp = c->elem_comp; p.comps = c->comps;
  p.X = c->X; p.gscale = 1.0;
  CU(cudaEventRecord(c->evk0, c->stream));
  for (int col = 0; col < c->n_colors; col++)
    for (int cpl = 0; cpl < 2; cpl++) {
      // linear strain model: K does not depend on the state, K x = residual kernel with u := x
      p.n_list = c->list_len[2 * cpl][col];
      p.elem_list = c->list_dev[2 * cpl][col];
      p.u = x_dev; p.res = y_dev; p.alpha = 1.0;
      p.res_scale = scale * alpha; p.thermal = 0.0;
      p.jvp_x = nullptr; p.jvp_y = nullptr;
      if (p.n_list > 0 && launch_elem<true, false, false, false>(c, p, cpl != 0)) return 1;
      // nonlinear strain model: the tangent about the current state is formed per element
      // (as for assembleJacobian) and multiplied with the element's slice of x in shared
      // memory instead of being scattered
      p.n_list = c->list_len[2 * cpl + 1][col];
      p.elem_list = c->list_dev[2 * cpl + 1][col];
      p.u = c->u; p.res = nullptr; p.alpha = alpha; p.thermal = 1.0;
      p.jvp_x = x_dev; p.jvp_y = y_dev; p.jvp_scale = scale;
      if (p.n_list > 0 && launch_elem<false, true, false, true>(c, p, cpl != 0)) return 1;
    }
  CU(cudaEventRecord(c->evk1, c->stream));
  if (halo_exchange(c, y_dev, true)) return 1;
  if (c->n_bc) {
    k_vec_zero_bcs<<<(6 * c->n_bc + 255) / 256, 256, 0, c->stream>>>(c->n_bc, c->bc_nodes,
                                                                     c->bc_vars, y_dev, c->n_owned);
    c->last_launches++;
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev1, c->stream));
  return 0;
  A2DS_CATCH(a2ds_add_jacobian_vec_product_dev)
}

extern "C" int a2ds_add_jacobian_vec_product(a2ds_ctx *c, double scale, double alpha,
                                             const double *x, double *y) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  const size_t nb = 6 * (size_t)c->n_nodes * sizeof(double);
  double *dx = nullptr, *dy = nullptr;
  if (scratch_vectors(c, 6 * (size_t)c->n_nodes, 6 * (size_t)c->n_nodes, &dx, &dy)) return 1;
  CU(cudaMemcpyAsync(dx, x, nb, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(dy, y, nb, cudaMemcpyHostToDevice, c->stream));
  int rc = a2ds_add_jacobian_vec_product_dev(c, scale, alpha, dx, dy);
  if (!rc) {
    CU(cudaMemcpyAsync(y, dy, nb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return rc;
  A2DS_CATCH(a2ds_add_jacobian_vec_product)
}

extern "C" int a2ds_last_timing(a2ds_ctx *c, float *ms, int *launches) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->ev1));
  CU(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  if (ms) *ms = c->last_ms;
  if (launches) *launches = c->last_launches;
  return 0;
  A2DS_CATCH(a2ds_last_timing)
}

extern "C" int a2ds_last_kernel_ms(a2ds_ctx *c, float *ms) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->evk1));
  CU(cudaEventElapsedTime(ms, c->evk0, c->evk1));
  return 0;
  A2DS_CATCH(a2ds_last_kernel_ms)
}

extern "C" int a2ds_region_begin(a2ds_ctx *c) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->evr0, c->stream));
  return 0;
  A2DS_CATCH(a2ds_region_begin)
}

extern "C" int a2ds_region_end(a2ds_ctx *c, float *ms) {
  A2DS_TRY
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->evr1, c->stream));
  CU(cudaEventSynchronize(c->evr1));
  CU(cudaEventElapsedTime(ms, c->evr0, c->evr1));
  return 0;
  A2DS_CATCH(a2ds_region_end)
}
