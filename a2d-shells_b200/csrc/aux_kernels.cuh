// aux_kernels.cuh — the small kernels around the element kernels, included by a2ds.cu: element
// -> block offset tables, boundary conditions, vector and matrix halo pack / unpack, and the
// HBM-bound matrix algebra of the buckling flow (copy / axpy, 6x6 BCSR mat-vec).
#ifndef A2DS_AUX_KERNELS_CUH
#define A2DS_AUX_KERNELS_CUH
#ifdef __CUDACC__
#define A2DS_COLOR_HD __host__ __device__ __forceinline__
#else
#define A2DS_COLOR_HD inline
#endif

#include <cuda_runtime.h>

// One BCSR block of a matrix as the kernels see it
struct BlockDev {
  const int *rowp, *cols, *row_map, *col_map;
  long long base;  // first 6x6 block of this BCSR block in the concatenated array
  int nrows, ident;
};

// element -> block offset table; replaces TACSSchurMat::addValues' findIndex +
// BCSRMat::addRowValues' bsearch (TACSSchurMat.cpp:453-531, BCSRMat.cpp:1778-1827)
// (npe nodes per element: npe * npe slots per element, slot = npe * i + j for node pair (i, j))
// Node ids >= n_indep are dependent nodes (rows behind the independent ones): their slots are
// left at -1 here and pointed at the matrix's fold scratch afterwards (k_dep_patch_offsets).
__global__ void k_build_offsets(int n_elems, int npe, const int *conn, int n_blocks, const BlockDev *blk,
                                int *off, int *missing, int n_indep) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int n2 = npe * npe;
  if (t >= n2 * (size_t)n_elems) return;
  const size_t e = t / n2;
  const int slot = (int)(t - e * n2);
  const int rn = conn[npe * e + slot / npe], cn = conn[npe * e + slot % npe];
  if (rn >= n_indep || cn >= n_indep) { off[t] = -1; return; }
  int found = -1;
  for (int b = 0; b < n_blocks && found < 0; b++) {
    const int rr = blk[b].row_map ? blk[b].row_map[rn] : (rn < blk[b].nrows ? rn : -1);
    const int cc = blk[b].col_map ? blk[b].col_map[cn] : cn;
    if (rr < 0 || rr >= blk[b].nrows || cc < 0) continue;
    int lo = blk[b].rowp[rr], hi = blk[b].rowp[rr + 1] - 1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1, v = blk[b].cols[mid];
      if (v == cc) { found = (int)(blk[b].base + mid); break; }
      if (v < cc) lo = mid + 1; else hi = mid - 1;
    }
  }
  off[t] = found;
  if (found < 0) atomicAdd(missing, 1);
}

// ---- dependent nodes ---------------------------------------------------------------------------
// A dependent node d is the weighted sum of the independent nodes conn[ptr[d] .. ptr[d + 1])
// (TACSAssembler::setDependentNodes, src/TACSAssembler.cpp:716-775).  On the device it is one
// more row behind the n_indep local rows of X / u / res, so that the element kernels gather
// and scatter it like any node; the kernels here fill those rows before the element pass and
// distribute what the elements added to them afterwards.
//
// v[n_indep + d] = sum_j w_j v[conn_j], accumulated from zero in list order
// (TACSBVec::endDistributeValues, src/bpmat/TACSBVec.cpp:930-975)
__global__ void k_dep_gather(int n_dep, int nc, int n_indep, const int *ptr, const int *conn,
                             const double *w, double *v) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_dep * nc) return;
  const int d = t / nc, k = t - d * nc;
  double z = 0.0;
  for (int j = ptr[d]; j < ptr[d + 1]; j++) z += w[j] * v[(size_t)nc * conn[j] + k];
  v[(size_t)nc * (n_indep + d) + k] = z;
}
// v[conn_j] += w_j v[n_indep + d], then the dependent row is cleared
// (TACSBVec::beginSetValues with TACS_ADD_VALUES / endSetValues, TACSBVec.cpp:855-910)
__global__ void k_dep_scatter(int n_dep, int nc, int n_indep, const int *ptr, const int *conn,
                              const double *w, double *v) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_dep * nc) return;
  const int d = t / nc, k = t - d * nc;
  const double z = v[(size_t)nc * (n_indep + d) + k];
  v[(size_t)nc * (n_indep + d) + k] = 0.0;
  for (int j = ptr[d]; j < ptr[d + 1]; j++) atomicAdd(&v[(size_t)nc * conn[j] + k], w[j] * z);
}
// element -> block offsets of the node pairs with a dependent node: fold slot s lives in block
// base + s of the value array (behind the matrix proper)
__global__ void k_dep_patch_offsets(int n_slots, const int *slot_pos, int base, int *off) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_slots) off[slot_pos[s]] = base + s;
}
// W^T K_e W (TACSAssembler::addMatValues -> addWeightValues, src/TACSAssembler.h:485-510): the
// element block a kernel added to fold slot s goes to every (independent row, independent
// column) pair behind it with the product of the two weights; the slot is cleared for the next
// assembly.  The target list is built on the host when the matrix is created.
__global__ void k_dep_fold(int n_slots, long long base, const int *fptr, const int *ftgt,
                           const double *fw, double *A) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= 36 * (size_t)n_slots) return;
  const int s = (int)(t / 36), k = (int)(t - 36 * (size_t)s);
  double *src = A + 36 * (size_t)(base + s) + k;
  const double v = *src;
  *src = 0.0;
  if (v == 0.0) return;
  for (int j = fptr[s]; j < fptr[s + 1]; j++) atomicAdd(&A[36 * (size_t)ftgt[j] + k], fw[j] * v);
}

// ---- element colouring on the device --------------------------------------------------------
// Elements sharing a node get different colours (A2DS_SCATTER_COLORED: one launch per colour,
// every block / residual row receives at most one contribution per launch).  Rule: greedy in the
// order of a hashed priority (color_key, ties impossible: the element index is in its low word) —
// an element takes the smallest colour none of its higher-priority neighbours holds.  On the
// device that is a Jones-Plassmann sweep: per round every uncoloured element whose uncoloured
// neighbours all have a lower key colours itself; two adjacent elements are never coloured in
// the same round, and the result does not depend on the timing — it is the colouring the
// sequential greedy pass in decreasing key order produces (color_elements_hashed on the host,
// bit-identical).  O(log n) rounds instead of the n-long dependency chain of natural order.
// Runs of four consecutive elements share one hash and are ordered by index inside the run: on
// meshes numbered along rows the greedy pass then sees the regular neighbourhood natural order
// sees, and needs 6 colours on a structured mesh (4 large, 2 small) where per-element hashes
// need 8 or 9 — for about 50 rounds instead of 20.
A2DS_COLOR_HD unsigned long long color_key(int e) {
  unsigned h = ((unsigned)e >> 2) + 0x9e3779b9u;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return ((unsigned long long)h << 32) | (0xffffffffu - (unsigned)e);
}
A2DS_COLOR_HD int color_key_elem(unsigned long long key) { return (int)(0xffffffffu - (unsigned)(key & 0xffffffffu)); }
#ifdef __CUDACC__
__global__ void k_color_round(int ne, const int *conn, const int *ptr, const int *adj, int *color,
                              int *left) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne || __ldcg(&color[e]) >= 0) return;
  const unsigned long long key = color_key(e);
  int nodes[4];
#pragma unroll
  for (int i = 0; i < 4; i++) nodes[i] = conn[4 * (size_t)e + i];
#pragma unroll
  for (int i = 0; i < 4; i++)
    for (int k = ptr[nodes[i]]; k < ptr[nodes[i] + 1]; k++) {
      const int o = adj[k];
      if (o != e && __ldcg(&color[o]) < 0 && color_key(o) > key) { atomicAdd(left, 1); return; }
    }
  // every neighbour still uncoloured has a lower key and waits for this element: the colours
  // read below are final
  for (int base = 0;; base += 64) {
    unsigned long long used = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
      for (int k = ptr[nodes[i]]; k < ptr[nodes[i] + 1]; k++) {
        const int co = __ldcg(&color[adj[k]]) - base;
        if (co >= 0 && co < 64) used |= 1ull << co;
      }
    if (~used) { color[e] = base + __ffsll((long long)~used) - 1; return; }
  }
}
#endif

// ---- non-zero pattern of the natural-order matrix on the device -------------------------------
// TACSAssembler::computeLocalNodeToNodeCSR + TacsSortAndUniquifyCSR (src/TACSAssembler.cpp:1839,
// src/utils/TacsUtilities.cpp:280): node -> nodes of its elements, sorted and unique per row.
// Three light kernels around two prefix sums: element incidences per node, then per node the
// (at most 4 per element) candidate columns are sorted and made unique in a per-thread array
// — once to count, once to write.  PAT_MAX_ELEMS bounds the elements around one node; meshes
// beyond it (very high-valence fans) take the host path.
static const int PAT_MAX_ELEMS = 24;
__global__ void k_pat_count(size_t n4, const int *conn, int *deg) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t < n4) atomicAdd(&deg[conn[t]], 1);
}
__global__ void k_pat_fill(size_t n4, const int *conn, const int *ptr, int *cursor, int *adj) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n4) return;
  const int n = conn[t];
  adj[ptr[n] + atomicAdd(&cursor[n], 1)] = (int)(t >> 2);
}
// WRITE = false: cnt[n] = number of distinct columns of row n (overflow: more than PAT_MAX_ELEMS
// elements around a node); WRITE = true: cols[rowp[n] ..] = the sorted distinct columns
template <bool WRITE>
__global__ void k_pat_rows(int nn, const int *conn, const int *ptr, const int *adj, int *cnt,
                           const int *rowp, int *cols, int *overflow) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  const int a = ptr[n], b = ptr[n + 1];
  if (b - a > PAT_MAX_ELEMS) {
    if (!WRITE) { cnt[n] = 0; atomicAdd(overflow, 1); }
    return;
  }
  int c[4 * PAT_MAX_ELEMS];
  int m = 0;
  for (int k = a; k < b; k++) {
    const int e = adj[k];
    for (int j = 0; j < 4; j++) {
      const int v = conn[4 * (size_t)e + j];
      // insertion into the sorted, distinct prefix c[0 .. m)
      int pos = m;
      while (pos > 0 && c[pos - 1] > v) pos--;
      if (pos > 0 && c[pos - 1] == v) continue;
      for (int i = m; i > pos; i--) c[i] = c[i - 1];
      c[pos] = v;
      m++;
    }
  }
  if (!WRITE) {
    cnt[n] = m;
  } else {
    int *out = cols + rowp[n];
    for (int i = 0; i < m; i++) out[i] = c[i];
  }
}

// residual BC rows: r = u - ubar on owned nodes (TACSBVec::applyBCs, TACSBVec.cpp:546-585)
// (rows [row_lo, row_hi) only: the streamed assembly finishes the residual chunk by chunk)
__global__ void k_res_bcs(int n_bc, const int *nodes, const int *vars, const double *vals,
                          const double *u, double *res, int row_lo, int row_hi) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * n_bc) return;
  const int b = t / 6, k = t - 6 * b, n = nodes[b];
  if (n >= row_lo && n < row_hi && (vars[b] & (1 << k))) res[6 * (size_t)n + k] = u[6 * (size_t)n + k] - vals[t];
}

// vector BC rows set to zero (TACSBVec::applyBCs without a state vector, TACSBVec.cpp:570-584)
__global__ void k_vec_zero_bcs(int n_bc, const int *nodes, const int *vars, double *y, int n_owned) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * n_bc) return;
  const int b = t / 6, k = t - 6 * b, n = nodes[b];
  if (n < n_owned && (vars[b] & (1 << k))) y[6 * (size_t)n + k] = 0.0;
}

// y[bc] = x[bc] on the owned BC rows: what the identity rows of the assembled matrix give
__global__ void k_vec_copy_bcs(int n_bc, const int *nodes, const int *vars, const double *x,
                               double *y, int n_owned) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * n_bc) return;
  const int b = t / 6, k = t - 6 * b, n = nodes[b];
  if (n < n_owned && (vars[b] & (1 << k))) y[6 * (size_t)n + k] = x[6 * (size_t)n + k];
}

// matrix BC rows: zero the constrained DOF rows of every block in the block row, 1.0 on
// the diagonal entry of the diagonal block (BCSRMat::zeroRow, BCSRMat.cpp:2005-2030;
// columns untouched, as the reference)
// One warp per (BC node, BCSR block): the lanes sweep the 36 entries of every block of the row
// (coalesced); blockIdx.y selects the matrix, so the tangent and the geometric stiffness of a
// fused assembly share one launch.
struct BcMats {
  const BlockDev *blk[3];
  double *A[3];
  int n_blocks[3];
};
__global__ void k_mat_bcs(int n_bc, const int *nodes, const int *vars, BcMats M) {
  const int im = blockIdx.y;
  const int n_blocks = M.n_blocks[im];
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n_bc * n_blocks) return;
  const int b = t / n_blocks, ib = t - b * n_blocks;
  const BlockDev B = M.blk[im][ib];
  double *A = M.A[im];
  const int n = nodes[b], mask = vars[b];
  const int row = B.row_map ? B.row_map[n] : n;
  if (row < 0 || row >= B.nrows) return;
  const int diag_col = B.col_map ? B.col_map[n] : n;
  for (int j = B.rowp[row]; j < B.rowp[row + 1]; j++) {
    double *a = &A[36 * (size_t)(B.base + j)];
    const bool diag = B.ident && B.cols[j] == diag_col;
    for (int e = lane; e < 36; e += 32) {
      const int ii = e / 6, jj = e - 6 * ii;
      if (mask & (1 << ii)) a[e] = (diag && ii == jj) ? 1.0 : 0.0;
    }
  }
}

// halo pack / unpack (VecDistGetVars6 and the TACS_ADD_VALUES scatter of
// TACSBVecDistribute.cpp:543-747)
__global__ void k_pack6(int n, const int *nodes, const double *v, double *buf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 6 * n) buf[t] = v[6 * (size_t)nodes[t / 6] + t % 6];
}
__global__ void k_unpack6(int n, const int *nodes, const double *buf, double *v, int add) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * n) return;
  double *d = &v[6 * (size_t)nodes[t / 6] + t % 6];
  if (add) *d += buf[t]; else *d = buf[t];
}

// y <- beta y + alpha x over all stored values (BCSRMat::axpy / copyValues, BCSRMat.cpp:2375,
// 2430).  Pure HBM streaming: 128-bit accesses, four independent loads in flight per thread.
// n2 = number of double2 (block values come in multiples of 36 doubles, 16-byte aligned).
template <bool COPY>
__global__ void __launch_bounds__(256) k_axpy(size_t n2, double alpha, const double2 *__restrict__ x,
                                              double2 *__restrict__ y) {
  // one contiguous 16 KB tile per block (4 x 128-bit per thread), no grid-stride loop: the
  // DRAM pages of a tile are touched by one block at one time
  const size_t base = blockIdx.x * (size_t)(4 * 256) + threadIdx.x;
  double2 xv[4], yv[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const size_t i = base + 256 * k;
    if (i < n2) {
      xv[k] = x[i];
      if (!COPY) yv[k] = y[i];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const size_t i = base + 256 * k;
    if (i < n2) {
      double2 r;
      if (COPY) r = xv[k];
      else { r.x = yv[k].x + alpha * xv[k].x; r.y = yv[k].y + alpha * xv[k].y; }
      y[i] = r;
    }
  }
}

// 6x6 BCSR mat-vec (BCSRMatVecMult6, BCSRMatMult6.cpp:82): one thread per scalar row, the six
// threads of a block row read the 288 contiguous bytes of each block between them.
// HBM bound: 288 B of values + 4 B of column index per block; x is reused through L1/L2.
__global__ void k_spmv6(int nrows, const int *__restrict__ rowp, const int *__restrict__ cols,
                        const double *__restrict__ A, const double *__restrict__ x,
                        double *__restrict__ y) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int row = (int)(t / 6), r = (int)(t - 6 * (size_t)row);
  if (row >= nrows) return;
  double acc = 0.0;
  const int end = rowp[row + 1];
  for (int k = rowp[row]; k < end; k++) {
    const double2 *a = reinterpret_cast<const double2 *>(A + 36 * (size_t)k + 6 * r);
    const double2 *xv = reinterpret_cast<const double2 *>(x + 6 * (size_t)__ldg(&cols[k]));
    const double2 a0 = __ldg(a), a1 = __ldg(a + 1), a2 = __ldg(a + 2);
    const double2 x0 = __ldg(xv), x1 = __ldg(xv + 1), x2 = __ldg(xv + 2);
    acc += a0.x * x0.x + a0.y * x0.y + a1.x * x1.x + a1.y * x1.y + a2.x * x2.x + a2.y * x2.y;
  }
  y[t] = acc;
}

// ---- matrix halo (TACSParallelMat flavour): ghost rows -> owners, added --------------------
// TACSMatDistribute::beginAssembly/endAssembly (src/bpmat/TACSMatDistribute.cpp:1036-1176):
// contributions a rank made to block rows it does not own travel to the owner and are added
// there.  The plan (which of my blocks go to which peer, and where arriving blocks land) is
// supplied by the host, which knows the global numbering (a2ds_mat_set_halo).
__global__ void k_pack36(int n, const int *blk, const double2 *A, double2 *buf) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= 18 * (size_t)n) return;
  const size_t b = t / 18, e = t - 18 * b;
  buf[t] = A[18 * (size_t)blk[b] + e];
}
__global__ void k_unpack36_add(int n, const int *blk, const double2 *buf, double2 *A) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= 18 * (size_t)n) return;
  const size_t b = t / 18, e = t - 18 * b;
  double2 *d = &A[18 * (size_t)blk[b] + e];
  const double2 v = buf[t];
  d->x += v.x; d->y += v.y;
}

#endif
