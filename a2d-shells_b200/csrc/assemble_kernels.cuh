// assemble_kernels.cuh — the element kernels of liba2ds_b200.so (sm_100a), included by a2ds.cu.
//
// k_assemble<RES, KMAT, GMAT, NL>: a warp draws batches of NB MITC4 elements from a work
// counter; per batch the node and Gauss-point phases run with one lane per (element, node) /
// (element, point), then element by element all 32 lanes form their columns of the strain
// matrices in registers (mitc4_math.h) — which ARE the fragments of the 24x24 contractions on
// the FP64 tensor path (mma.sync m8n8k4 f64) — stage the element matrices in shared memory and
// scatter them with full-warp RED.E.ADD.F64 through the element -> block offset tables.
// k_mass<RES, MAT>: mass matrix / inertial residual with the same batching and scatter.
// Roofline notes live in DESIGN.md, measurements in profiles/.
#ifndef A2DS_ASSEMBLE_KERNELS_CUH
#define A2DS_ASSEMBLE_KERNELS_CUH

#include <cuda_runtime.h>

#include "mitc4_math.h"
#include "mitc4_tying.h"

using namespace a2ds;
#ifndef A2DS_ZWAIT
#define A2DS_ZWAIT 2   // 1: zeroing of the output matrices inside k_assemble_t with the round protocol of
#endif                 // ZeroPlan.  Built and measured, not faster than the memsets it replaces
                       // (profiles/README.md).  2 (default): double buffering — the kernel zeroes the
                       // SPARE value array of each matrix (the one the next assembly will add into) with
                       // bulk copies spread over the lanes; nothing in the launch depends on them, so
                       // there is no fence, no counter and no wait: rounds are only the way the work is
                       // spread over the trips.  0: memsets in front of the kernel.
// work description of the zeroing a launch of an element kernel does on the side (see "zeroing
// of the output matrices" below)
struct ZeroPlan {
  double2 *zK, *zG;        // arrays to zero (null: none)
  long long nK, nG;        // lengths in double2 units
  int cK, cG;              // chunk of one warp in one round (double2 units, whole blocks)
  float inv_round;         // 1 / (blocks per round)
  int rounds, ahead;       // number of rounds; rounds zeroed before the first batch
  int *done;               // per round: number of warps that have zeroed it
};

struct KParams {
  const int *elem_list;  // elements to process (NULL: 0..n_list-1)
  int n_list;
  const int *conn;       // 4 local nodes per element
  const int *elem_comp;
  const CompData *comps;
  const double *X;       // 3 per node
  const double *u;       // 6 per node
  double *res;           // 6 per node
  double *Kval;          // concatenated block values of the tangent matrix
  const int *Koff;       // 16 block offsets per element (-1: not stored here)
  double *Gval;
  const int *Goff;
  double alpha;          // scale of the tangent (and of the mass matrix in k_mass)
  double gscale;         // scale of the geometric stiffness (assembleMatCombo; 1 otherwise)
  double res_scale;      // residual entries are multiplied by this before the RED
  double thermal;        // 1: residual of the state;  0: matrix-free product K x (u := x)
  // matrix-free tangent product of the nonlinear model: y += jvp_scale * K_e(u) x_e instead
  // of the matrix scatter (null: scatter)
  const double *jvp_x;
  double *jvp_y;
  double jvp_scale;
  int scratch_bytes;
  int *work_counter;     // dynamic batch scheduling: next batch index (zeroed per launch)
  // zeroing of the output matrices inside k_assemble_t (cooperative launch; null: the caller
  // zeroed them): see ZeroPlan
  const struct ZeroPlan *zplan;
  // A2DS_ZWAIT == 2 (double buffering): the plan by value, zval.rounds == 0: nothing to zero
  ZeroPlan zval;
};
#if A2DS_ZWAIT == 2
#define A2DS_ZP(p) ((p).zval)
#else
#define A2DS_ZP(p) (*(p).zplan)
#endif

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

static const int MAX_WARPS_PER_BLOCK = 4;
static const int MAX_DEVICES = 64;  // per-device caches of launch configurations
// Register budget: the residual / tangent kernels run at 12 warps per SM (168
// registers); the variants that carry the B1(q) fragments across the tangent pass
// (geometric stiffness, nonlinear model) need the full 255 registers (8 warps per SM).
#ifndef A2DS_MB_G
#define A2DS_MB_G 2
#endif
#ifndef A2DS_MB_K
#define A2DS_MB_K 3
#endif
#define A2DS_MIN_BLOCKS(GMAT, NL) (((GMAT) || (NL)) ? A2DS_MB_G : A2DS_MB_K)

// scatter one staged 24x24 element matrix.  Per 8 of the 16 node-pair blocks: 8 full-warp
// REDs carry entries 0..31 of one block each (consecutive lanes -> consecutive doubles) and
// ONE more full-warp RED carries the four remaining entries 32..35 of all 8 blocks (lane =
// 4 * block + entry): 18 REDs per matrix, no divergent tail.  Every slot has a block
// (a2ds_mat_create refuses patterns with missing blocks), so there is no validity branch;
// loads are batched so the REDs do not wait on shared memory one by one.
template <bool SCALED = false>
__device__ __forceinline__ void scatter_matrix(const double *E, double *vals, int off16,
                                               int lane, double scale = 1.0) {
  const unsigned FULL = 0xffffffffu;
  const int r0 = lane / 6, c0 = lane - 6 * r0;   // entry `lane` of a 6x6 block
  const int src0 = r0 * KE_LD + c0;
  const int tb = lane >> 2;                      // tail RED: block tb of the current 8
  const int src1 = 6 * (tb >> 2) * KE_LD + 6 * (tb & 3) + 5 * KE_LD + 2 + (lane & 3);
#pragma unroll
  for (int half = 0; half < 2; half++) {
    double v0[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int b = 8 * half + k;
      v0[k] = E[6 * (b >> 2) * KE_LD + 6 * (b & 3) + src0];
      if (SCALED) v0[k] *= scale;
    }
    double v1 = E[12 * half * KE_LD + src1];   // blocks 8..15 start two block rows down
    if (SCALED) v1 *= scale;
    const int offt = __shfl_sync(FULL, off16, 8 * half + tb);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int off = __shfl_sync(FULL, off16, 8 * half + k);
      atomicAdd(vals + 36 * (size_t)off + lane, v0[k]);
    }
    atomicAdd(vals + 36 * (size_t)offt + 32 + (lane & 3), v1);
  }
}

// DMMA accumulators of the 6 upper tiles -> staged symmetric 24x24.  In the fragment
// layout of mitc4_math.h the accumulator of tile (ti, tj) holds, on lane (c, qp'),
// K[3 c + ti][6 qp' + tj] and K[3 c + ti][6 qp' + 3 + tj]  (c = lane >> 2, qp' = lane & 3).
__device__ __forceinline__ void stage_tiles(double *E, const double (&acc)[6][2], double scale,
                                            int lane) {
  const int rowb = 3 * (lane >> 2), colb = 6 * (lane & 3);
  int idx = 0;
#pragma unroll
  for (int ti = 0; ti < 3; ti++)
#pragma unroll
    for (int tj = ti; tj < 3; tj++, idx++) {
      const int row = rowb + ti, col0 = colb + tj, col1 = colb + 3 + tj;
      const double a0 = scale * acc[idx][0], a1 = scale * acc[idx][1];
      E[row * KE_LD + col0] = a0;
      E[row * KE_LD + col1] = a1;
      if (ti != tj) {
        E[col0 * KE_LD + row] = a0;
        E[col1 * KE_LD + row] = a1;
      }
    }
}

// all 9 tiles of an unsymmetric 24x24 (Z = B1^T W) into the staging area
__device__ __forceinline__ void stage_tiles_full(double *E, const double (&acc)[9][2], int lane) {
  const int rowb = 3 * (lane >> 2), colb = 6 * (lane & 3);
#pragma unroll
  for (int ti = 0; ti < 3; ti++)
#pragma unroll
    for (int tj = 0; tj < 3; tj++) {
      E[(rowb + ti) * KE_LD + colb + tj] = acc[3 * ti + tj][0];
      E[(rowb + ti) * KE_LD + colb + 3 + tj] = acc[3 * ti + tj][1];
    }
}

// G = Z + Z^T + geometric blocks, written block by block (64 generalised node pairs, 2 per
// lane) from the staged Z into a second buffer
__device__ __forceinline__ void symmetrize_add_geo(const ElemGeom &gm, const ElemWork &wk,
                                                   const double *Pq4, const double *Z, double *G,
                                                   double scale, int lane) {
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const int pair = lane + 32 * pass, pr = pair >> 3, pc = pair & 7;
    double blk[9];
    geo_block(gm, wk, Pq4, pr, pc, blk);
    const int r0 = 6 * (pr & 3) + (pr >= 4 ? 3 : 0), c0 = 6 * (pc & 3) + (pc >= 4 ? 3 : 0);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        G[(r0 + i) * KE_LD + c0 + j] =
            scale * (blk[3 * i + j] + Z[(r0 + i) * KE_LD + c0 + j] + Z[(c0 + j) * KE_LD + r0 + i]);
  }
}

// 64 geometric-stiffness 3x3 blocks (generalised node pairs), 2 per lane, added in place
__device__ __forceinline__ void add_geo_blocks(const ElemGeom &gm, const ElemWork &wk,
                                               const double *Pq4, double *E, double scale,
                                               int lane) {
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const int pair = lane + 32 * pass, pr = pair >> 3, pc = pair & 7;
    double blk[9];
    geo_block(gm, wk, Pq4, pr, pc, blk);
    const int r0 = 6 * (pr & 3) + (pr >= 4 ? 3 : 0), c0 = 6 * (pc & 3) + (pc >= 4 ? 3 : 0);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) E[(r0 + i) * KE_LD + c0 + j] += scale * blk[3 * i + j];
  }
}

// A warp prepares NB elements at a time: the node phase and the Gauss-point phase then run
// on NB*4 distinct work items (one lane each) instead of 8 redundant copies per element.
// Build-time knobs for occupancy experiments (defaults = the measured configuration):
//   A2DS_NB          elements per batch (2, 4 or 8).  Shared memory per warp scales with it:
//                    27.7 KB at 4, 19.4 KB at 2 for the geometric-stiffness / nonlinear variants
//   A2DS_PREFETCH_G  0: no double-buffered gather for those variants (-0.7 KB per warp at NB 2)
//   A2DS_MB_G        blocks per SM the launch bounds ask for (3 -> 168 registers, 12 warps/SM)
// e.g. -DA2DS_NB=2 -DA2DS_PREFETCH_G=0 -DA2DS_MB_G=3: 18.7 KB per warp, 12 instead of 8 warps
// per SM for the fused kernel (192 B of spills per thread) — see profiles/README.md.
#ifndef A2DS_NB
#define A2DS_NB 4
#endif
#ifndef A2DS_PREFETCH_G
#define A2DS_PREFETCH_G 1
#endif
#if A2DS_PREFETCH_G
#define A2DS_RAW1(ws) (ws).raw1
#define A2DS_GOFF1 1
#else   // never selected at run time (PF is false), only has to name something that exists
#define A2DS_RAW1(ws) (ws).raw0
#define A2DS_GOFF1 0
#endif
static const int NB = A2DS_NB;
static_assert(NB == 2 || NB == 4 || NB == 8,
              "one lane per (element, node): NB * 4 <= 32; the offset gather moves 32 entries per pass");
struct alignas(16) RawBatch {   // gathered inputs of one batch, filled by cp.async
  double xq[NB][36];       // per element: X[12] then q[24]
  int koff[NB][16];
  int comp[NB];
};
struct WarpScratch {
  RawBatch raw0;
  int nodes[NB][4];
  ElemGeom geo[NB];
  double E[24 * KE_LD];   // staging of a 24x24 element matrix for the scatter (tangent)
  // ---- only the geometric-stiffness / nonlinear variants use what follows; the linear
  //      residual / tangent kernels allocate up to here (12 instead of 8 warps per SM) ----
  double E2[24 * KE_LD];  // second staging buffer (geometric stiffness)
  ElemWork work;
  double Pq[NB][4][6];    // per Gauss point T T^T
#if A2DS_PREFETCH_G
  RawBatch raw1;          // double buffer: batch i+1 lands (cp.async) while batch i is processed
  int goff[2][NB][16];    // block offsets of the geometric stiffness matrix (per raw buffer)
#else
  int goff[1][NB][16];    // block offsets of the geometric stiffness matrix
#endif
};

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

#if A2DS_ZWAIT == 2
__device__ __forceinline__ void zero_round(const ZeroPlan &z, int r, int gw, int n_gw, unsigned src, int lane);
#endif

template <bool RES, bool KMAT, bool GMAT, bool NL>
__global__ void __launch_bounds__(MAX_WARPS_PER_BLOCK * 32, A2DS_MIN_BLOCKS(GMAT, NL))
    k_assemble(const KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#if A2DS_ZWAIT == 2
  // double buffering (see k_assemble_t): the spare value arrays of the matrices are zeroed on
  // the side, one round per trip, from a zeroed 1 KB buffer
  __shared__ alignas(16) double zero_src1[128];
  const bool zon = (KMAT || GMAT) && p.zval.rounds > 0;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp, n_gw = gridDim.x * (blockDim.x >> 5);
  int zr = 0;
  unsigned zsrc = 0;
  if (KMAT || GMAT) {
    if (threadIdx.x < 128) zero_src1[threadIdx.x] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    zsrc = (unsigned)__cvta_generic_to_shared(zero_src1);
    __syncthreads();
  }
#endif
  WarpScratch &ws = *reinterpret_cast<WarpScratch *>(smem_raw + (size_t)warp * p.scratch_bytes);
  ElemWork &wk = ws.work;
  // full scratch + asynchronous prefetch of the next batch
  const bool PF = (GMAT || NL) && A2DS_PREFETCH_G != 0;
  const unsigned FULL = 0xffffffffu;
  Want w;
  w.res = RES; w.kmat = KMAT; w.gmat = GMAT; w.nonlinear = NL; w.thermal = p.thermal;
  const bool need_state = GMAT || NL;

  const int n_groups = (p.n_list + NB - 1) / NB;

  // element id and node id this lane is responsible for in a batch: lane = 4 j + m
  auto batch_ids = [&](int grp_, int &e_out, int &nd_out) {
    const int j = (lane >> 2) & (NB - 1);
    const int idx = grp_ * NB + j;
    e_out = -1; nd_out = 0;
    if (grp_ < n_groups && idx < p.n_list) {
      e_out = p.elem_list ? __ldg(&p.elem_list[idx]) : idx;
      nd_out = __ldg(&p.conn[4 * e_out + (lane & 3)]);
    }
  };
  // asynchronous gather of a batch into raw buffer `rb` (addresses from the ids above)
  auto issue_gather = [&](RawBatch &rb, int (*goffb)[16], int e_l, int nd_l) {
#pragma unroll
    for (int r = 0; r < (NB * 36 + 31) / 32; r++) {
      const int sidx = lane + 32 * r;
      const int j = sidx / 36, k = sidx - 36 * j;
      const bool isx = k < 12;
      const int node = isx ? k / 3 : (k - 12) / 6;
      const int comp_k = isx ? k - 3 * node : (k - 12) - 6 * node;
      const int src_lane = (4 * j + node) & 31;
      const int nd = __shfl_sync(FULL, nd_l, src_lane);
      const int ej = __shfl_sync(FULL, e_l, (4 * j) & 31);
      if (sidx < NB * 36 && ej >= 0) {
        const double *src = isx ? &p.X[3 * (size_t)nd + comp_k] : &p.u[6 * (size_t)nd + comp_k];
        cp_async8(&rb.xq[j][k], src);
      }
    }
#pragma unroll
    for (int r = 0; r < NB * 16 / 32; r++) {
      const int sidx = lane + 32 * r;
      const int j = sidx >> 4, k = sidx & 15;
      const int ej = __shfl_sync(FULL, e_l, (4 * j) & 31);
      if (ej >= 0) {
        if (KMAT && p.Koff) cp_async4(&rb.koff[j][k], &p.Koff[16 * (size_t)ej + k]);
        if (GMAT) cp_async4(&goffb[j][k], &p.Goff[16 * (size_t)ej + k]);
      }
    }
    if ((lane & 3) == 0 && lane < 4 * NB && e_l >= 0) cp_async4(&rb.comp[lane >> 2], &p.elem_comp[e_l]);
  };

  // dynamic scheduling: warps draw the next batch from one counter, so that the batches in
  // flight at any time stay close together in the element order (L2 locality of the
  // read-modify-write scatter) however the warps drift
  auto next_group = [&]() {
    int g = 0;
    if (lane == 0) g = atomicAdd(p.work_counter, 1);
    return __shfl_sync(FULL, g, 0);
  };
  int grp = next_group();
  int grp_nxt = PF ? next_group() : 0;
  int buf = 0;
  int e_cur, nd_cur;
  batch_ids(grp, e_cur, nd_cur);
  if (PF && grp < n_groups) issue_gather(ws.raw0, ws.goff[0], e_cur, nd_cur);
  for (; grp < n_groups; buf ^= 1) {
    // the draw for the trip after this one (after next with prefetch) is issued now and read
    // at the end of the trip: its latency never shows
    int drawn = 0;
    if (lane == 0) drawn = atomicAdd(p.work_counter, 1);
    const int base = grp * NB;
    const int cnt = min(NB, p.n_list - base);
#if A2DS_ZWAIT == 2
    if (zon && zr < p.zval.rounds) { zero_round(p.zval, zr, gw, n_gw, zsrc, lane); zr++; }
#endif
    // ids of the NEXT batch: requested now, consumed after the geometry phases
    int e_nxt = -1, nd_nxt = 0;
    if (PF) batch_ids(grp_nxt, e_nxt, nd_nxt);

    // ---- the batch gathered during the previous trip (or right now without prefetch) -----
    if (!PF) issue_gather(ws.raw0, ws.goff[0], e_cur, nd_cur);
    cp_async_wait_all();
    __syncwarp();
    const RawBatch &rb = (PF && buf) ? A2DS_RAW1(ws) : ws.raw0;
    const int (*goffb)[16] = ws.goff[(PF && buf) ? A2DS_GOFF1 : 0];
    if (lane < 4 * NB) ws.nodes[lane >> 2][lane & 3] = nd_cur;
#pragma unroll
    for (int r = 0; r < (NB * 36 + 31) / 32; r++) {
      const int sidx = lane + 32 * r;
      const int j = sidx / 36, k = sidx - 36 * j;
      if (sidx < NB * 36) {
        const double v = rb.xq[j][k];
        if (k < 12) ws.geo[j].X[k] = v; else ws.geo[j].q[k - 12] = v;
      }
    }
    __syncwarp();

    // ---- node phase: lane = (element of the batch, node) --------------------------------
    {
      const int j = (lane >> 2) & (NB - 1);
      if (lane < 4 * NB && j < cnt) phase_node(p.comps[rb.comp[j]], ws.geo[j], lane & 3);
    }
    __syncwarp();
    // ---- Gauss point phase: lane = (element of the batch, Gauss point) ------------------
    {
      const int j = (lane >> 2) & (NB - 1);
      if (lane < 4 * NB && j < cnt)
        phase_qp(p.comps[rb.comp[j]], ws.geo[j], lane & 3, RES || GMAT || NL, need_state, NL,
                 need_state ? &ws.Pq[j][lane & 3][0] : (double *)0);
    }
    // start the gather of the next batch into the other raw buffer
    if (PF && grp_nxt < n_groups)
      issue_gather(buf ? ws.raw0 : A2DS_RAW1(ws), ws.goff[buf ? 0 : A2DS_GOFF1], e_nxt, nd_nxt);
    if (PF) { e_cur = e_nxt; nd_cur = nd_nxt; }
    __syncwarp();

#pragma unroll 1
    for (int j = 0; j < cnt; j++) {
      const ElemGeom &gm = ws.geo[j];
      const CompData &c = p.comps[rb.comp[j]];

      // ---- column phase + contractions on the FP64 tensor path ----------------------
      // The lane's columns of B, w C B and B1(q) ARE the DMMA fragments (mitc4_math.h), so
      // operands go from the FMA pipe to the tensor path without leaving registers.
      //      K = B^T (w C B),   G = B1^T W + W^T B1      (upper tiles only)
      // Order: B0/W -> strains, residual -> K pass -> B1 -> G pass, so that B0 and B1 are
      // never live together (the nonlinear model needs B1 first: B = B0 + B1).
      double Bc[9][3], Wc[9][3], Bq[9][3];
      double kacc[6][2], zacc[9][2];
      if (NL) lane_b1(gm, wk, lane, Bq);
      lane_b0w(c, gm, lane, w, Bq, Bc, Wc);
      if (RES || GMAT || NL) {
        double r3[3];
        lane_stress(c, gm, wk, lane, w, Wc, r3);
        if (RES) {
          // residual: sum over the 4 Gauss points (lane bits 0..1)
#pragma unroll
          for (int k = 0; k < 3; k++) {
            r3[k] += __shfl_xor_sync(FULL, r3[k], 1);
            r3[k] += __shfl_xor_sync(FULL, r3[k], 2);
          }
          if ((lane & 3) == 0) {
            double *r = &p.res[6 * (size_t)ws.nodes[j][lane_m(lane)] + 3 * lane_h(lane)];
            atomicAdd(r, p.res_scale * r3[0]);
            atomicAdd(r + 1, p.res_scale * r3[1]);
            atomicAdd(r + 2, p.res_scale * r3[2]);
          }
        }
      }
      if (KMAT) {
#pragma unroll
        for (int t = 0; t < 6; t++) kacc[t][0] = kacc[t][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 9; ks++) {
          int idx = 0;
#pragma unroll
          for (int ti = 0; ti < 3; ti++)
#pragma unroll
            for (int tj = ti; tj < 3; tj++, idx++) dmma884(kacc[idx], Bc[ks][ti], Wc[ks][tj]);
        }
      }
      if (GMAT) {
        // Z = B1^T W over all 9 tiles (not symmetric); G = Z + Z^T is formed when the tiles
        // are staged: 9 instead of 12 DMMAs per k-step
        lane_b1(gm, wk, lane, Bq);
#pragma unroll
        for (int t = 0; t < 9; t++) zacc[t][0] = zacc[t][1] = 0.0;
        // the drilling strain is linear in the state: row 8 of B1 is zero, 8 k-steps suffice
#pragma unroll
        for (int ks = 0; ks < 8; ks++)
#pragma unroll
          for (int ti = 0; ti < 3; ti++)
#pragma unroll
            for (int tj = 0; tj < 3; tj++) dmma884(zacc[3 * ti + tj], Bq[ks][ti], Wc[ks][tj]);
      }
      // ---- stage, add the geometric blocks, scatter -------------------------------------
      if (KMAT) stage_tiles(ws.E, kacc, p.alpha, lane);
      if (GMAT) stage_tiles_full(ws.E2, zacc, lane);
      if (GMAT || NL) {
        __syncwarp();  // the per Gauss point stresses (lane_stress) are published
        if (lane < 9) sum_tying_stress(wk, lane);
      }
      __syncwarp();
      if (NL && KMAT) {
        add_geo_blocks(gm, wk, &ws.Pq[j][0][0], ws.E, p.alpha, lane);
        __syncwarp();
      }
      if (KMAT) {
        if (NL && !GMAT && p.jvp_x) {
          // matrix-free: the staged tangent times the element's slice of x (the second
          // staging buffer is free in this variant)
          const int node = ws.nodes[j][(lane / 6) & 3];
          if (lane < 24) ws.E2[lane] = p.jvp_x[6 * (size_t)node + lane % 6];
          __syncwarp();
          if (lane < 24) {
            double y = 0.0;
#pragma unroll
            for (int k = 0; k < 24; k++) y += ws.E[lane * KE_LD + k] * ws.E2[k];
            atomicAdd(&p.jvp_y[6 * (size_t)node + lane % 6], p.jvp_scale * y);
          }
        } else {
          scatter_matrix(ws.E, p.Kval, rb.koff[j][lane & 15], lane);
        }
      }
      if (GMAT) {
        if (KMAT) __syncwarp();  // E is reused for G once K has left
        symmetrize_add_geo(gm, wk, &ws.Pq[j][0][0], ws.E2, ws.E, p.gscale, lane);
        __syncwarp();
        scatter_matrix(ws.E, p.Gval, goffb[j][lane & 15], lane);
      }
      __syncwarp();
    }
    drawn = __shfl_sync(FULL, drawn, 0);
    if (PF) { grp = grp_nxt; grp_nxt = drawn; }
    else { grp = drawn; batch_ids(grp, e_cur, nd_cur); }
  }
#if A2DS_ZWAIT == 2
  if (zon) {   // rounds this warp has not reached; the copies read zero_src1 until they complete
    for (; zr < p.zval.rounds; zr++) zero_round(p.zval, zr, gw, n_gw, zsrc, lane);
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
#endif
}


// =====================================================================================
// k_assemble_t<RES, KMAT, GMAT, NL>: the same assembly with the element contraction at the
// tying-point level (mitc4_tying.h): 7 instead of 9 DMMA k-steps for the tangent, 6 instead
// of 8 for the geometric stiffness, which accumulates Z + Z^T directly in the 6 upper tiles
// (two MMAs per off-diagonal tile; the diagonal tiles are symmetrised when the geometric
// blocks are added), so there is no second staging tile.  Components without
// membrane-bending coupling only (run_assembly sends coupled ones to k_assemble).
// =====================================================================================
// Staging tile of k_assemble_t: the 24 x 24 element matrix as 4 x 4 node-pair blocks of 36
// doubles in their BCSR order, block (bi, bj) at ET_R * bi + ET_C * bj.  ET_C = 36 = 4 (mod 16)
// and ET_R = 145 = 1 (mod 16) make every access pattern of the kernel conflict free on the 16
// 8-byte banks a half-warp touches: the MMA accumulator stores (lane = (c, q): block row c >> 1,
// block column q), the scatter (32 consecutive entries of a block; 4 tail entries of 4 blocks)
// and the geometric pass (16 node pairs (a, b) of one u/d class per half-warp).
static const int ET_R = 145, ET_C = 36, ET_SIZE = 3 * ET_R + 3 * ET_C + 36;
__device__ __forceinline__ int et_at(int bi, int rr, int bj, int cc) {
  return ET_R * bi + ET_C * bj + 6 * rr + cc;
}
// ---- zeroing of the output matrices inside k_assemble_t -------------------------------------
// A memset in front of the kernel writes 5.2 GB of zeros that are evicted to DRAM and read back
// by the first RED of every block (0.8 ms of a 5.8 ms step, 3 x the algorithmic traffic).
// Instead every warp zeroes ONE small chunk per trip, a few rounds ahead of where the element
// batches drawn at that time scatter: round r = chunk r of every warp = one contiguous slice of
// the value arrays (zeroing needs the store bandwidth of ALL SMs: a few dedicated zero blocks
// reach only ~30 GB/s each, measured).  The stores are issued at the top of a trip and
// published after the batched phases (bar.warp.sync + red.release.gpu on done[r]: by then the
// warp's earlier REDs and these stores have long drained, so the release costs nothing).
// Before a batch scatters, the warp checks that the round its highest block offset falls into
// has been zeroed by ALL warps (done[r] == number of warps, acquire; rounds complete in order
// per warp, the last check is cached).  A warp that has to wait zeroes ahead instead of idling,
// so the scheme cannot deadlock for any element order (without locality it degenerates to
// zeroing everything first).  All warps are co-resident (cooperative launch).
// The zeros are still in L2 when the REDs arrive: nothing is read from DRAM and every block is
// written to DRAM once.  No __threadfence(), no call: either one, merely present in the
// kernel, costs every variant 15-17 % (measured; profiles/README.md).
#ifndef A2DS_MB_T
#define A2DS_MB_T 2
#endif
#ifndef A2DS_CTA_SYNC
#define A2DS_CTA_SYNC 1
#endif
#ifndef A2DS_ELEM_SYNC
#define A2DS_ELEM_SYNC 0
#endif
struct BatchTmp {   // node-phase outputs only the Gauss-point phase reads; overlaid on the staging
  double dr[12], etn[4], pad_[4];   // tile E, which is only live inside the per-element loop
};                  // 20 doubles = 4 (mod 16)
struct NodeView {   // what phase_node / qp_geometry address as one record
  double *X, *q, *fn, *dr, *wn, *cdr, *etn;
};
struct ElemRecT {   // ElemRec without the batch-phase-only arrays
  double fn[12], wn[12], cdr[36];
  NodeTab t0[4], t1[4];
  QpRec qp[4];
  double pad_[12];  // 548 doubles = 4 (mod 16): the 16 (element, point) lanes hit 16 banks
};
struct WarpScratchT {
  RawBatch raw0;
  int nodes[NB][4];
  ElemRecT rec[NB];
  union {
    double E[ET_SIZE];      // staging tile (block-major, see et_at)
    BatchTmp tmp[NB];
  };
  TyWork work;
  RawBatch raw1;            // double buffer: batch i+1 lands (cp.async) while batch i is processed
  alignas(16) int goff[2][NB][16];
};
static_assert(sizeof(BatchTmp) * NB <= sizeof(double) * ET_SIZE, "batch inputs overlay E");
// block-shared part of the dynamic shared memory, in front of the per-warp scratch: the plans of
// the 45 entries of H_tt (Gauss-point weights and slots; constant)
struct BlockSharedT {
  alignas(16) double zero_src[128];   // source of the bulk copies that zero the matrices (ZeroPlan)
  TyPlan plan[45];
  int draw[2];   // block-synchronous scheduling: first batch of the block in trips t, t + 1
  int pad_[2];
};
static const int BLOCK_SHARED_T = (int)((sizeof(BlockSharedT) + 15) & ~size_t(15));

// record view used by the per-lane functions of mitc4_tying.h (they address ElemRec members)
struct RecView {
  const double *fn, *wn, *cdr;
  const NodeTab *t0, *t1;
  const QpRec *qp;
};

// MMA accumulators of the 6 upper tiles -> staging tile (lower tiles mirrored).  Lane (c, q)
// holds K[3 c + ti][6 q + tj] and K[3 c + ti][6 q + 3 + tj]: block (c >> 1, q).
template <bool SCALE>
__device__ __forceinline__ void stage_tiles_t(double *E, const double (&acc)[6][2], double scale,
                                              int lane) {
  const int c = lane >> 2, q = lane & 3;
  const int up = ET_R * (c >> 1) + ET_C * q + 18 * (c & 1);        // + 6 ti + tj (+3)
  const int lo = ET_R * q + ET_C * (c >> 1) + 3 * (c & 1);         // + 6 tj (+18) + ti
  int idx = 0;
#pragma unroll
  for (int ti = 0; ti < 3; ti++)
#pragma unroll
    for (int tj = ti; tj < 3; tj++, idx++) {
      const double a0 = SCALE ? scale * acc[idx][0] : acc[idx][0];
      const double a1 = SCALE ? scale * acc[idx][1] : acc[idx][1];
      E[up + 6 * ti + tj] = a0;
      E[up + 6 * ti + 3 + tj] = a1;
      if (ti != tj) {
        E[lo + 6 * tj + ti] = a0;
        E[lo + 6 * (3 + tj) + ti] = a1;
      }
    }
}
// scatter of the staged blocks: as scatter_matrix, on the block-major tile
#ifndef A2DS_OFF_LDS
#define A2DS_OFF_LDS 1   // block offsets of the scatter by broadcast LDS.128 from the gathered table (0: by shuffles; measured +0.3 ... 1 %)
#endif
__device__ __forceinline__ void scatter_matrix_t(const double *E, double *vals, int off16, int lane,
                                                 const int *offrow = nullptr) {
  const unsigned FULL = 0xffffffffu;
  const int tb = lane >> 2;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    double v0[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int b = 8 * half + k;
      v0[k] = E[ET_R * (b >> 2) + ET_C * (b & 3) + lane];
    }
    const double v1 = E[ET_R * (2 * half + (tb >> 2)) + ET_C * (tb & 3) + 32 + (lane & 3)];
#if A2DS_OFF_LDS
    const int4 oa = *reinterpret_cast<const int4 *>(offrow + 8 * half);
    const int4 ob = *reinterpret_cast<const int4 *>(offrow + 8 * half + 4);
    const int offs[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
    const int offt = offrow[8 * half + tb];
    (void)FULL; (void)off16;
#pragma unroll
    for (int k = 0; k < 8; k++) atomicAdd(vals + 36 * (size_t)offs[k] + lane, v0[k]);
#else
    const int offt = __shfl_sync(FULL, off16, 8 * half + tb);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int off = __shfl_sync(FULL, off16, 8 * half + k);
      atomicAdd(vals + 36 * (size_t)off + lane, v0[k]);
    }
#endif
    atomicAdd(vals + 36 * (size_t)offt + 32 + (lane & 3), v1);
  }
}


// one chunk (round r, warp gw) of one value array (see ZeroPlan), written by the TMA unit from a
// small zeroed shared-memory buffer (bulk copies issued by lane 0): plain stores of 21 KB per
// trip hold the warp ~3 us at the LSU (measured: in-kernel zeroing then costs what the memset
// did), bulk copies cost it a dozen instructions
static const int ZERO_SRC_BYTES = 1024;
__device__ __forceinline__ void zero_chunk(double2 *zb, long long n2, int c2, int r, int gw, int n_gw,
                                           unsigned src) {
  if (!zb) return;
  const long long s0 = ((long long)r * n_gw + gw) * c2;
  const long long left = n2 - s0;
  int bytes = 16 * (left < c2 ? (left > 0 ? (int)left : 0) : c2);
  unsigned char *dst = reinterpret_cast<unsigned char *>(zb + s0);
  while (bytes > 0) {
    const int n = bytes > ZERO_SRC_BYTES ? ZERO_SRC_BYTES : bytes;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(n) : "memory");
    dst += n; bytes -= n;
  }
}
#if A2DS_ZWAIT == 2
// the same chunk with the bulk copies spread over the lanes (segment i of ZERO_SRC_BYTES by lane
// i mod 32): one predicated instruction per lane instead of a loop in lane 0.  Bulk async-groups
// are per thread: every lane commits, and waits for, its own copies.  (An L2 evict_first hint on
// the copies changes nothing: measured.)
__device__ __forceinline__ void zero_chunk_lanes(double2 *zb, long long n2, int c2, int r, int gw, int n_gw,
                                                 unsigned src, int lane) {
  if (!zb) return;
  const long long s0 = ((long long)r * n_gw + gw) * c2;
  const long long left = n2 - s0;
  const int bytes = 16 * (left < c2 ? (left > 0 ? (int)left : 0) : c2);
  unsigned char *dst = reinterpret_cast<unsigned char *>(zb + s0);
  for (int o = lane * ZERO_SRC_BYTES; o < bytes; o += 32 * ZERO_SRC_BYTES) {
    const int n = bytes - o > ZERO_SRC_BYTES ? ZERO_SRC_BYTES : bytes - o;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + o), "r"(src), "r"(n) : "memory");
  }
}
__device__ __forceinline__ void zero_round(const ZeroPlan &z, int r, int gw, int n_gw, unsigned src, int lane) {
  zero_chunk_lanes(z.zK, z.nK, z.cK, r, gw, n_gw, src, lane);
  zero_chunk_lanes(z.zG, z.nG, z.cG, r, gw, n_gw, src, lane);
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
#else
__device__ __forceinline__ void zero_round(const ZeroPlan &z, int r, int gw, int n_gw, unsigned src, int lane) {
  if (lane == 0) {
    zero_chunk(z.zK, z.nK, z.cK, r, gw, n_gw, src);
    zero_chunk(z.zG, z.nG, z.cG, r, gw, n_gw, src);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
#endif
// round r of this warp is zeroed: wait for the bulk copies, then count it (release)
__device__ __forceinline__ void zero_publish(int *done, int r, int lane) {
  if (lane == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(done + r) : "memory");
  }
}

#ifndef A2DS_GEO_MMA
#define A2DS_GEO_MMA 2   // 0: scalar bending scalars; 1: MMA, pairs in accumulator layout (staging
#endif                   // conflicts: measured slower); 2: MMA + redistribution through shared memory
#ifndef A2DS_G_ORDER
#define A2DS_G_ORDER 1
#endif
// generalised node pair e (0, 1) of a lane in the geometric phase.  Scalar path: a half-warp
// takes the 16 node pairs (a, b) of ONE class (u-u, u-d, d-u, d-d), which are 16 different
// banks of the staging tile (and a uniform fold branch per half-warp)
__device__ __forceinline__ void geo_pair(int lane, int e, int &pr, int &pc) {
#if A2DS_GEO_MMA == 1
  pr = lane >> 2; pc = 2 * (lane & 3) + e;         // accumulator layout of the MMA
#else
  const int g = 2 * e + (lane >> 4);
  pr = (lane & 3) + 4 * (g >> 1); pc = ((lane >> 2) & 3) + 4 * (g & 1);
#endif
}
// The bending scalars of the geometric stiffness,
//   mq_qp(p, pp) = [a_p; b_p]^T [[0, Sigma], [Sigma, 0]] [a_pp; b_pp]       (8 x 8 per Gauss point),
// as one m8n8k4 MMA per Gauss point: A = the coefficient 4-vectors of the 8 generalised nodes,
// B = Sigma4 times them.  Lane (row = lane >> 2, k = lane & 3) receives mq(row, 2 k), mq(row, 2 k + 1).
template <class Rec>
__device__ __forceinline__ void geo_mq_mma(const Rec &gm, const TyWork &wk, int lane, double (&acc)[4][2]) {
  const int row = lane >> 2, k = lane & 3;
  // ca / cb are adjacent [4][8][2] arrays: lane offsets once, the Gauss point is an immediate
  const double *cab = &wk.ca[0][0][0];
  const int offA = 64 * (k >> 1) + 2 * row + (k & 1);   // (k < 2 ? ca : cb)[.][row][k & 1]
  const int offC = 64 * (1 - (k >> 1)) + 2 * row;       // (k < 2 ? cb : ca)[.][row][0..1]
  const int sa = 2 * (k & 1), sb = 2 - (k & 1);         // (s3, s5) or (s5, s4) of (w s3, w s4, w s5)
#pragma unroll
  for (int qp = 0; qp < 4; qp++) {
    const double a = cab[offA + 16 * qp];
    const double *sg = gm.qp[qp].sg;
    const double b = sg[sa] * cab[offC + 16 * qp] + sg[sb] * cab[offC + 16 * qp + 1];
    acc[qp][0] = acc[qp][1] = 0.0;
    dmma884(acc[qp], a, b);
  }
}
// The 64 geometric-stiffness blocks of an element, two per lane (pairs: geo_pair).
template <class Rec>
__device__ __forceinline__ void geo_mq(const Rec &gm, TyWork &wk, int lane, double (&mq)[2][4]) {
#if A2DS_GEO_MMA == 1
  double acc[4][2];
  geo_mq_mma(gm, wk, lane, acc);
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int qp = 0; qp < 4; qp++) mq[e][qp] = acc[qp][e];
#elif A2DS_GEO_MMA == 2
  // MMA, then through shared memory into the class-wise pair map of the staging pass; the
  // scratch overlays H, ca, cb (consumed: H by the column phase, ca / cb by the MMA fragments)
  double acc[4][2];
  geo_mq_mma(gm, wk, lane, acc);
  double *mqs = &wk.H[0];                  // [qp][p][pp], 256 doubles
  __syncwarp();                            // every lane has read its fragments
  const int row = lane >> 2, k = lane & 3;
#pragma unroll
  for (int qp = 0; qp < 4; qp++)
    *reinterpret_cast<double2 *>(&mqs[64 * qp + 8 * row + 2 * k]) = make_double2(acc[qp][0], acc[qp][1]);
  __syncwarp();
#pragma unroll
  for (int e = 0; e < 2; e++) {
    int pr, pc;
    geo_pair(lane, e, pr, pc);
#pragma unroll
    for (int qp = 0; qp < 4; qp++) mq[e][qp] = mqs[64 * qp + 8 * pr + pc];
  }
#else
#pragma unroll
  for (int e = 0; e < 2; e++) {
    int pr, pc;
    geo_pair(lane, e, pr, pc);
#pragma unroll
    for (int qp = 0; qp < 4; qp++)
      mq[e][qp] = geo_bending_mq(wk.ca[qp][pr], wk.cb[qp][pr], wk.ca[qp][pc], wk.cb[qp][pc], gm.qp[qp].sg);
  }
#endif
}

template <bool RES, bool KMAT, bool GMAT, bool NL>
__global__ void __launch_bounds__(MAX_WARPS_PER_BLOCK * 32, A2DS_MB_T) k_assemble_t(const KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  BlockSharedT &bs = *reinterpret_cast<BlockSharedT *>(smem_raw);
  WarpScratchT &ws = *reinterpret_cast<WarpScratchT *>(smem_raw + BLOCK_SHARED_T +
                                                        (size_t)warp * p.scratch_bytes);
  TyWork &wk = ws.work;
  const unsigned FULL = 0xffffffffu;
  Want w;
  w.res = RES; w.kmat = KMAT; w.gmat = GMAT; w.nonlinear = NL; w.thermal = p.thermal;
  const bool need_state = GMAT || NL;
  const int n_groups = (p.n_list + NB - 1) / NB;
  if (threadIdx.x < 45) ty_plan(threadIdx.x, bs.plan[threadIdx.x]);
#if A2DS_ZWAIT
  bs.zero_src[threadIdx.x & 127] = 0.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // read by the async proxy (bulk copies)
  const unsigned zsrc = (unsigned)__cvta_generic_to_shared(bs.zero_src);
#endif
  __syncthreads();
  // lane constants of the column phase: opaque, so that they stay in registers across the
  // element loop instead of being recomputed per element
  LaneConst lc;
  lane_const(lane, lc);
#ifndef A2DS_NO_OPAQUE
  asm volatile("" : "+d"(lc.Nxi), "+d"(lc.Neta), "+d"(lc.N), "+d"(lc.Nq[0]), "+d"(lc.Nq[1]),
                    "+d"(lc.Nq[2]), "+d"(lc.Nq[3]));
#endif

  // zero row / column of H
  if (lane < TY_LD) { wk.H[TY_LD * 9 + lane] = 0.0; wk.H[TY_LD * lane + 9] = 0.0; wk.sigt[lane] = 0.0; }

#if A2DS_ZWAIT
  // zeroing of the output matrices by rounds (see ZeroPlan)
#if A2DS_ZWAIT == 2
  const bool zon = (KMAT || GMAT) && p.zval.rounds > 0;
#else
  const bool zon = (KMAT || GMAT) && p.zplan != nullptr;
#endif
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp, n_gw = gridDim.x * (blockDim.x >> 5);
  int zr = 0;        // rounds this warp has zeroed (the last one possibly not yet published)
  int zknown = -1;   // highest round known complete on all warps
  bool zpend = false;
#if A2DS_ZWAIT == 1
  if (zon) {
    const ZeroPlan &z = A2DS_ZP(p);
    for (; zr < z.rounds && zr < z.ahead; zr++) {
      zero_round(z, zr, gw, n_gw, zsrc, lane);
      zero_publish(z.done, zr, lane);
    }
  }
#endif
  (void)zknown; (void)zpend;
#endif

  auto batch_ids = [&](int grp_, int &e_out, int &nd_out) {
    const int j = (lane >> 2) & (NB - 1);
    const int idx = grp_ * NB + j;
    e_out = -1; nd_out = 0;
    if (grp_ < n_groups && idx < p.n_list) {
      e_out = p.elem_list ? __ldg(&p.elem_list[idx]) : idx;
      nd_out = __ldg(&p.conn[4 * e_out + (lane & 3)]);
    }
  };
  // asynchronous gather of a batch: lane = (element j, node m) fetches its node's coordinates
  // and state (the rows of X / u are 24 / 48 bytes) and a quarter of the element's offsets
  auto issue_gather = [&](RawBatch &rb, int (*goffb)[16], int e_l, int nd_l) {
    if (lane < 4 * NB && e_l >= 0) {
      const int j = lane >> 2, m = lane & 3;
      const double *xs = &p.X[3 * (size_t)nd_l];
      cp_async8(&rb.xq[j][3 * m], xs);
      cp_async8(&rb.xq[j][3 * m + 1], xs + 1);
      cp_async8(&rb.xq[j][3 * m + 2], xs + 2);
      const double *us = &p.u[6 * (size_t)nd_l];
      cp_async16(&rb.xq[j][12 + 6 * m], us);
      cp_async16(&rb.xq[j][12 + 6 * m + 2], us + 2);
      cp_async16(&rb.xq[j][12 + 6 * m + 4], us + 4);
      if (KMAT && p.Koff) cp_async16(&rb.koff[j][4 * m], &p.Koff[16 * (size_t)e_l + 4 * m]);
      if (GMAT) cp_async16(&goffb[j][4 * m], &p.Goff[16 * (size_t)e_l + 4 * m]);
      if (m == 0) cp_async4(&rb.comp[j], &p.elem_comp[e_l]);
    }
  };
  auto next_group = [&]() {
    int g = 0;
    if (lane == 0) g = atomicAdd(p.work_counter, 1);
    return __shfl_sync(FULL, g, 0);
  };
  // Block-synchronous trips for the big variants: the warps of a block draw their batches
  // together (one atomic per block and trip) and start every trip at a block barrier, so that
  // they walk the same code at the same time — the batched phases are ~45 KB of straight-line
  // SASS per trip and the instruction cache serves four warps with one fetch.  Measured:
  // -4 % for the fused / geometric-stiffness kernels, +3 % for the residual / linear tangent
  // ones (short trips against the barrier wait), which keep drawing per warp.
  constexpr bool CS = A2DS_CTA_SYNC && GMAT;
  const int wpb = blockDim.x >> 5;
  int base_cur = 0, trip = 0;
  int grp, grp_nxt;
  if constexpr (CS) {
    if (threadIdx.x == 0) {
      bs.draw[0] = atomicAdd(p.work_counter, wpb);
      bs.draw[1] = atomicAdd(p.work_counter, wpb);
    }
    __syncthreads();
    base_cur = bs.draw[0];
    grp = base_cur + warp; grp_nxt = bs.draw[1] + warp;
  } else {
    grp = next_group();
    grp_nxt = next_group();
  }
  int buf = 0;
  int e_cur, nd_cur;
  batch_ids(grp, e_cur, nd_cur);
  if (grp < n_groups) issue_gather(ws.raw0, ws.goff[0], e_cur, nd_cur);
  for (; CS ? (base_cur < n_groups) : (grp < n_groups); buf ^= 1, trip++) {
    int drawn = 0, base_nxt = 0;
    if constexpr (CS) {
      __syncthreads();
      base_nxt = bs.draw[(trip + 1) & 1];   // drawn one trip ago (or in the prologue)
      grp_nxt = base_nxt + warp;
      if (threadIdx.x == 0) drawn = atomicAdd(p.work_counter, wpb);   // for trip + 2, stored at the end
    } else {
      if (lane == 0) drawn = atomicAdd(p.work_counter, 1);
    }
    const int base = grp * NB;
    const int cnt = grp < n_groups ? min(NB, p.n_list - base) : 0;
#if A2DS_ZWAIT
    if (zon && zr < A2DS_ZP(p).rounds) {   // issued here, published after the batched phases
      zero_round(A2DS_ZP(p), zr, gw, n_gw, zsrc, lane);
#if A2DS_ZWAIT == 1
      zpend = true;
#else
      zr++;   // spare array: nobody waits for it inside this launch
#endif
    }
#endif
    int e_nxt = -1, nd_nxt = 0;
    batch_ids(grp_nxt, e_nxt, nd_nxt);
    cp_async_wait_all();
    __syncwarp();
    const RawBatch &rb = buf ? ws.raw1 : ws.raw0;
    const int (*goffb)[16] = ws.goff[buf];
    if (lane < 4 * NB) ws.nodes[lane >> 2][lane & 3] = nd_cur;
    // ---- batched phases: lane = (element of the batch, node | Gauss point) ----------------
    const int jb = (lane >> 2) & (NB - 1);
    const bool act = lane < 4 * NB && jb < cnt;
    NodeView nv;
    nv.X = const_cast<double *>(&rb.xq[jb][0]); nv.q = const_cast<double *>(&rb.xq[jb][12]);
    nv.dr = ws.tmp[jb].dr; nv.etn = ws.tmp[jb].etn;
    nv.fn = ws.rec[jb].fn; nv.wn = ws.rec[jb].wn; nv.cdr = ws.rec[jb].cdr;
    if (act) phase_node(p.comps[rb.comp[jb]], nv, lane & 3);
    __syncwarp();
    if (act) {
      ElemRecT &rc = ws.rec[jb];
      const int m = lane & 3;
      node_tab(rc.t0[m], nv.X, 3, nv.fn, &nv.fn[3 * m], m);
      if (need_state) node_tab(rc.t1[m], nv.q, 6, nv.dr, &nv.fn[3 * m], m);
      phase_qp_t(p.comps[rb.comp[jb]], nv, rc.qp[lane & 3], lane & 3, w);
    }
    if (grp_nxt < n_groups)
      issue_gather(buf ? ws.raw0 : ws.raw1, ws.goff[buf ^ 1], e_nxt, nd_nxt);
    e_cur = e_nxt; nd_cur = nd_nxt;
    __syncwarp();
#if A2DS_ZWAIT == 1
    if (zon) {
      const int n_zr = A2DS_ZP(p).rounds;
      int *zdone = A2DS_ZP(p).done;
      if (zpend) { zero_publish(zdone, zr, lane); zr++; zpend = false; }
      // the round the highest block offset of this batch falls into must be complete (rounded
      // up: one round early costs nothing, one round late is a race)
      int mxw = 0;
#pragma unroll
      for (int r = 0; r < NB * 16 / 32; r++) {
        const int sidx = lane + 32 * r, j = sidx >> 4, k = sidx & 15;
        if (j < cnt) {
          if (KMAT) mxw = max(mxw, rb.koff[j][k]);
          if (GMAT) mxw = max(mxw, goffb[j][k]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mxw = max(mxw, __shfl_xor_sync(FULL, mxw, o));
      const int rneed = min(n_zr - 1, (int)((float)mxw * A2DS_ZP(p).inv_round) + 1);
      while (rneed > zknown) {
        int v = 0;
        if (lane == 0) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(zdone + rneed) : "memory");
        v = __shfl_sync(FULL, v, 0);
        if (v >= n_gw) { zknown = rneed; break; }
        if (zr < n_zr) {   // zero ahead instead of idling
          zero_round(A2DS_ZP(p), zr, gw, n_gw, zsrc, lane);
          zero_publish(zdone, zr, lane);
          zr++;
        }
      }
    }
#endif

    // element prologue: H_tt (45 entries over 32 lanes: Gauss-point sums of Q) and the tying
    // stresses.  A pure latency chain (plan -> Q -> 4 FMA -> store); it is issued for element
    // j + 1 right before the last scatter of element j, whose REDs cover it.
    auto prologue = [&](int jn) {
      const ElemRecT &rn = ws.rec[jn];
      RecView gn;
      gn.fn = rn.fn; gn.wn = rn.wn; gn.cdr = rn.cdr; gn.t0 = rn.t0; gn.t1 = rn.t1; gn.qp = rn.qp;
      if (KMAT || GMAT) {
        ty_H_entry(gn, bs.plan[lane], wk.H);
        if (lane + 32 < 45) ty_H_entry(gn, bs.plan[lane + 32], wk.H);
#if A2DS_GEO_MMA == 2
        // the geometric phase used H as scratch: restore its zero row / column
        if (need_state && lane < TY_LD) { wk.H[TY_LD * 9 + lane] = 0.0; wk.H[TY_LD * lane + 9] = 0.0; }
#endif
      }
      if ((RES || need_state) && lane < 9) ty_sum_stress(gn, wk, lane);
    };
    if (cnt > 0) prologue(0);
    __syncwarp();
#pragma unroll 1
    for (int j = 0; j < (A2DS_ELEM_SYNC && CS ? NB : cnt); j++) {
#if A2DS_ELEM_SYNC
      if constexpr (CS) {
        __syncthreads();
        if (j >= cnt) continue;
      }
#endif
      const ElemRecT &rc = ws.rec[j];
      RecView gm;
      gm.fn = rc.fn; gm.wn = rc.wn; gm.cdr = rc.cdr; gm.t0 = rc.t0; gm.t1 = rc.t1; gm.qp = rc.qp;
      // (H_tt and the tying stresses of this element were formed by `prologue` during the
      // previous element's scatter — or in front of the loop for the first one)

      // ---- column phase: the lane's rows of Bt, W = H Bt (and Bt1) ARE the DMMA fragments ----
      LaneFrag f;
      double B1[6][3];
      lane_fragments(gm, wk, lane, lc, w, f, B1);
      if (RES) {
        double r3[3];
        lane_residual(gm, wk, lane, f, r3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          r3[k] += __shfl_xor_sync(FULL, r3[k], 1);
          r3[k] += __shfl_xor_sync(FULL, r3[k], 2);
        }
        if ((lane & 3) == 0) {
          double *r = &p.res[6 * (size_t)ws.nodes[j][lane_m(lane)] + 3 * lane_h(lane)];
          atomicAdd(r, p.res_scale * r3[0]);
          atomicAdd(r + 1, p.res_scale * r3[1]);
          atomicAdd(r + 2, p.res_scale * r3[2]);
        }
      }
      if (KMAT) {
        double kacc[6][2];
#pragma unroll
        for (int t = 0; t < 6; t++) kacc[t][0] = kacc[t][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < TY_ROWS; ks++) {
          int idx = 0;
#pragma unroll
          for (int ti = 0; ti < 3; ti++)
#pragma unroll
            for (int tj = ti; tj < 3; tj++, idx++) dmma884(kacc[idx], f.B[ks][ti], f.W[ks][tj]);
        }
        stage_tiles_t<true>(ws.E, kacc, p.alpha, lane);
      }
      double gacc[6][2];
      if (GMAT) {
        // G = Bt1^T W + W^T Bt1 in the 6 upper tiles: off-diagonal tiles take both products,
        // diagonal tiles take Z_tt = Bt1_t^T W_t and are symmetrised below
#pragma unroll
        for (int t = 0; t < 6; t++) gacc[t][0] = gacc[t][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 6; ks++) {
          // all six B1^T W products first, then the three W^T B1 of the off-diagonal tiles:
          // no MMA waits on the accumulator of the one issued just before it
          int idx = 0;
#pragma unroll
          for (int ti = 0; ti < 3; ti++)
#pragma unroll
            for (int tj = ti; tj < 3; tj++, idx++) {
              dmma884(gacc[idx], B1[ks][ti], f.W[ks][tj]);
#if !A2DS_G_ORDER
              if (ti != tj) dmma884(gacc[idx], f.W[ks][ti], B1[ks][tj]);
#endif
            }
#if A2DS_G_ORDER
          dmma884(gacc[1], f.W[ks][0], B1[ks][1]);
          dmma884(gacc[2], f.W[ks][0], B1[ks][2]);
          dmma884(gacc[4], f.W[ks][1], B1[ks][2]);
#endif
        }
      }
      __syncwarp();   // E staged; coefficient pairs of the geometric phase published
      if (KMAT) {
        if (NL) {
          double blk[2][9], mq[2][4];
          geo_mq(gm, wk, lane, mq);
#pragma unroll
          for (int e = 0; e < 2; e++) {
            int pr, pc;
            geo_pair(lane, e, pr, pc);
            geo_block_from_mq(gm, wk, pr, pc, mq[e], blk[e]);
            const int at = et_at(pr & 3, 3 * (pr >> 2), pc & 3, 3 * (pc >> 2));
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
              for (int jj = 0; jj < 3; jj++) ws.E[at + 6 * i + jj] += p.alpha * blk[e][3 * i + jj];
          }
          __syncwarp();
        }
        if (!GMAT && j + 1 < cnt) prologue(j + 1);   // covered by the scatter below
        if (NL && !GMAT && p.jvp_x) {
          const int node = ws.nodes[j][(lane / 6) & 3];
          double xv = 0.0;
          if (lane < 24) xv = p.jvp_x[6 * (size_t)node + lane % 6];
          double y = 0.0;
#pragma unroll
          for (int k = 0; k < 24; k++) {
            const double xk = __shfl_sync(FULL, xv, k);
            if (lane < 24) y += ws.E[et_at(lane / 6, lane % 6, k / 6, k % 6)] * xk;
          }
          if (lane < 24) atomicAdd(&p.jvp_y[6 * (size_t)node + lane % 6], p.jvp_scale * y);
        } else {
          scatter_matrix_t(ws.E, p.Kval, rb.koff[j][lane & 15], lane, rb.koff[j]);
        }
      }
      if (GMAT) {
        if (KMAT) __syncwarp();   // E is reused for G once K has left
        stage_tiles_t<false>(ws.E, gacc, 1.0, lane);
        __syncwarp();
        double v[2][9], mq[2][4];
        geo_mq(gm, wk, lane, mq);
#pragma unroll
        for (int e = 0; e < 2; e++) {
          int pr, pc;
          geo_pair(lane, e, pr, pc);
          geo_block_from_mq(gm, wk, pr, pc, mq[e], v[e]);
          const int at = et_at(pr & 3, 3 * (pr >> 2), pc & 3, 3 * (pc >> 2));
          const int tr = et_at(pc & 3, 3 * (pc >> 2), pr & 3, 3 * (pr >> 2));   // transposed block
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int jj = 0; jj < 3; jj++) {
              double z = ws.E[at + 6 * i + jj];
              if (i == jj) z += ws.E[tr + 6 * jj + i];   // diagonal tiles: Z + Z^T
              v[e][3 * i + jj] = p.gscale * (v[e][3 * i + jj] + z);
            }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 2; e++) {
          int pr, pc;
          geo_pair(lane, e, pr, pc);
          const int at = et_at(pr & 3, 3 * (pr >> 2), pc & 3, 3 * (pc >> 2));
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int jj = 0; jj < 3; jj++) ws.E[at + 6 * i + jj] = v[e][3 * i + jj];
        }
        __syncwarp();
        if (j + 1 < cnt) prologue(j + 1);   // H and the tying stresses are free from here on
        scatter_matrix_t(ws.E, p.Gval, goffb[j][lane & 15], lane, goffb[j]);
      }
      if (!GMAT && !KMAT && j + 1 < cnt) prologue(j + 1);
      __syncwarp();
    }
    if constexpr (CS) {
      // slot trip & 1 held this trip's base, read by everybody before this trip's barrier
      if (threadIdx.x == 0) bs.draw[trip & 1] = drawn;
      base_cur = base_nxt; grp = grp_nxt;
    } else {
      drawn = __shfl_sync(FULL, drawn, 0);
      grp = grp_nxt; grp_nxt = drawn;
    }
  }
#if A2DS_ZWAIT
  // rounds this warp has not reached yet (short lists, warps without a batch)
  if (zon) {
    const ZeroPlan &z = A2DS_ZP(p);
#if A2DS_ZWAIT == 1
    if (zpend) { zero_publish(z.done, zr, lane); zr++; }
    for (; zr < z.rounds; zr++) {
      zero_round(z, zr, gw, n_gw, zsrc, lane);
      zero_publish(z.done, zr, lane);
    }
#else
    for (; zr < z.rounds; zr++) zero_round(z, zr, gw, n_gw, zsrc, lane);
    // the bulk copies read bs.zero_src: they must have completed before the block retires
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
  }
#endif
}

// ---- mass path (gamma terms of assembleJacobian, TACS_MASS_MATRIX, inertial residual) ----
// M_e = sum_qp w (m0 N^T N on u, m1 on u-d, m2 on d-d) folded onto the rotations
// (TACSShellElement.h:410-447, 614-648); RES adds M_e * p.u (p.u = second time derivative
// of the state) to the residual, MAT adds p.alpha * M_e to the matrix p.Kval.
// Memory-bound (2.6 kB of matrix per element, a few hundred flops): same batched geometry
// phases and scatter as k_assemble, no tensor-core part.
template <bool RES, bool MAT>
__global__ void __launch_bounds__(MAX_WARPS_PER_BLOCK * 32, 3) k_mass(const KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  WarpScratch &ws = *reinterpret_cast<WarpScratch *>(smem_raw + (size_t)warp * p.scratch_bytes);
  const unsigned FULL = 0xffffffffu;
  const int n_groups = (p.n_list + NB - 1) / NB;
  (void)warps_per_block;
  // batches drawn from the work counter, as in the element kernels (dynamic scheduling)
  for (;;) {
    int grp = 0;
    if (lane == 0) grp = atomicAdd(p.work_counter, 1);
    grp = __shfl_sync(FULL, grp, 0);
    if (grp >= n_groups) break;
    const int base = grp * NB, cnt = min(NB, p.n_list - base);
    // lane = 4 j + m: node m of element j
    const int jl = (lane >> 2) & (NB - 1);
    int e_l = -1, nd_l = 0;
    if (jl < cnt) {
      e_l = p.elem_list ? __ldg(&p.elem_list[base + jl]) : base + jl;
      nd_l = __ldg(&p.conn[4 * e_l + (lane & 3)]);
    }
    if (lane < 4 * NB && e_l >= 0) {
      const int m = lane & 3;
#pragma unroll
      for (int k = 0; k < 3; k++) ws.geo[jl].X[3 * m + k] = p.X[3 * (size_t)nd_l + k];
#pragma unroll
      for (int k = 0; k < 6; k++) ws.geo[jl].q[6 * m + k] = RES ? p.u[6 * (size_t)nd_l + k] : 0.0;
      if (m == 0) ws.raw0.comp[jl] = __ldg(&p.elem_comp[e_l]);
    }
    if (MAT) {
#pragma unroll
      for (int r = 0; r < NB * 16 / 32; r++) {
        const int sidx = lane + 32 * r, j = sidx >> 4, k = sidx & 15;
        const int ej = __shfl_sync(FULL, e_l, (4 * j) & 31);
        if (ej >= 0) ws.raw0.koff[j][k] = __ldg(&p.Koff[16 * (size_t)ej + k]);
      }
    }
    __syncwarp();
    if (lane < 4 * NB && jl < cnt) mass_node(ws.geo[jl], lane & 3);
    __syncwarp();
    if (lane < 4 * NB && jl < cnt) mass_qp(ws.geo[jl], lane & 3);
    __syncwarp();
#pragma unroll 1
    for (int j = 0; j < cnt; j++) {
      const ElemGeom &gm = ws.geo[j];
      const CompData &c = p.comps[ws.raw0.comp[j]];
#pragma unroll
      for (int pass = 0; pass < 2; pass++) {
        const int pair = lane + 32 * pass, pr = pair >> 3, pc = pair & 7;
        double blk[9];
        mass_block(c, gm, pr, pc, blk);
        const int r0 = 6 * (pr & 3) + (pr >= 4 ? 3 : 0), c0 = 6 * (pc & 3) + (pc >= 4 ? 3 : 0);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int k = 0; k < 3; k++) ws.E[(r0 + i) * KE_LD + c0 + k] = blk[3 * i + k];
      }
      __syncwarp();
      if (RES) {
        const int node = __shfl_sync(FULL, nd_l, 4 * j + (lane / 6 & 3));
        if (lane < 24) {
          double r = 0.0;
#pragma unroll
          for (int k = 0; k < 24; k++) r += ws.E[lane * KE_LD + k] * gm.q[k];
          atomicAdd(&p.res[6 * (size_t)node + lane % 6], p.res_scale * r);
        }
      }
      if (MAT) scatter_matrix<true>(ws.E, p.Kval, ws.raw0.koff[j][lane & 15], lane, p.alpha);
      __syncwarp();
    }
  }
}

#endif
