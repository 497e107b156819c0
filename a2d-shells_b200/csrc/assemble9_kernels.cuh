// assemble9_kernels.cuh — k_assemble9<RES, KMAT, GMAT, NL>: residual, tangent and geometric
// stiffness of the 9-node MITC shells (TACSQuad9Shell, NL: TACSQuad9NonlinearShell) assembled
// into the same 6 x 6 BCSR matrices and node vectors as the 4-node path.  Replaces, for these
// element classes, the element loop of TACSAssembler::assembleRes / assembleJacobian /
// assembleMatType(STIFFNESS, GEOMETRIC_STIFFNESS) (src/TACSAssembler.cpp:4000-4242) with
// TACSShellElement::addResidual / addJacobian / getMatType inside
// (src/elements/shell/TACSShellElement.h:303-771).  The geometric stiffness is the exact
// linear-in-(state, temperature) part of the nonlinear tangent, G = L^T C B1(q) + B1(q)^T C L +
// sum_i sigma_i d2e_i, where the reference takes a central difference of that tangent (:705-751).
//
// One thread block per element in flight, elements drawn from a counter.  The element math
// (mitc9_math.h) runs as block-level phases over shared memory.  Warp 7 is the PRODUCER: it
// works one element ahead and fills the geometry record of the next element (gather, 9 node
// frames with the reference's rounding, 28 tying points, 9 Gauss-point frames, strains and
// stresses of the state) while warps 0..6, the CONSUMERS, turn the current record into the
// derivative tables (28 x 54 tying, 9 x 54 drill) and, per row of three Gauss points, the
// 3 x 54 columns of B and w det C B.  The tangent K = sum B^T (w det C B) is contracted on the
// FP64 tensor path (mma.m8n8k4): consumer warp w owns the 8 rows 8 w .. 8 w + 7 of the 54 x 54
// (padded to 56) element matrix and accumulates the tiles on and right of the diagonal (28 of
// 49, the matrix is symmetric); the accumulators are added to the BCSR blocks straight from
// registers, mirrored for the off-diagonal tiles — there is no staged element matrix.  The two
// groups meet at one block barrier per element; the consumers synchronise among themselves on
// a named barrier.
#ifndef A2DS_ASSEMBLE9_KERNELS_CUH
#define A2DS_ASSEMBLE9_KERNELS_CUH

#include "mitc9_math.h"

static const int Q9_CONSUMERS = 224; // 7 warps: one per 8-row tile of the 56 x 56 padded matrix
static const int Q9_THREADS = 256;   // + the producer warp
static const int Q9_QB = 3;          // Gauss points per contraction batch (one eta row)
#ifndef A2DS_Q9_MINB
#define A2DS_Q9_MINB 2   // thread blocks per SM the register budget is sized for
#endif

struct Meta9 {
  int elem;       // element of the record, -1: the list is exhausted
  int comp;
  int nodes[9];
  int off[81];
};
struct Elem9Block {
  a2ds::Elem9 E[2];   // record of the element being contracted / being prepared
  Meta9 M[2];
  a2ds::Shape9 H;     // shape function tables, filled once per block
  a2ds::Tab9 T;       // derivative tables of the element being contracted (ends with B, CB)
  double B2[Q9_QB - 1][a2ds::Q9_KROWS][a2ds::Q9_LDB];    // B / CB of the other points of the batch
  double CB2[Q9_QB - 1][a2ds::Q9_KROWS][a2ds::Q9_LDB];
  // Once the last batch is contracted the tables are dead and hold the element matrix on its
  // way out: T.B .. CB2 (4896 doubles, contiguous) the 54 x 54 contraction result, T.Gt .. T.Gt1
  // (3136 doubles) the 54 x 54 geometric term, both in BCSR block order (k9_at).
};
static_assert(offsetof(Elem9Block, B2) == offsetof(Elem9Block, T) + offsetof(a2ds::Tab9, CB) +
                                              sizeof(double) * a2ds::Q9_KROWS * a2ds::Q9_LDB,
              "T.B, T.CB, B2, CB2 must be contiguous");
static_assert(offsetof(a2ds::Tab9, Gt1) == offsetof(a2ds::Tab9, Gt) + sizeof(double) * a2ds::Q9_NTY * a2ds::Q9_LD,
              "T.Gt, T.Gt1 must be contiguous");
// entry (r, c) of the 54 x 54 element matrix in the order the 81 BCSR blocks want it
__device__ __forceinline__ int k9_at(int r, int c) {
  const int br = r / 6, bc = c / 6;
  return 36 * (9 * br + bc) + 6 * (r - 6 * br) + (c - 6 * bc);
}

__device__ __forceinline__ void q9_consumer_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(Q9_CONSUMERS) : "memory");
}

// KMAT and GMAT are separate instantiations (one w det C B table): <.., 1, 0, NL> tangent
// (alpha K into Kval through Koff), <0, 0, 1, 0> geometric stiffness (gscale G into Gval through Goff)
template <bool RES, bool KMAT, bool GMAT, bool NL>
__global__ void __launch_bounds__(Q9_THREADS, A2DS_Q9_MINB) k_assemble9(KParams p) {
  static_assert(!(KMAT && GMAT), "tangent and geometric stiffness are separate launches");
  constexpr bool STATE = RES || GMAT || NL;   // strains / stresses of the state are needed
  constexpr bool BIL = GMAT || NL;            // bilinear strain terms are needed
  using namespace a2ds;
  extern __shared__ __align__(16) unsigned char smem9[];
  Elem9Block &S = *reinterpret_cast<Elem9Block *>(smem9);
  const Shape9 &H = S.H;
  const unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 46) q9_shape_tables(S.H, tid);
  // rows 9..11 and columns 54, 55 of the B tables stay zero
  for (int i = tid; i < Q9_KROWS * Q9_LDB; i += Q9_THREADS) {
    (&S.T.B[0][0])[i] = 0.0; (&S.T.CB[0][0])[i] = 0.0;
    for (int qq = 0; qq < Q9_QB - 1; qq++) { (&S.B2[qq][0][0])[i] = 0.0; (&S.CB2[qq][0][0])[i] = 0.0; }
  }
  __syncthreads();

  // ---- producer: the geometry record of the next element of the list (one warp) ----
  auto produce = [&](Elem9 &E, Meta9 &M) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(p.work_counter, 1);
    idx = __shfl_sync(FULL, idx, 0);
    if (idx >= p.n_list) {
      if (lane == 0) M.elem = -1;
      return;
    }
    const int e = p.elem_list ? __ldg(&p.elem_list[idx]) : idx;
    if (lane == 0) { M.elem = e; M.comp = __ldg(&p.elem_comp[e]); }
    if (lane < 9) M.nodes[lane] = __ldg(&p.conn[9 * (size_t)e + lane]);
    if (KMAT || GMAT) {
      const int *off = GMAT ? p.Goff : p.Koff;
      for (int k = lane; k < 81; k += 32) M.off[k] = __ldg(&off[81 * (size_t)e + k]);
    }
    __syncwarp();
    const CompData &c = p.comps[M.comp];
    if (lane < 27) E.X[lane] = __ldg(&p.X[3 * (size_t)M.nodes[lane / 3] + lane % 3]);
    for (int d = lane; d < Q9_NV; d += 32) E.q[d] = __ldg(&p.u[6 * (size_t)M.nodes[d / 6] + d % 6]);
    __syncwarp();
    if (lane < 9) q9_node(c, E, lane);
    __syncwarp();
    if (lane < Q9_NTY) q9_tying(E, H, lane, NL);
    if (lane < 9) q9_qp(c, E, H, lane, BIL);
    __syncwarp();
    if (STATE && lane < 9) q9_qp_state(c, E, H, lane, p.thermal, NL, BIL);
    if (BIL) {
      __syncwarp();
      if (lane < Q9_NTY) q9_sigt(E, H, lane);
    }
  };

  // ---- consumers: tables, columns of B, contraction, scatter of one record (7 warps) ----
  auto consume = [&](const Elem9 &E, const Meta9 &M) {
    const CompData &c = p.comps[M.comp];
    Tab9 &Tb = S.T;
    auto Bq = [&](int qq) -> double(*)[Q9_LDB] { return qq == 0 ? Tb.B : S.B2[qq - 1]; };
    auto CBq = [&](int qq) -> double(*)[Q9_LDB] { return qq == 0 ? Tb.CB : S.CB2[qq - 1]; };
    if (KMAT || GMAT) {
      // the previous element's matrix went out through the B / CB tables: their padding (rows
      // 9..11, columns 54, 55), which the contraction reads, has to be zero again
      double *tab = &Tb.B[0][0];   // T.B, T.CB, B2, CB2: 2 * Q9_QB tables of Q9_KROWS x Q9_LDB
      const int per = (Q9_KROWS - 9) * Q9_LDB + 9 * (Q9_LDB - Q9_NV);
      for (int i = tid; i < 2 * Q9_QB * per; i += Q9_CONSUMERS) {
        const int t = i / per, k = i - per * t;
        const int at = k < (Q9_KROWS - 9) * Q9_LDB ? 9 * Q9_LDB + k
                                                   : Q9_LDB * ((k - (Q9_KROWS - 9) * Q9_LDB) / (Q9_LDB - Q9_NV)) + Q9_NV +
                                                         (k - (Q9_KROWS - 9) * Q9_LDB) % (Q9_LDB - Q9_NV);
        tab[Q9_KROWS * Q9_LDB * t + at] = 0.0;
      }
    }
    for (int i = tid; i < Q9_NTY * Q9_NV; i += Q9_CONSUMERS) {
      Tb.Gt[i / Q9_NV][i % Q9_NV] = q9_gt(E, H, i / Q9_NV, i % Q9_NV);
      if (BIL) Tb.Gt1[i / Q9_NV][i % Q9_NV] = q9_gt1(E, H, i / Q9_NV, i % Q9_NV);
    }
    for (int i = tid; i < Q9_NN * Q9_NV; i += Q9_CONSUMERS) Tb.Dn[i / Q9_NV][i % Q9_NV] = q9_dn(E, H, i / Q9_NV, i % Q9_NV);
    q9_consumer_barrier();
    double acc[7][2];
#pragma unroll
    for (int t = 0; t < 7; t++) acc[t][0] = acc[t][1] = 0.0;
    double r = 0.0;
    for (int b = 0; b < 3; b++) {
      if (tid < Q9_QB * Q9_NV) {
        const int qq = tid / Q9_NV, col = tid - Q9_NV * qq, q = Q9_QB * b + qq;
        double Bk[9], B1k[9], Sk[9];
        q9_bcol(E, Tb, H, q, col, Bk, BIL ? B1k : nullptr);
        if (NL) {   // B of the nonlinear model = L + Bil(q, .)
#pragma unroll
          for (int k = 0; k < 9; k++) Bk[k] += B1k[k];
        }
        q9_stress(c.Cs, GMAT ? B1k : Bk, Sk);
        double(*Bt)[Q9_LDB] = Bq(qq), (*Ct)[Q9_LDB] = CBq(qq);
        const double w = E.qw[q];
#pragma unroll
        for (int k = 0; k < 9; k++) { Bt[k][col] = Bk[k]; Ct[k][col] = w * Sk[k]; }
      }
      q9_consumer_barrier();
      if (KMAT || GMAT) {
        // K = B^T (w det C B) is symmetric: warp w contracts the four tiles (w, (w + d) mod 7),
        // d = 0..3 — every unordered tile pair exactly once, the same work for every warp;
        // Z = L^T (w det C B1): all 49 tiles (d = 0..6), G = Z + Z^T is formed by the scatter
#pragma unroll
        for (int qq = 0; qq < Q9_QB; qq++) {
          const double(*Bt)[Q9_LDB] = Bq(qq), (*Ct)[Q9_LDB] = CBq(qq);
#pragma unroll
          for (int ks = 0; ks < 3; ks++) {
            const int kr = 4 * ks + (lane & 3);
            const double a = Bt[kr][8 * warp + (lane >> 2)];
#pragma unroll
            for (int d = 0; d < (GMAT ? 7 : 4); d++) {
              int tj = warp + d;
              if (tj >= 7) tj -= 7;
              dmma884(acc[d], a, Ct[kr][8 * tj + (lane >> 2)]);
            }
          }
        }
      }
      if (RES && tid < Q9_NV) {
#pragma unroll
        for (int qq = 0; qq < Q9_QB; qq++) {
          const double(*Bt)[Q9_LDB] = Bq(qq);
          const double *s = E.sq[Q9_QB * b + qq];
#pragma unroll
          for (int k = 0; k < 9; k++) r += Bt[k][tid] * s[k];
        }
      }
      q9_consumer_barrier();
    }
    if (RES && tid < Q9_NV)
      atomicAdd(&p.res[6 * (size_t)M.nodes[tid / 6] + tid % 6], p.res_scale * r);
    if (KMAT || GMAT) {
      double *vals = GMAT ? p.Gval : p.Kval;
      const double scale = GMAT ? p.gscale : p.alpha;
      // the tables are dead (barrier at the end of the last batch): stage the accumulators in
      // block order — accumulator (row, col) of tile (warp, tj): rows 8 warp + lane / 4, columns
      // 8 tj + 2 (lane % 4) + i; K: off-diagonal tiles mirrored — and the geometric term beside it
      double *Ke = &Tb.B[0][0], *Kg = &Tb.Gt[0][0];
      const int gr = 8 * warp + (lane >> 2);
      if (gr < Q9_NV) {
#pragma unroll
        for (int d = 0; d < (GMAT ? 7 : 4); d++) {
          int tj = warp + d;
          if (tj >= 7) tj -= 7;
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const int gc = 8 * tj + 2 * (lane & 3) + i;
            if (gc >= Q9_NV) continue;
            Ke[k9_at(gr, gc)] = acc[d][i];
            if (!GMAT && d > 0) Ke[k9_at(gc, gr)] = acc[d][i];   // the mirrored tile of the symmetric K
          }
        }
      }
      // geometric term of the 81 node pairs (four 3 x 3 quadrants each): all of it for G, the
      // stress term of the nonlinear tangent
      if (BIL) {
        for (int item = tid; item < 4 * 81; item += Q9_CONSUMERS) {
          const int pair = item >> 2, ha = (item >> 1) & 1, hb = item & 1;
          double o[9];
          q9_geo_quadrant(E, H, pair / 9, pair % 9, ha, hb, o);
          double *dst = Kg + 36 * pair + 18 * ha + 3 * hb;
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) dst[6 * i + j] = o[3 * i + j];
        }
      }
      q9_consumer_barrier();
      // 81 blocks x 36 entries leave as coalesced REDs: consecutive threads, consecutive doubles
      // of a block.  G = Z + Z^T + geometric term: the transposed entry is read from the tile.
      for (int idx = tid; idx < 81 * 36; idx += Q9_CONSUMERS) {
        const int blk = idx / 36, ent = idx - 36 * blk;
        double v = Ke[idx];
        if (GMAT) {
          const int br = blk / 9, bc = blk - 9 * br, rr = ent / 6, cc = ent - 6 * rr;
          v += Ke[36 * (9 * bc + br) + 6 * cc + rr];
        }
        if (BIL) v += Kg[idx];
        atomicAdd(vals + 36 * (size_t)M.off[blk] + ent, scale * v);
      }
    }
  };

  int cur = 0;
  if (warp == 7) produce(S.E[0], S.M[0]);
  for (;;) {
    __syncthreads();   // record `cur` is complete, record `cur ^ 1` is free again
    if (S.M[cur].elem < 0) break;
#if defined(A2DS_Q9_TIMING_NO_CONSUME)   // timing experiments only: results are wrong
    if (warp == 7) produce(S.E[cur ^ 1], S.M[cur ^ 1]);
#elif defined(A2DS_Q9_TIMING_NO_PRODUCE)
    if (warp == 7) { if (lane == 0) S.M[cur ^ 1].elem = (int)atomicAdd(p.work_counter, 1) < p.n_list ? 0 : -1; }
    else consume(S.E[cur], S.M[cur]);
#else
    if (warp == 7) produce(S.E[cur ^ 1], S.M[cur ^ 1]);
    else consume(S.E[cur], S.M[cur]);
#endif
    cur ^= 1;
  }
}


// ---- mass path of the 9-node shells: gamma M into a matrix and / or the inertial residual M u'' ----
// (TACS_MASS_MATRIX, the gamma term of assembleJacobian, TACSShellElement.h:410-447, 614-648).
// One warp per element: node normals, w det of the 9 Gauss points, then the 81 node-pair blocks
// (a scalar per pair times fixed 3 x 3 patterns, q9_mass_pair); p.u holds the second time
// derivatives (as for k_mass), p.alpha the scale of the matrix.
struct Mass9Warp {
  a2ds::Elem9 E;
  int nodes[9];
  int off[81];
};
template <bool RES, bool MAT>
__global__ void __launch_bounds__(128) k_mass9(KParams p, const a2ds::Shape9 *Hg) {
  using namespace a2ds;
  extern __shared__ __align__(16) unsigned char smem9m[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Mass9Warp &W = reinterpret_cast<Mass9Warp *>(smem9m)[warp];
  const Shape9 &H = *Hg;
  const unsigned FULL = 0xffffffffu;
  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(p.work_counter, 1);
    idx = __shfl_sync(FULL, idx, 0);
    if (idx >= p.n_list) break;
    const int e = p.elem_list ? __ldg(&p.elem_list[idx]) : idx;
    const CompData &c = p.comps[__ldg(&p.elem_comp[e])];
    if (lane < 9) W.nodes[lane] = __ldg(&p.conn[9 * (size_t)e + lane]);
    if (MAT)
      for (int k = lane; k < 81; k += 32) W.off[k] = __ldg(&p.Koff[81 * (size_t)e + k]);
    __syncwarp();
    if (lane < 27) W.E.X[lane] = __ldg(&p.X[3 * (size_t)W.nodes[lane / 3] + lane % 3]);
    if (RES)
      for (int d = lane; d < Q9_NV; d += 32) W.E.q[d] = __ldg(&p.u[6 * (size_t)W.nodes[d / 6] + d % 6]);
    __syncwarp();
    if (lane < 9) q9_node_normal(W.E, lane);
    __syncwarp();
    if (lane < 9) q9_qp_det(W.E, H, lane);
    __syncwarp();
    for (int pair = lane; pair < 81; pair += 32) {
      const int ma = pair / 9, mb = pair - 9 * ma;
      double blk[36];
      q9_mass_pair(c, W.E, H, ma, mb, blk);
      if (MAT) {
        double *dst = p.Kval + 36 * (size_t)W.off[pair];
#pragma unroll
        for (int k = 0; k < 36; k++)
          if (blk[k] != 0.0) atomicAdd(dst + k, p.alpha * blk[k]);
      }
      if (RES) {
        double *r = &p.res[6 * (size_t)W.nodes[ma]];
#pragma unroll
        for (int i = 0; i < 6; i++) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < 6; j++) s += blk[6 * i + j] * W.E.q[6 * mb + j];
          atomicAdd(r + i, p.res_scale * s);
        }
      }
    }
    __syncwarp();
  }
}
// fills the shape tables once per device (k_mass9 reads them from global memory)
__global__ void k_shape9_tables(a2ds::Shape9 *H) {
  if (threadIdx.x < 46) a2ds::q9_shape_tables(*H, threadIdx.x);
}

#endif
