// assemble9_kernels.cuh — k_assemble9<RES, KMAT>: residual and tangent of the 9-node MITC shell
// (TACSQuad9Shell, linear strain model) assembled into the same 6 x 6 BCSR matrices and node
// vectors as the 4-node path.  Replaces, for that element class, the element loop of
// TACSAssembler::assembleRes / assembleJacobian / assembleMatType(STIFFNESS)
// (src/TACSAssembler.cpp:4000-4242) with TACSShellElement::addResidual / addJacobian inside
// (src/elements/shell/TACSShellElement.h:303-672).
//
// One thread block of 7 warps per element, elements drawn from a counter.  The element math
// (mitc9_math.h) runs as block-level phases over shared memory: 9 node frames, 28 tying points +
// 9 Gauss-point frames, the derivative tables (28 x 54 tying, 9 x 54 drill), then per row of
// three Gauss points the 3 x 54 columns of B and w det C B.  The tangent K = sum B^T (w det C B)
// is contracted on the FP64 tensor path (mma.m8n8k4): warp w owns the 8 rows 8 w .. 8 w + 7 of
// the 54 x 54 (padded to 56) element matrix and accumulates the tiles on and right of the
// diagonal (28 of 49, the matrix is symmetric); the accumulators are added to the BCSR blocks
// straight from registers, mirrored for the off-diagonal tiles — there is no staged element matrix.
#ifndef A2DS_ASSEMBLE9_KERNELS_CUH
#define A2DS_ASSEMBLE9_KERNELS_CUH

#include "mitc9_math.h"

static const int Q9_THREADS = 224;   // 7 warps: one per 8-row tile of the 56 x 56 padded matrix
static const int Q9_QB = 3;          // Gauss points per contraction batch (one eta row)

struct Elem9Block {
  a2ds::Elem9 E;
  double B2[Q9_QB - 1][a2ds::Q9_KROWS][a2ds::Q9_LD];    // B / CB of the other points of the batch
  double CB2[Q9_QB - 1][a2ds::Q9_KROWS][a2ds::Q9_LD];
  int nodes[9];
  int off[81];
  int elem;
};

template <bool RES, bool KMAT>
__global__ void __launch_bounds__(Q9_THREADS, 2) k_assemble9(KParams p) {
  using namespace a2ds;
  extern __shared__ __align__(16) unsigned char smem9[];
  Elem9Block &S = *reinterpret_cast<Elem9Block *>(smem9);
  Elem9 &E = S.E;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // batch slot qq: tables of Gauss point 3 b + qq of the current batch b
  auto Bq = [&](int qq) -> double(*)[Q9_LD] { return qq == 0 ? E.B : S.B2[qq - 1]; };
  auto CBq = [&](int qq) -> double(*)[Q9_LD] { return qq == 0 ? E.CB : S.CB2[qq - 1]; };
  // rows 9..11 and columns 54, 55 of the tables stay zero
  for (int i = tid; i < Q9_KROWS * Q9_LD; i += Q9_THREADS) {
    (&E.B[0][0])[i] = 0.0; (&E.CB[0][0])[i] = 0.0;
    for (int qq = 0; qq < Q9_QB - 1; qq++) { (&S.B2[qq][0][0])[i] = 0.0; (&S.CB2[qq][0][0])[i] = 0.0; }
  }
  for (;;) {
    if (tid == 0) S.elem = atomicAdd(p.work_counter, 1);
    __syncthreads();
    const int idx = S.elem;
    if (idx >= p.n_list) break;
    const int e = p.elem_list ? __ldg(&p.elem_list[idx]) : idx;
    const CompData &c = p.comps[__ldg(&p.elem_comp[e])];
    if (tid < 9) S.nodes[tid] = __ldg(&p.conn[9 * (size_t)e + tid]);
    if (KMAT && tid >= 32 && tid < 32 + 81) S.off[tid - 32] = __ldg(&p.Koff[81 * (size_t)e + tid - 32]);
    __syncthreads();
    if (tid < 27) E.X[tid] = __ldg(&p.X[3 * (size_t)S.nodes[tid / 3] + tid % 3]);
    else if (tid >= 32 && tid < 32 + 54) {
      const int d = tid - 32;
      E.q[d] = __ldg(&p.u[6 * (size_t)S.nodes[d / 6] + d % 6]);
    }
    __syncthreads();
    if (tid < 9) q9_node(c, E, tid);
    __syncthreads();
    if (tid < 28) q9_tying(E, tid);
    else if (tid >= 32 && tid < 41) q9_qp(c, E, tid - 32);
    __syncthreads();
    for (int i = tid; i < Q9_NTY * Q9_NV; i += Q9_THREADS) E.Gt[i / Q9_NV][i % Q9_NV] = q9_gt(E, i / Q9_NV, i % Q9_NV);
    for (int i = tid; i < Q9_NN * Q9_NV; i += Q9_THREADS) E.Dn[i / Q9_NV][i % Q9_NV] = q9_dn(E, i / Q9_NV, i % Q9_NV);
    if (RES && tid >= 192 && tid < 201) q9_qp_state(c, E, tid - 192, p.thermal);
    __syncthreads();

    double acc[7][2];
#pragma unroll
    for (int t = 0; t < 7; t++) acc[t][0] = acc[t][1] = 0.0;
    double r = 0.0;
    for (int b = 0; b < 3; b++) {
      if (tid < Q9_QB * Q9_NV) {
        const int qq = tid / Q9_NV, col = tid - Q9_NV * qq, q = Q9_QB * b + qq;
        double Bk[9], Sk[9];
        q9_bcol(E, q, col, Bk);
        q9_stress(c.Cs, Bk, Sk);
        double(*Bt)[Q9_LD] = Bq(qq), (*Ct)[Q9_LD] = CBq(qq);
        const double w = E.qw[q];
#pragma unroll
        for (int k = 0; k < 9; k++) { Bt[k][col] = Bk[k]; Ct[k][col] = w * Sk[k]; }
      }
      __syncthreads();
      if (KMAT) {
#pragma unroll
        for (int qq = 0; qq < Q9_QB; qq++) {
          const double(*Bt)[Q9_LD] = Bq(qq), (*Ct)[Q9_LD] = CBq(qq);
#pragma unroll
          for (int ks = 0; ks < 3; ks++) {
            const int kr = 4 * ks + (lane & 3);
            const double a = Bt[kr][8 * warp + (lane >> 2)];
#pragma unroll
            for (int tj = 0; tj < 7; tj++)
              if (tj >= warp) dmma884(acc[tj], a, Ct[kr][8 * tj + (lane >> 2)]);
          }
        }
      }
      if (RES && tid < Q9_NV) {
#pragma unroll
        for (int qq = 0; qq < Q9_QB; qq++) {
          const double(*Bt)[Q9_LD] = Bq(qq);
          const double *s = E.sq[Q9_QB * b + qq];
#pragma unroll
          for (int k = 0; k < 9; k++) r += Bt[k][tid] * s[k];
        }
      }
      __syncthreads();
    }
    if (RES && tid < Q9_NV)
      atomicAdd(&p.res[6 * (size_t)S.nodes[tid / 6] + tid % 6], p.res_scale * r);
    if (KMAT) {
      // accumulator (row, col) of tile (warp, tj): rows 8 warp + lane / 4, columns 8 tj + 2 (lane % 4) + i
      const int gr = 8 * warp + (lane >> 2);
      if (gr < Q9_NV) {
        const int br = gr / 6, rr = gr - 6 * br;
#pragma unroll
        for (int tj = 0; tj < 7; tj++) {
          if (tj < warp) continue;
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const int gc = 8 * tj + 2 * (lane & 3) + i;
            if (gc >= Q9_NV) continue;
            const int bc = gc / 6, cc = gc - 6 * bc;
            const double v = p.alpha * acc[tj][i];
            atomicAdd(p.Kval + 36 * (size_t)S.off[9 * br + bc] + 6 * rr + cc, v);
            if (tj > warp) atomicAdd(p.Kval + 36 * (size_t)S.off[9 * bc + br] + 6 * cc + rr, v);
          }
        }
      }
    }
    // S.elem, S.nodes and S.off are rewritten only after the barriers at the top of the next trip
  }
}

#endif
