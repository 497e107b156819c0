// FP64 pipe microbenchmarks for the roofline denominators (B200, sm_100a):
//   dfma  : dependent-chain-free DFMA throughput (vector FP64 pipe)
//   dmma  : mma.sync.m8n8k4.f64 throughput (FP64 tensor path)
//   mixed : both interleaved, to see whether the pipes overlap
//   red   : FP64 atomicAdd (RED.E.ADD.F64) throughput into an L2-resident array
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double *out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_mixed(double *out, int iters, double a, double b) {
  double c0[ILP], c1[ILP], acc[2 * ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x; c1[i] = i; acc[2 * i] = i; acc[2 * i + 1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      dmma884(c0[i], c1[i], a, b);
      acc[2 * i] = fma(acc[2 * i], a, b);
      acc[2 * i + 1] = fma(acc[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i] + acc[2 * i] + acc[2 * i + 1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// pattern 0: each thread adds to 36 consecutive doubles of "its" block (block stride 288 B);
// neighbouring threads hit neighbouring blocks -> like one-thread-per-block scatter.
// pattern 1: a warp adds 32 consecutive doubles (fully coalesced RED).
__global__ void k_red(double *buf, size_t nblocks, int pattern, int reps) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t nthreads = gridDim.x * (size_t)blockDim.x;
  if (pattern == 0) {
    for (int r = 0; r < reps; r++)
      for (size_t b = tid; b < nblocks; b += nthreads)
        for (int k = 0; k < 36; k++) atomicAdd(&buf[36 * b + k], 1.0);
  } else {
    size_t n = 36 * nblocks;
    for (int r = 0; r < reps; r++)
      for (size_t i = tid; i < n; i += nthreads) atomicAdd(&buf[i], 1.0);
  }
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s  SMs %d  clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int nsm = p.multiProcessorCount;
  double *out;
  CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int iters = 20000;
  for (int threads = 128; threads <= 1024; threads *= 2) {
    int blocks = nsm * (2048 / threads);
    k_dfma<8><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * iters * (double)blocks * threads;
    printf("dfma  threads/blk %4d  %.3f ms  %.2f TFLOP/s\n", threads, ms, flops / ms * 1e-9);
  }
  for (int wps = 4; wps <= 64; wps *= 2) {  // warps per SM
    int threads = 128, blocks = nsm * wps * 32 / threads;
    k_dmma<8><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * 8 * iters * (double)blocks * threads / 32;
    printf("dmma  warps/SM %3d  %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
  }
  for (int wps = 8; wps <= 64; wps *= 2) {
    int threads = 128, blocks = nsm * wps * 32 / threads;
    k_mixed<4><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k_mixed<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    double f_mma = 2.0 * 256 * 4 * iters * (double)blocks * threads / 32;
    double f_fma = 2.0 * 8 * iters * (double)blocks * threads;
    printf("mixed warps/SM %3d  %.3f ms  dmma %.2f + dfma %.2f = %.2f TFLOP/s\n", wps, ms,
           f_mma / ms * 1e-9, f_fma / ms * 1e-9, (f_mma + f_fma) / ms * 1e-9);
  }
  // RED.F64 throughput: L2-resident (64 MB) and HBM-resident (4 GB) targets
  for (int big = 0; big < 2; big++) {
    size_t nblocks = big ? (size_t)14 * 1000 * 1000 : (size_t)220 * 1000;
    double *buf;
    CK(cudaMalloc(&buf, nblocks * 288));
    CK(cudaMemset(buf, 0, nblocks * 288));
    for (int pattern = 0; pattern < 2; pattern++) {
      int reps = big ? 1 : 20;
      k_red<<<nsm * 8, 256>>>(buf, nblocks, pattern, 1);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      k_red<<<nsm * 8, 256>>>(buf, nblocks, pattern, reps);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      double n = 36.0 * nblocks * reps;
      printf("red.f64 %s pattern %d: %.3f ms  %.2f G atomics/s  %.1f GB/s (8 B each)\n",
             big ? "4GB " : "63MB", pattern, ms, n / ms * 1e-6, 8 * n / ms * 1e-6);
    }
    cudaFree(buf);
  }
  return 0;
}
