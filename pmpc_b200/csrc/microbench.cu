// Roofline denominators measured on the device at hand (pmpc_microbench, include/pmpc_b200.h):
// a STREAM-style copy for HBM bandwidth and register-resident DFMA chains for the fp64 pipe.
#include <cuda_runtime.h>

#include "../../include/pmpc_b200.h"

namespace {

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

// 8 independent FMA chains per thread: enough ILP to cover the DFMA latency at 8+ warps per SMSP.
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// fp64 tensor-core path (DMMA): 4 independent m8n8k4 accumulator chains per warp, 256 FMAs per instruction
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c0[2] = {0.0, 0.0}, c1[2] = {1.0, 0.0}, c2[2] = {2.0, 0.0}, c3[2] = {3.0, 0.0};
  const double av = a + 1e-12 * threadIdx.x, bv = b;
  for (int i = 0; i < iters; ++i) {
#define PMPC_DMMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" \
                                  : "+d"(c[0]), "+d"(c[1]) : "d"(av), "d"(bv))
    PMPC_DMMA(c0); PMPC_DMMA(c1); PMPC_DMMA(c2); PMPC_DMMA(c3);
#undef PMPC_DMMA
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = (c0[0] + c0[1]) + (c1[0] + c1[1]) + (c2[0] + c2[1]) + (c3[0] + c3[1]);
}

// DFMA fed from shared memory: WIDE = 1: one broadcast LDS.64 per FMA; WIDE = 2: one broadcast LDS.128 per two FMAs
// (the operand pattern of a warp-per-particle matrix product with one operand in shared memory)
template <int WIDE>
__global__ void dfma_lds_kernel(double* out, int iters) {
  __shared__ __align__(16) double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 / (1.0 + i);
  __syncthreads();
  double x[8];
  for (int k = 0; k < 8; ++k) x[k] = threadIdx.x + k;
  const double a = 0.999999 + 1e-9 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    const int base = (i & 63) * 16;
    if (WIDE == 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = fma(sm[base + k], a, x[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const double2 v = *reinterpret_cast<const double2*>(&sm[base + k]);
        x[k] = fma(v.x, a, x[k]);
        x[k + 1] = fma(v.y, a, x[k + 1]);
      }
    }
  }
  double s = 0.0;
  for (int k = 0; k < 8; ++k) s += x[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int pmpc_microbench(int device, int kind, double* result) {
  if (!result) return PMPC_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return PMPC_ERR_NO_DEVICE; }
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return PMPC_ERR_CUDA;
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  int rc = PMPC_OK;
  if (kind == 0) {
    const size_t bytes = (size_t)1 << 30, n = bytes / sizeof(double2);
    double2 *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) { cudaFree(a); return PMPC_ERR_CUDA; }
    cudaMemset(a, 1, bytes);
    for (int rep = 0; rep < 8; ++rep) {
      cudaEventRecord(e0);
      copy_kernel<<<prop.multiProcessorCount * 16, 512>>>(a, b, n);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
    }
    *result = 2.0 * (double)bytes / (best * 1e-3) / 1e9;
    cudaFree(a);
    cudaFree(b);
  } else if (kind == 1) {
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 1 << 14;
    double* out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return PMPC_ERR_CUDA;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
    }
    *result = 2.0 * 8.0 * (double)iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    cudaFree(out);
  } else if (kind >= 2 && kind <= 4) {
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 1 << 14;
    double* out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return PMPC_ERR_CUDA;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      if (kind == 2) dmma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
      else if (kind == 3) dfma_lds_kernel<1><<<blocks, threads>>>(out, iters);
      else dfma_lds_kernel<2><<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
    }
    // kind 2: 4 DMMA x 256 FMAs per warp per iteration; kinds 3, 4: 8 FMAs per thread per iteration
    const double fl = (kind == 2) ? 2.0 * 4.0 * 256.0 * (double)iters * (double)blocks * (threads / 32)
                                  : 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    *result = fl / (best * 1e-3) / 1e12;
    cudaFree(out);
  } else {
    rc = PMPC_ERR_BAD_ARG;
  }
  if (cudaGetLastError() != cudaSuccess) rc = PMPC_ERR_CUDA;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}
