// Kernel entry points and the per-dimension launch template (one translation unit per (NX, NU)).
#pragma once
#include "solver.cuh"

#ifndef PMPC_MINB
#define PMPC_MINB 2
#endif
#include <cstdlib>
namespace pmpc {

// One CTA per problem, one thread per particle (GROUPED) or one thread per single-particle
// problem (!GROUPED).  BLK bounds the CTA size so small-M problems may use the full register file.
template <int NX, int NU, bool GROUPED, int BLK, int VAR>
__global__ void __launch_bounds__(BLK, (BLK <= 128 ? PMPC_MINB : 1)) scp_kernel(const __grid_constant__ DevProb pb) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ double red[ST_COUNT];
  __shared__ double red2[(BLK / 32) * RCAP];  // per-warp partials of the small CTA reductions (Solver::cta_reduce)
  Solver<NX, NU, GROUPED> S(pb, red, red2);
  S.template run<VAR>(pb, dyn_smem);
}

static __global__ void queue_init_kernel(int* q, int v) { *q = v; }

// Two problems of at most 16 particles per warp (Solver PPW = 2), HW warps per CTA whose sweeps are phase-aligned
// (Solver::half_phase).
template <int NX, int NU, int VAR>
__global__ void __launch_bounds__(256, 1) scp_kernel_half(const __grid_constant__ DevProb pb) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ double red[ST_COUNT];
  __shared__ double red2[RCAP];
  Solver<NX, NU, true, 2> S(pb, red, red2);
  // persistent CTAs: the first pair of a warp is its own index, further pairs come from the queue (starts at grid x warps)
  const int W = (int)(blockDim.x >> 5), npairs = (pb.B + 1) / 2;
  int pair = (int)blockIdx.x * W + (int)(threadIdx.x >> 5);
  for (bool first = true; pair < npairs; first = false) {
    if (!first) S.assign_pair(pair);
    S.template run_half<VAR>(pb, dyn_smem, first);
    pair = ((threadIdx.x & 31) == 0) ? atomicAdd(pb.queue, 1) : 0;
    pair = __shfl_sync(0xffffffffu, pair, 0);
  }
  S.half_idle();
}

template <int NX, int NU, int VAR>
cudaError_t launch_half(DevProb pb, int blocks, int threads, cudaStream_t stream, int* occ_out) {
  using S = Solver<NX, NU, true, 2>;
  (void)blocks; (void)threads;
  auto kern = scp_kernel_half<NX, NU, VAR>;
  const int one = (S::table_doubles(pb.N) + 15) / 16 * 16;  // one uniform cost table per half-warp problem
  const int want_ring = pb.ring;
  // warps per CTA: as many as keep 8 warps resident per SM (the register file's limit) — 4 (two CTAs per SM), else 2, 1;
  // the cost tables are dropped (global reads instead) when they do not fit beside the ring at that occupancy
  const char* we = getenv("PMPC_HALF_W");
  int W = we ? atoi(we) : 4, occ = 0;
  if (W != 8 && W != 4 && W != 2 && W != 1) W = 4;
  size_t smem = 0;
  for (;; W >>= 1) {
    const size_t cap = (227 * 1024 - 2048) / (size_t)max(1, 8 / W) - 1024;
    size_t ring_bytes = (size_t)S::ring_doubles(W) * 8 + (size_t)W * S::DEPTH * 8;
    pb.ring = (S::RING_SHAPE && want_ring) ? 1 : 0;
    if (!pb.ring) ring_bytes = 0;
    pb.tab_doubles = ((size_t)2 * W * one * 8 + ring_bytes <= cap) ? 2 * W * one : 0;
    smem = (size_t)pb.tab_doubles * 8 + ring_bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W, smem)) != cudaSuccess) return e;
    if (occ * W >= 8 || W == 1) break;
  }
  if (occ_out != nullptr) {
    *occ_out = occ * W;  // warps (= pairs of problems) per SM
    return cudaSuccess;
  }
  int dev = 0, sms = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int nblk = min((pb.B + 2 * W - 1) / (2 * W), occ * sms);  // one wave of resident CTAs; the rest comes through the queue
  queue_init_kernel<<<1, 1, 0, stream>>>(pb.queue, nblk * W);
  kern<<<nblk, 32 * W, smem, stream>>>(pb);
  return cudaGetLastError();
}

template <int NX, int NU, bool GROUPED, int BLK, int VAR>
cudaError_t launch_variant(DevProb pb, int blocks, int threads, cudaStream_t stream, int* occ_out) {
  using S = Solver<NX, NU, GROUPED>;
  // fused-path shared memory: [uniform cost table][per-warp ring][per-warp mbarriers]
  const int nw = threads / 32;
  const size_t max_smem = 200 * 1024;
  // table rounded to 128 bytes so that the ring behind it keeps the 16-byte alignment bulk copies need
  pb.tab_doubles = (GROUPED && (size_t)S::table_doubles(pb.N) * 8 <= 32 * 1024) ? (S::table_doubles(pb.N) + 15) / 16 * 16 : 0;
  size_t ring_bytes = (size_t)S::ring_doubles(nw) * 8 + (size_t)nw * S::DEPTH * 8;
  pb.ring = (GROUPED && S::RING_SHAPE && pb.ring && (size_t)pb.tab_doubles * 8 + ring_bytes <= max_smem) ? 1 : 0;
  if (!pb.ring) ring_bytes = 0;
  const size_t smem = (size_t)pb.tab_doubles * 8 + ring_bytes;
  auto kern = scp_kernel<NX, NU, GROUPED, BLK, VAR>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (occ_out != nullptr)  // query only: resident CTAs per SM for this launch configuration
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_out, kern, threads, smem);
  kern<<<blocks, threads, smem, stream>>>(pb);
  return cudaGetLastError();
}

}  // namespace pmpc
