// Kernel entry point and launch helper of the warp-per-particle kernel (wpp.cuh); one translation unit per
// (NX, NU, SLEW) instantiation (wpp_inst.cu).
#pragma once
#include <stdlib.h>

#include "wpp.cuh"
#include "wpp_plan.h"

namespace pmpc {

// warps per CTA: bounded by the per-warp shared-memory slab (n = 32: 28.5 KB per warp) and by 255 registers per thread
template <int NX>
struct WppCfg {
  static constexpr int WMAX = (NX >= 32) ? 7 : 8;
};

template <int NX, int NU, bool SLEW>
__global__ void __launch_bounds__(32 * WppCfg<NX>::WMAX, 1) wpp_kernel(const __grid_constant__ DevProb pb, int Tc, int ppw) {
  __shared__ double red[ST_COUNT];
  __shared__ double red2[RCAP];
  Solver<NX, NU, true> S0(pb, red, red2, (int)blockIdx.x / Tc);
  Wpp<NX, NU, SLEW> wk(pb, wpp_smem);
  wk.run(S0, Tc, ppw);
}

template <int NX, int NU, bool SLEW>
cudaError_t wpp_plan(const DevProb& pb, WppPlan* pl) {
  using K = Wpp<NX, NU, SLEW>;
  const int M = pb.M, B = pb.B;
  const int D = (M == 1) ? 0 : pb.D;
  int dev0 = 0, sms0 = 148;
  cudaGetDevice(&dev0);
  cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
  // Warps per CTA: a warp's sweep is a dependent chain, and warps that share an SM slow each other down (C5 with 128
  // particles: 175 ms at 7 warps per SM, 131 ms at 2; C4: 21.9 ms at 8, 18.0 ms at 1-4), so use only as many warps per
  // CTA as it takes to give every particle of the launch a warp: ceil(B M / SMs), at least 2, at most WMAX.
  const long long need = ((long long)B * M + sms0 - 1) / sms0;
  pl->W = (int)std::min<long long>(WppCfg<NX>::WMAX, std::max<long long>(2, need));
  pl->W = std::min(pl->W, M);
  if (const char* e = getenv("PMPC_WPP_W")) {  // experiments: warps per CTA (fewer warps = more SMs per problem)
    const int w = atoi(e);
    if (w >= 1) pl->W = std::min(pl->W, w);
  }
  pl->threads = 32 * pl->W;
  pl->smem = ((size_t)pl->W * K::SLAB + K::cred_doubles(pl->W)) * sizeof(double);
  auto kern = wpp_kernel<NX, NU, SLEW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem);
  if (e != cudaSuccess) return e;
  int occ = 0, dev = 0, sms = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, pl->threads, pl->smem)) != cudaSuccess) return e;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long resident = (long long)std::max(occ, 1) * sms;
  const int full = (M + pl->W - 1) / pl->W;  // CTAs per problem at one particle per warp
  if (full <= 1) {
    pl->Tc = 1;
    pl->ppw = 1;
  } else if ((long long)B * full <= resident) {
    pl->Tc = full;
    pl->ppw = 1;
  } else {
    const int tc = (int)std::max<long long>(1, resident / B);
    pl->ppw = (M + pl->W * tc - 1) / (pl->W * tc);
    pl->Tc = (M + pl->W * pl->ppw - 1) / (pl->W * pl->ppw);
  }
  pl->blocks = B * pl->Tc;
  pl->cooperative = pl->Tc > 1 ? 1 : 0;
  pl->cap = std::max(D * D + D, (int)ST_COUNT) + 8;
  pl->tm_doubles = (size_t)B * ((size_t)(pl->Tc + 1) * pl->cap + (size_t)pl->Tc + 1);
  pl->wscr_doubles = ((size_t)pl->blocks * pl->W + pl->blocks) * (size_t)pl->cap;
  return cudaSuccess;
}

template <int NX, int NU, bool SLEW>
cudaError_t wpp_launch(const DevProb& pb, const WppPlan& pl, cudaStream_t stream) {
  auto kern = wpp_kernel<NX, NU, SLEW>;
  if (pl.cooperative) {
    DevProb pbc = pb;
    int Tc = pl.Tc, ppw = pl.ppw;
    void* args[] = {&pbc, &Tc, &ppw};
    return cudaLaunchCooperativeKernel((const void*)kern, dim3(pl.blocks), dim3(pl.threads), args, pl.smem, stream);
  }
  kern<<<pl.blocks, pl.threads, pl.smem, stream>>>(pb, pl.Tc, pl.ppw);
  return cudaGetLastError();
}

}  // namespace pmpc
