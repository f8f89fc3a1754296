// Kernel entry points and the per-dimension launch template (one translation unit per (NX, NU)).
#pragma once
#include "solver.cuh"

namespace pmpc {

// One CTA per problem, one thread per particle (GROUPED) or one thread per single-particle
// problem (!GROUPED).  BLK bounds the CTA size so small-M problems may use the full register file.
template <int NX, int NU, bool GROUPED, int BLK>
__global__ void __launch_bounds__(BLK, (BLK <= 128 ? 3 : 1)) scp_kernel(const __grid_constant__ DevProb pb) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ double red[ST_COUNT];
  Solver<NX, NU, GROUPED> S(pb, red);
  S.run(dyn_smem);
}

// Launches the SCP / QP kernel for compile-time dims; defined by explicit instantiation in
// inst_<NX>_<NU>.cu.  Returns the cudaError_t of the launch.
template <int NX, int NU>
cudaError_t launch_scp(const DevProb& pb, cudaStream_t stream);

template <int NX, int NU, bool GROUPED, int BLK>
cudaError_t launch_variant(DevProb pb, int blocks, int threads, cudaStream_t stream) {
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
  auto kern = scp_kernel<NX, NU, GROUPED, BLK>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<blocks, threads, smem, stream>>>(pb);
  return cudaGetLastError();
}

template <int NX, int NU>
cudaError_t launch_scp_impl(const DevProb& pb, cudaStream_t stream) {
  if (pb.M == 1) {
    const int threads = 128;
    return launch_variant<NX, NU, false, 128>(pb, (pb.B + threads - 1) / threads, threads, stream);
  }
  const int threads = ((pb.M + 31) / 32) * 32;
  if (threads <= 128) return launch_variant<NX, NU, true, 128>(pb, pb.B, threads, stream);
  return launch_variant<NX, NU, true, 1024>(pb, pb.B, threads, stream);
}

}  // namespace pmpc
