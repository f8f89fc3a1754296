// Kernel entry points and the per-dimension launch template (one translation unit per (NX, NU)).
#pragma once
#include "solver.cuh"

namespace pmpc {

// One CTA per problem, one thread per particle (GROUPED) or one thread per single-particle
// problem (!GROUPED).  BLK bounds the CTA size so small-M problems may use the full register file.
template <int NX, int NU, bool GROUPED, int BLK>
__global__ void __launch_bounds__(BLK) scp_kernel(const __grid_constant__ DevProb pb) {
  __shared__ double red[ST_COUNT];
  Solver<NX, NU, GROUPED> S(pb, red);
  S.run();
}

// Launches the SCP / QP kernel for compile-time dims; defined by explicit instantiation in
// inst_<NX>_<NU>.cu.  Returns the cudaError_t of the launch.
template <int NX, int NU>
cudaError_t launch_scp(const DevProb& pb, cudaStream_t stream);

template <int NX, int NU>
cudaError_t launch_scp_impl(const DevProb& pb, cudaStream_t stream) {
  if (pb.M == 1) {
    const int threads = 128;
    const int blocks = (pb.B + threads - 1) / threads;
    scp_kernel<NX, NU, false, 128><<<blocks, threads, 0, stream>>>(pb);
  } else {
    const int threads = ((pb.M + 31) / 32) * 32;
    if (threads <= 128)
      scp_kernel<NX, NU, true, 128><<<pb.B, threads, 0, stream>>>(pb);
    else
      scp_kernel<NX, NU, true, 1024><<<pb.B, threads, 0, stream>>>(pb);
  }
  return cudaGetLastError();
}

}  // namespace pmpc
