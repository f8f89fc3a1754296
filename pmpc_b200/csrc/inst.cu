// One kernel instantiation per translation unit (so `make -j` compiles them in parallel):
//   -DPMPC_NX=.. -DPMPC_NU=..   compile-time state / control dimensions
//   -DPMPC_KIND=0|1|2           0: one thread per single-particle problem (128-thread CTAs)
//                               1: one CTA per problem, M <= 128 particles   2: M <= 1024 particles
//                               3: two problems of M <= 16 particles per warp (fused active-set passes only)
//   -DPMPC_VAR=0..3             fused-path variant: bit 0 on-the-fly Dubins, bit 1 lean (no slew)
#include "kernels.cuh"

#define PMPC_CAT_(a, b, c, d) launch_##a##_##b##_##c##_##d
#define PMPC_CAT(a, b, c, d) PMPC_CAT_(a, b, c, d)

namespace pmpc {
cudaError_t PMPC_CAT(PMPC_NX, PMPC_NU, PMPC_KIND, PMPC_VAR)(const DevProb& pb, int blocks, int threads,
                                                            cudaStream_t stream, int* occ_out) {
#if PMPC_KIND == 3
  return launch_half<PMPC_NX, PMPC_NU, PMPC_VAR>(pb, blocks, threads, stream, occ_out);
#else
  return launch_variant<PMPC_NX, PMPC_NU, (PMPC_KIND != 0), (PMPC_KIND == 2 ? 1024 : 128), PMPC_VAR>(
      pb, blocks, threads, stream, occ_out);
#endif
}
}  // namespace pmpc
