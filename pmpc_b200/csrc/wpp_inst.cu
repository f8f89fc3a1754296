// One instantiation of the warp-per-particle kernel per translation unit:
//   -DPMPC_NX=.. -DPMPC_NU=.. -DPMPC_SLEW=0|1
#include <algorithm>

#include "wpp_kernels.cuh"

#define PMPC_CAT_(a, b, c, d) a##_##b##_##c##_##d
#define PMPC_CAT(a, b, c, d) PMPC_CAT_(a, b, c, d)

namespace pmpc {
cudaError_t PMPC_CAT(wpp_plan, PMPC_NX, PMPC_NU, PMPC_SLEW)(const DevProb& pb, WppPlan* pl) {
  return wpp_plan<PMPC_NX, PMPC_NU, (PMPC_SLEW != 0)>(pb, pl);
}
cudaError_t PMPC_CAT(wpp_launch, PMPC_NX, PMPC_NU, PMPC_SLEW)(const DevProb& pb, const WppPlan& pl, cudaStream_t stream) {
  return wpp_launch<PMPC_NX, PMPC_NU, (PMPC_SLEW != 0)>(pb, pl, stream);
}
}  // namespace pmpc
