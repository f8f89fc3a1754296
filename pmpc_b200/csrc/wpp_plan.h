// Launch plan of the warp-per-particle kernel (wpp_kernels.cuh), shared with the host dispatch in api.cu.
#pragma once
#include <stddef.h>

namespace pmpc {
// W warps per CTA, Tc CTAs per problem (team), ppw particles per warp.  A team of more than one CTA exchanges through
// global-memory flags, so all CTAs must be co-resident: the kernel is then launched cooperatively and the grid is kept
// within the resident capacity by giving every warp more particles.
struct WppPlan {
  int W, Tc, ppw, blocks, threads, cap, cooperative;
  size_t smem, tm_doubles, wscr_doubles;
};
}  // namespace pmpc
