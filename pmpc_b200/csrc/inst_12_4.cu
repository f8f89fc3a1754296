// Explicit instantiation of the SCP kernels for NX=12, NU=4 (generated layout, one TU per shape
// so the shapes compile in parallel).
#include "kernels.cuh"
namespace pmpc {
template <>
cudaError_t launch_scp<12, 4>(const DevProb& pb, cudaStream_t stream) { return launch_scp_impl<12, 4>(pb, stream); }
}  // namespace pmpc
