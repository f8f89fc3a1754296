// Explicit instantiation of the SCP kernels for NX=4, NU=2 (generated layout, one TU per shape
// so the shapes compile in parallel).
#include "kernels.cuh"
namespace pmpc {
template <>
cudaError_t launch_scp<4, 2>(const DevProb& pb, cudaStream_t stream) { return launch_scp_impl<4, 2>(pb, stream); }
}  // namespace pmpc
