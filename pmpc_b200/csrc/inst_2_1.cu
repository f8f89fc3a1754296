// Explicit instantiation of the SCP kernels for NX=2, NU=1 (generated layout, one TU per shape
// so the shapes compile in parallel).
#include "kernels.cuh"
namespace pmpc {
template <>
cudaError_t launch_scp<2, 1>(const DevProb& pb, cudaStream_t stream) { return launch_scp_impl<2, 1>(pb, stream); }
}  // namespace pmpc
