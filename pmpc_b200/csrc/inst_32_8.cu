// Explicit instantiation of the SCP kernels for NX=32, NU=8 (generated layout, one TU per shape
// so the shapes compile in parallel).
#include "kernels.cuh"
namespace pmpc {
template <>
cudaError_t launch_scp<32, 8>(const DevProb& pb, cudaStream_t stream) { return launch_scp_impl<32, 8>(pb, stream); }
}  // namespace pmpc
