// Explicit instantiation of the SCP kernels for NX=8, NU=4 (generated layout, one TU per shape
// so the shapes compile in parallel).
#include "kernels.cuh"
namespace pmpc {
template <>
cudaError_t launch_scp<8, 4>(const DevProb& pb, cudaStream_t stream) { return launch_scp_impl<8, 4>(pb, stream); }
}  // namespace pmpc
