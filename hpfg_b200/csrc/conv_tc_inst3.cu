// Explicit instantiations of the tensor-core convolution kernels for KC=32, BN=16 (3x3 and 1x1); see conv_tc_kernel.cuh.
#include "conv_tc_kernel.cuh"
namespace hpfg {
template int tc_launch<3, 32, 16>(int, int, int, const CUtensorMap &, const CUtensorMap &, const TcConvParams &, cudaStream_t);
template int tc_launch<1, 32, 16>(int, int, int, const CUtensorMap &, const CUtensorMap &, const TcConvParams &, cudaStream_t);
}  // namespace hpfg
