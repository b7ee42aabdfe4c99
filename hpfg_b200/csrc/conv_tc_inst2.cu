// Explicit instantiations of the tensor-core convolution kernels for KC=16, BN=128 (3x3 and 1x1); see conv_tc_kernel.cuh.
#include "conv_tc_kernel.cuh"
namespace hpfg {
template int tc_launch<3, 16, 128>(int, int, int, const CUtensorMap &, const CUtensorMap &, const TcConvParams &, cudaStream_t);
template int tc_launch<1, 16, 128>(int, int, int, const CUtensorMap &, const CUtensorMap &, const TcConvParams &, cudaStream_t);
}  // namespace hpfg
