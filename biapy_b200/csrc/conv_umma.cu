// tcgen05 / TMA implicit-GEMM convolution kernels (placeholder until the kernels land in this file).
#include "common.cuh"

namespace b200 {
bool conv_fprop_umma_supported(const b200_tensor*, const b200_tensor*, const b200_tensor*, int, int, int) { return false; }
int conv_fprop_umma(const b200_tensor*, const void*, const float*, const b200_tensor*, const b200_tensor*, int, int, int, int,
                    cudaStream_t) {
  set_error("conv_fprop_umma: not built");
  return B200_ERR_UNSUPPORTED;
}
bool conv_wgrad_umma_supported(const b200_tensor*, const b200_tensor*, int, int, int) { return false; }
int conv_wgrad_umma(const b200_tensor*, const b200_tensor*, float*, float*, int, int, int, cudaStream_t) {
  set_error("conv_wgrad_umma: not built");
  return B200_ERR_UNSUPPORTED;
}
}  // namespace b200

B200_EXPORT int b200_umma_selftest(int32_t, void*) {
  b200::set_error("umma_selftest: not built");
  return B200_ERR_UNSUPPORTED;
}
