// placeholder, replaced by the tcgen05/TMA implementation
#include "common.cuh"
namespace r2n2 {
int conv_fwd_umma(const void*, const void*, const float*, const void*, const void*, void*, int, int,
                  int, int, int, int, int, int, int, int, int, int, cudaStream_t) {
  set_error("conv_fwd[umma]: not built");
  return R2N2_ERR_UNSUPPORTED;
}
int conv_wgrad_umma(const void*, const void*, float*, int, int, int, int, int, int, int, int, int,
                    int, int, cudaStream_t) {
  set_error("conv_wgrad[umma]: not built");
  return R2N2_ERR_UNSUPPORTED;
}
}  // namespace r2n2
