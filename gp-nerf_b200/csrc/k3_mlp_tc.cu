// K3 (bf16 tensor-core path) – placeholder until the tcgen05 kernels land.
#include "common.cuh"

int gpnerf_density_mlp_tc(const float*, const float*, const float*, const gpnerf_head_weights_t*, int, int, int,
                          const int32_t*, float*, float*, cudaStream_t) {
  gpnerf::set_error("bf16 tcgen05 density head not built", cudaSuccess);
  return GPNERF_E_UNSUPPORTED;
}
int gpnerf_color_mlp_tc(const float*, const float*, const int32_t*, const gpnerf_head_weights_t*, int, int,
                        const int32_t*, float*, cudaStream_t) {
  gpnerf::set_error("bf16 tcgen05 colour head not built", cudaSuccess);
  return GPNERF_E_UNSUPPORTED;
}
