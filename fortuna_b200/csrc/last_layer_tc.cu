// K2 (tensor-core path): S-batched last-layer GEMM on tcgen05 / TMEM / TMA.  Placeholder until the
// kernel lands: reports "unsupported" so fb200_last_layer_logits routes every shape to the CUDA-core path.
#include "common.cuh"

namespace fb200 {

int last_layer_tc_supported(int, int, int, int, int, int, int) { return 0; }

int last_layer_tc_launch(const void*, const void*, const float*, const float*, void*, int, int, int, int, int,
                         cudaStream_t) {
  return FB200_E_UNSUPPORTED;
}

}  // namespace fb200
