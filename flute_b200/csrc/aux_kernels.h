// Host launchers of the auxiliary kernels (internal).
#pragma once

#include <cuda_runtime.h>

#include "../../include/flute_b200.h"

namespace fb {

int dequantize_launch(const void* Q, const void* S, const void* table2, void* What, int N, int K, int bits, int group,
                      int tile_p, int bf16, cudaStream_t stream);
int hadamard_launch(const void* in, void* out, long rows, int h, int bf16, cudaStream_t stream);

}  // namespace fb
