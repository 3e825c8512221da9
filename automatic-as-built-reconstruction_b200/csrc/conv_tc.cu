// conv_tc.cu - tcgen05 / TMEM gather-GEMM (bf16 and 3xTF32).  Placeholder until the tensor
// core kernels land: reports "not handled" so callers use the exact-fp32 FFMA tiles.
#include "conv.cuh"

namespace scn {
int osgemm_tc(const float *, const float *, const float *, float *, int, int, long long,
              const TileView &, int, int, cudaStream_t) { return 1; }
int weight_grad_tc(const float *, const float *, float *, int, int, RuleBook *, int, int, int,
                   cudaStream_t) { return 1; }
}  // namespace scn
