// conv.cuh - interfaces shared by the FFMA (conv.cu) and tcgen05 (conv_tc.cu) convolution paths
#pragma once
#include "metadata.cuh"

namespace scn {

struct TileView {            // device view of a TileBook, passed by value to kernels
  int identity;
  int n_tiles;
  int k_flip;                // >= 0: table offset k uses weight slice k_flip - k (submanifold dX), else k
  int k_base;                // first filter offset of this book: its table offset j is filter offset k = k_base + j
  const int32_t *perm;
  const uint32_t *tile_mask;
  const int32_t *tile_off;
  const int4 *order;         // work-item order, tiles heaviest first: {tile, its offset mask, its first entry, 0} - ONE load
                             // per work item instead of order -> (mask, offset) (null: identity books)
  const int32_t *entries;
};

TileView make_view(const TileBook &tb, int k_flip = -1);

// Y[stationary] = bias + sum_k X[partner_k] @ Wg[k]; Kd = reduction width (row width of X), N = row
// width of Y.  transpose_w = 0: Wg[k] = W[k] with W [K,Kd,N]; transpose_w = 1: Wg[k] = W[k]^T with
// W [K,N,Kd] (the dX pass reads the forward weights in place).
// k_flip >= 0: the gather lists were built for the mirrored problem (submanifold dX through the
// forward lists): table offset k pairs with weight slice k_flip - k.
int osgemm(const float *X, const float *W, const float *bias, float *Y, int Kd, int N,
           const TileBook &tb, int precision, int transpose_w, cudaStream_t s, int k_flip = -1,
           const int64_t *weight_tag = nullptr);

// tensor-core variants. Return 0 = done, >0 = shape/precision not handled (caller uses the FFMA
// kernels), <0 = -(error) with scn_last_error set.
// K = offsets of the whole filter (what the weight tensor / its packed images hold), Kb = offsets of the tile book
int osgemm_tc(const float *X, const float *W, const float *bias, float *Y, int Kd, int N,
              long long n_rows, const TileView &tv, int K, int Kb, int precision, int transpose_w, cudaStream_t s,
              double prof_bytes, double prof_flops, const int64_t *weight_tag);
int dw_partial_tc(const float *X, const float *dY, const int32_t *pairs, const DwWork *work, float *partial,
                  int Cin, int Cout, int xcol, int ycol, int n_work, long long ident_n, int ident_chunk,
                  int precision, cudaStream_t s);

// layer-graph forward: all weight operand images in one launch, fresh until prepack_weights_end()
int prepack_weights_batch(int n, const int64_t *const *tags, const float *const *W, const int *K, const int *Cin,
                          const int *Cout, int precision, cudaStream_t s);
void prepack_weights_end();

int transpose_weights(const float *W, float *Wt, int K, int Cin, int Cout, cudaStream_t s);

// dW[k] = X[in_k]^T @ dY[out_k] for every offset (xcol / ycol: which pair column indexes X / dY); d_bias = column sums
int weight_grad(const float *X, const float *dY, float *dW, int Cin, int Cout, RuleBook *rb, int xcol, int ycol,
                int precision, cudaStream_t s);
int bias_grad(const float *d_out, float *d_bias, long long n, int C, cudaStream_t s);

// Weight gradients run on the companion stream of `s`.  Per-layer calls join it before returning; while
// g_defer_dw_join is set (layer-graph reverse sweep, per host thread) the join is left to dw_join_pending(s).
extern thread_local bool g_defer_dw_join, g_dw_join_pending;
extern bool g_dw_companion;
int dw_join_pending(cudaStream_t s);

}  // namespace scn
