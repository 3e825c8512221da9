// rpn.cu - the RPN's per-site work on the device: anchor generation from the sparse
// maps' coordinates and the RPN head's three 1x1 convolutions.
//
// Reference:
//   maskrcnn_benchmark/modeling/rpn/anchor_generator_sparse3d.py:88-104 (grid_anchors) and :137-146 (forward):
//     every step copies get_spatial_locations() of each level to the HOST (Metadata.cpp:149-168), builds
//     centroids = (loc[:, 0:3].float() + 0) / voxel_scale * stride, adds the A base anchors and moves the result
//     back to the device (rpn_sparse3d.py:196-198); examples_bidx_2_sizes (:174-185) counts the rows of every
//     sample with one torch.sum per sample.
//   maskrcnn_benchmark/modeling/rpn/rpn_sparse3d.py:97-131 (RPNHead): features.t()[None, :, :, None] through
//     nn.Conv2d(C, C, 1) + ReLU, then nn.Conv2d(C, A S, 1) and nn.Conv2d(C, 7 A S, 1), permuted back to
//     [1, n, A, S] / [1, n, A, 7 S] - a transpose, three cuDNN calls and two permutes per level, 6 levels.
// Here: anchors are written by one kernel per level straight from the grid's device coordinates (no host copy),
// with the per-sample row ranges from a binary search on the batch-contiguous rows; the head runs on the
// library's row-major [n, C] layout through the 1x1 gather-GEMM (identity tile book: tcgen05 where the width
// allows, FFMA tiles otherwise) - the Conv2d weights [Cout, Cin, 1, 1] are consumed in place as the transposed
// operand, so no transpose, permute or weight copy exists.
#include "conv.cuh"
#include "../../include/scn_b200.h"

namespace scn {

// anchors[(r * A + a) * 7 + j] = centroid(r)[j] + base[a][j], centroid = (x,y,z) / voxel_scale * stride, 0, 0, 0, 0.
// Each float operation rounded separately, in the reference's order (division, multiplication, addition).
struct Stride3 { float v[3]; };
__global__ void k_grid_anchors(const int32_t *__restrict__ coords, long long n, const float *__restrict__ base, int A,
                               float voxel_scale, Stride3 st, float *__restrict__ out) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over n * A * 7
  if (i >= n * A * 7) return;
  const int j = (int)(i % 7);
  const long long ra = i / 7;
  const int a = (int)(ra % A);
  const long long r = ra / A;
  float c = 0.f;
  if (j < 3) c = __fmul_rn(__fdiv_rn((float)coords[r * 4 + j] + 0.f, voxel_scale), st.v[j]);
  out[i] = __fadd_rn(c, base[a * 7 + j]);
}
// scope[b] = {first row of sample b, one past its last row} * A  (rows are batch-contiguous ascending)
__global__ void k_example_scope(const int32_t *__restrict__ coords, long long n, int B, int A, int64_t *__restrict__ scope) {
  pdl_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  auto lower = [&](int v) {          // first row with batch >= v
    long long lo = 0, hi = n;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (coords[mid * 4 + 3] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  scope[2 * b] = lower(b) * A;
  scope[2 * b + 1] = lower(b + 1) * A;
}

__global__ void k_relu_inplace(float *__restrict__ x, long long n) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = fmaxf(x[i], 0.f);
}
// d *= (y > 0)   (gradient of ReLU from its output)
__global__ void k_relu_mask(float *__restrict__ d, const float *__restrict__ y, long long n) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(y[i] > 0.f)) d[i] = 0.f;
}
__global__ void k_add_inplace(float *__restrict__ y, const float *__restrict__ t, long long n) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] += t[i];
}
static int ew_grid(long long n) {
  long long b = (n + 255) / 256;
  if (b > (long long)num_sms() * 8) b = (long long)num_sms() * 8;
  return (int)(b < 1 ? 1 : b);
}

static TileBook identity_book(int64_t n) {
  TileBook tb;
  tb.identity = true; tb.built = true; tb.K = 1;
  tb.n_rows = tb.n_partner = n;
  tb.n_tiles = cdiv(n, TILE_M);
  tb.n_pairs = n;
  return tb;
}
static void identity_rulebook(RuleBook &rb, int64_t n) {
  rb.kind = 0; rb.K = 1; rb.identity = true;
  rb.n_in = rb.n_out = n;
  memset(rb.counts, 0, sizeof(rb.counts));
  memset(rb.pair_off, 0, sizeof(rb.pair_off));
  rb.counts[0] = n; rb.pair_off[1] = n; rb.total_pairs = n;
}

}  // namespace scn

using namespace scn;

extern "C" {

int scn_grid_anchors(scn_metadata_t *m, const int64_t *ss, const float *base_anchors, int64_t n_anchors,
                     float voxel_scale, const float *stride, float *anchors_out, int64_t *scope_out, int64_t batch_size,
                     void *stream) {
  SCN_CHECK(m && ss && stride, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  Grid *g = find_grid(m, ss);
  SCN_CHECK(g, "no active sites at spatial size [%lld,%lld,%lld]", (long long)ss[0], (long long)ss[1], (long long)ss[2]);
  SCN_CHECK(n_anchors > 0 && voxel_scale != 0.f, "bad anchor arguments");
  SCN_CHECK(g->batch_sorted, "anchor scopes need batch-contiguous rows");
  const long long total = g->n_active * n_anchors * 7;
  if (total > 0) {
    SCN_CHECK(base_anchors && anchors_out, "null pointer");
    Stride3 st{{stride[0], stride[1], stride[2]}};
    SCN_LAUNCH(k_grid_anchors, cdiv(total, 256), 256, 0, s, g->coords, g->n_active, base_anchors, (int)n_anchors, voxel_scale, st,
                                                   anchors_out);
    SCN_LAUNCHED();
  }
  if (scope_out && batch_size > 0) {
    SCN_LAUNCH(k_example_scope, cdiv(batch_size, 128), 128, 0, s, g->coords, g->n_active, (int)batch_size, (int)n_anchors, scope_out);
    SCN_LAUNCHED();
  }
  return 0;
}

// hidden = relu(x Wc^T + bc); logits = hidden Wl^T + bl; reg = hidden Wr^T + br.  Weights in nn.Conv2d layout
// [Cout, Cin] (1x1 kernels), all buffers on the device; hidden [n, C] is kept by the caller for the backward pass.
int scn_rpn_head_forward(const float *x, int64_t n, int64_t C, const float *w_conv, const float *b_conv,
                         const float *w_cls, const float *b_cls, int64_t n_cls, const float *w_box, const float *b_box,
                         int64_t n_box, float *hidden, float *logits, float *reg, int precision, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) return 0;
  SCN_CHECK(x && w_conv && w_cls && w_box && hidden && logits && reg, "null pointer");
  const TileBook tb = identity_book(n);
  SCN_TRY(osgemm(x, w_conv, b_conv, hidden, (int)C, (int)C, tb, precision, /*transpose_w=*/1, s));
  SCN_LAUNCH(k_relu_inplace, ew_grid(n * C), 256, 0, s, hidden, n * C);
  SCN_LAUNCHED();
  SCN_TRY(osgemm(hidden, w_cls, b_cls, logits, (int)C, (int)n_cls, tb, precision, 1, s));
  SCN_TRY(osgemm(hidden, w_box, b_box, reg, (int)C, (int)n_box, tb, precision, 1, s));
  return 0;
}

// gradients of everything; d_hidden is scratch [n, C] supplied by the caller, d_x may be NULL
int scn_rpn_head_backward(const float *x, const float *hidden, int64_t n, int64_t C, const float *w_conv,
                          const float *w_cls, int64_t n_cls, const float *w_box, int64_t n_box, const float *d_logits,
                          const float *d_reg, float *d_hidden, float *d_x, float *dw_conv, float *db_conv,
                          float *dw_cls, float *db_cls, float *dw_box, float *db_box, int precision, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    if (dw_conv) SCN_CUDA(cudaMemsetAsync(dw_conv, 0, (size_t)C * C * 4, s));
    if (db_conv) SCN_CUDA(cudaMemsetAsync(db_conv, 0, (size_t)C * 4, s));
    if (dw_cls) SCN_CUDA(cudaMemsetAsync(dw_cls, 0, (size_t)n_cls * C * 4, s));
    if (db_cls) SCN_CUDA(cudaMemsetAsync(db_cls, 0, (size_t)n_cls * 4, s));
    if (dw_box) SCN_CUDA(cudaMemsetAsync(dw_box, 0, (size_t)n_box * C * 4, s));
    if (db_box) SCN_CUDA(cudaMemsetAsync(db_box, 0, (size_t)n_box * 4, s));
    return 0;
  }
  SCN_CHECK(x && hidden && d_logits && d_reg && d_hidden && w_conv && w_cls && w_box, "null pointer");
  const TileBook tb = identity_book(n);
  RuleBook rb;
  identity_rulebook(rb, n);
  // d_hidden = d_logits Wl + d_reg Wr   (W [Cout, Cin] read as the [Kd = Cout, N = Cin] operand: no transpose)
  float *part = nullptr;
  SCN_TRY(workspace_t(&part, WS_CHAIN, (size_t)n * C, s));
  SCN_TRY(osgemm(d_logits, w_cls, nullptr, d_hidden, (int)n_cls, (int)C, tb, precision, 0, s));
  SCN_TRY(osgemm(d_reg, w_box, nullptr, part, (int)n_box, (int)C, tb, precision, 0, s));
  SCN_LAUNCH(k_add_inplace, ew_grid(n * C), 256, 0, s, d_hidden, part, n * C);
  SCN_LAUNCHED();
  // head weights: dW [Cout, Cin] = d_out^T hidden
  if (dw_cls) SCN_TRY(weight_grad(d_logits, hidden, dw_cls, (int)n_cls, (int)C, &rb, 0, 1, precision, s));
  if (dw_box) SCN_TRY(weight_grad(d_reg, hidden, dw_box, (int)n_box, (int)C, &rb, 0, 1, precision, s));
  SCN_TRY(bias_grad(d_logits, db_cls, n, (int)n_cls, s));
  SCN_TRY(bias_grad(d_reg, db_box, n, (int)n_box, s));
  SCN_LAUNCH(k_relu_mask, ew_grid(n * C), 256, 0, s, d_hidden, hidden, n * C);
  SCN_LAUNCHED();
  if (dw_conv) SCN_TRY(weight_grad(d_hidden, x, dw_conv, (int)C, (int)C, &rb, 0, 1, precision, s));
  SCN_TRY(bias_grad(d_hidden, db_conv, n, (int)C, s));
  if (d_x) SCN_TRY(osgemm(d_hidden, w_conv, nullptr, d_x, (int)C, (int)C, tb, precision, 0, s));
  return 0;
}

}  // extern "C"
