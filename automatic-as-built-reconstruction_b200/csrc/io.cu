// io.cu - InputLayer / OutputLayer feature movement, SparseToDense, voxel quantisation.
//
// Reference: SparseConvNet/sparseconvnet/SCN/CPU/IOLayers.cpp:12-140 (segment sum / mean in
// list order, inverse scatter), CPU/SparseToDense.cpp:8-87, and the numpy quantiser in
// data3d/suncg_utils/suncg_dataset.py:126-188.
#include "metadata.cuh"
#include "../../include/scn_b200.h"

namespace scn {

// out[row][c] = sum_i mult * in[member_i][c] in ascending point order (IOLayers.cpp:17-28)
__global__ void k_input_fwd(const float *__restrict__ in, float *__restrict__ out,
                            const int32_t *__restrict__ csr_off, const int32_t *__restrict__ members,
                            long long n_rows, int C, int mode) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * C) return;
  const long long row = i / C;
  const int c = (int)(i - row * C);
  const int b = csr_off[row], e = csr_off[row + 1];
  float acc = 0.f;
  if (mode == 1) acc = in[(long long)members[b] * C + c];
  else if (mode == 2) acc = in[(long long)members[e - 1] * C + c];
  else {
    const float mult = (mode == 4 && e > b) ? 1.f / (float)(e - b) : 1.f;
    for (int j = b; j < e; ++j) acc += mult * in[(long long)members[j] * C + c];
  }
  out[i] = acc;
}

// d_in[point][c] = mult(row) * d_out[row][c]   (gather form of IOLayers.cpp:31-46; no atomics)
__global__ void k_input_bwd(float *__restrict__ d_in, const float *__restrict__ d_out,
                            const int32_t *__restrict__ point_row, const int32_t *__restrict__ csr_off,
                            const int32_t *__restrict__ members, long long n_points, int C, int mode,
                            int average) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points * C) return;
  const long long p = i / C;
  const int c = (int)(i - p * C);
  const int row = point_row[p];
  const int b = csr_off[row], e = csr_off[row + 1];
  float v = d_out[(long long)row * C + c];
  if (mode == 1) v = (members[b] == (int)p) ? v : 0.f;
  else if (mode == 2) v = (members[e - 1] == (int)p) ? v : 0.f;
  else if (average) v *= 1.f / (float)(e - b);
  d_in[i] = v;
}

// out[row][c] = sum over the row's selected points of d[point][c]  (no averaging)
__global__ void k_output_bwd(float *__restrict__ d_in, const float *__restrict__ d_out,
                             const int32_t *__restrict__ csr_off, const int32_t *__restrict__ members,
                             long long n_rows, int C, int mode) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * C) return;
  const long long row = i / C;
  const int c = (int)(i - row * C);
  const int b = csr_off[row], e = csr_off[row + 1];
  float acc = 0.f;
  if (mode == 1) acc = d_out[(long long)members[b] * C + c];
  else if (mode == 2) acc = d_out[(long long)members[e - 1] * C + c];
  else
    for (int j = b; j < e; ++j) acc += d_out[(long long)members[j] * C + c];
  d_in[i] = acc;
}

// ---- SparseToDense ----------------------------------------------------------------------
// dense[b][c][off] with off = (x*Y+y)*Z+z (ConvolutionRules.h:110-128, SparseToDense.cpp:14-17)
__global__ void k_s2d_fwd(const float *__restrict__ in, float *__restrict__ out,
                          const int32_t *__restrict__ coords, long long n, int C, long long Y,
                          long long Z, long long V) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * C) return;
  // consecutive threads take consecutive rows of one plane: reads stride C, writes follow the
  // site order (neighbouring z are neighbouring addresses)
  const int c = (int)(i / n);
  const long long r = i - (long long)c * n;
  const int4 p = reinterpret_cast<const int4 *>(coords)[r];
  const long long off = ((long long)p.x * Y + p.y) * Z + p.z;
  out[((long long)p.w * C + c) * V + off] = in[r * C + c];
}
__global__ void k_s2d_bwd(float *__restrict__ d_in, const float *__restrict__ d_out,
                          const int32_t *__restrict__ coords, long long n, int C, long long Y,
                          long long Z, long long V) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * C) return;
  const int c = (int)(i / n);
  const long long r = i - (long long)c * n;
  const int4 p = reinterpret_cast<const int4 *>(coords)[r];
  const long long off = ((long long)p.x * Y + p.y) * Z + p.z;
  d_in[r * C + c] = d_out[((long long)p.w * C + c) * V + off];
}

// ---- voxel quantisation -------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_double(double *addr, double v) {
  unsigned long long *a = (unsigned long long *)addr;
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

__global__ void k_quant_min(const double *__restrict__ xyz, long long n, double scale, double *mn) {
  pdl_sync();
  __shared__ double sm[3][256];
  double m[3] = {1e300, 1e300, 1e300};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
#pragma unroll
    for (int d = 0; d < 3; ++d) m[d] = fmin(m[d], __dmul_rn(xyz[i * 3 + d], scale));
  for (int d = 0; d < 3; ++d) sm[d][threadIdx.x] = m[d];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int d = 0; d < 3; ++d) sm[d][threadIdx.x] = fmin(sm[d][threadIdx.x], sm[d][threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x < 3) atomic_min_double(&mn[threadIdx.x], sm[threadIdx.x][0]);
}

struct Full3 { long long v[3]; };

__global__ void k_quant_flag(const double *__restrict__ xyz, long long n, double scale,
                             const double *__restrict__ mn, Full3 full, uint8_t *__restrict__ keep,
                             int32_t *__restrict__ flag) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool k = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    // two roundings like numpy (a = x*scale; a -= min): no FMA contraction
    const double a = __dsub_rn(__dmul_rn(xyz[i * 3 + d], scale), mn[d]);
    k = k && (a >= 0.0) && (a < (double)full.v[d]);
  }
  keep[i] = k ? 1 : 0;
  flag[i] = k ? 1 : 0;
}

__global__ void k_quant_emit(const double *__restrict__ xyz, long long n, double scale,
                             const double *__restrict__ mn, const int32_t *__restrict__ pos,
                             long long batch_idx, int64_t *__restrict__ coords) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = pos[i];
  if (pos[i + 1] == p) return;
  int64_t *o = coords + (long long)p * 4;
#pragma unroll
  for (int d = 0; d < 3; ++d) o[d] = (int64_t)__dsub_rn(__dmul_rn(xyz[i * 3 + d], scale), mn[d]);  // trunc, a >= 0
  o[3] = batch_idx;
}

// ---- batched voxelisation front end ------------------------------------------
// points [N, C] float32 (columns 0-2 = xyz in metres) of B buildings back to back, first[b] = first point of
// building b.  a = xyz @ M in float64 (data3d/suncg_utils/suncg_dataset.py:126-137: M = eye(3) * scale times the
// augmentations), a -= a.min(0) per building (:140-147), keep 0 <= a < full_scale (:171-183), truncate.
struct Vox { double m[9]; long long full[3]; double scale; };

__global__ void k_fill_double(double *p, double v, long long n) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__device__ __forceinline__ int vox_building(const int64_t *__restrict__ first, int B, long long i) {
  int lo = 0, hi = B;                       // last b with first[b] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (first[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}
// a[d] = (x m[0][d] + y m[1][d]) + z m[2][d], every operation rounded separately (no FMA): exact - and therefore
// equal to numpy's matmul in any summation order - whenever M is diagonal (the shipped configs: zoom / rotation off)
__device__ __forceinline__ void vox_transform(const float *__restrict__ p, const Vox &v, double a[3]) {
  const double x = p[0], y = p[1], z = p[2];
#pragma unroll
  for (int d = 0; d < 3; ++d)
    a[d] = __dadd_rn(__dadd_rn(__dmul_rn(x, v.m[d]), __dmul_rn(y, v.m[3 + d])), __dmul_rn(z, v.m[6 + d]));
}

// per-building minimum of the transformed coordinates: a block owns VOX_PER_BLOCK consecutive points; when they
// all belong to one building (all but <= B - 1 blocks) it reduces in shared memory and issues 3 atomics
constexpr int VOX_PER_BLOCK = 256 * 8;
__global__ void __launch_bounds__(256)
k_vox_min(const float *__restrict__ pts, int C, long long n, const int64_t *__restrict__ first, int B,
          Vox v, double *__restrict__ mn /* [B,3] */) {
  pdl_sync();
  __shared__ double sm[3][256];
  const long long p0 = (long long)blockIdx.x * VOX_PER_BLOCK;
  const long long p1 = min(p0 + VOX_PER_BLOCK, n);
  const int b0 = vox_building(first, B, p0), b1 = vox_building(first, B, p1 - 1);
  double m[3] = {1e300, 1e300, 1e300};
  for (long long i = p0 + threadIdx.x; i < p1; i += 256) {
    double a[3];
    vox_transform(pts + i * C, v, a);
    if (b0 == b1) {
#pragma unroll
      for (int d = 0; d < 3; ++d) m[d] = fmin(m[d], a[d]);
    } else {
      const int b = vox_building(first, B, i);
#pragma unroll
      for (int d = 0; d < 3; ++d) atomic_min_double(&mn[b * 3 + d], a[d]);
    }
  }
  if (b0 != b1) return;                      // (block-uniform)
#pragma unroll
  for (int d = 0; d < 3; ++d) sm[d][threadIdx.x] = m[d];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int d = 0; d < 3; ++d) sm[d][threadIdx.x] = fmin(sm[d][threadIdx.x], sm[d][threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x < 3) atomic_min_double(&mn[b0 * 3 + threadIdx.x], sm[threadIdx.x][0]);
}

__global__ void k_vox_flag(const float *__restrict__ pts, int C, long long n, const int64_t *__restrict__ first, int B,
                           Vox v, const double *__restrict__ mn, int32_t *__restrict__ flag) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[3];
  vox_transform(pts + i * C, v, a);
  const int b = vox_building(first, B, i);
  bool k = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double q = __dsub_rn(a[d], mn[b * 3 + d]);
    k = k && (q >= 0.0) && (q < (double)v.full[d]);
  }
  flag[i] = k ? 1 : 0;
}

__global__ void k_vox_emit(const float *__restrict__ pts, int C, long long n, const int64_t *__restrict__ first, int B,
                           Vox v, const double *__restrict__ mn, const int32_t *__restrict__ pos,
                           int64_t *__restrict__ coords, float *__restrict__ feats, int xyz_feature) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = pos[i];
  if (pos[i + 1] == p) return;
  double a[3];
  vox_transform(pts + i * C, v, a);
  const int b = vox_building(first, B, i);
  int64_t *o = coords + (long long)p * 4;
  float *f = feats + (long long)p * C;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double q = __dsub_rn(a[d], mn[b * 3 + d]);
    o[d] = (int64_t)q;                                               // trunc, q >= 0 (a.long())
    f[d] = xyz_feature ? (float)__ddiv_rn(q, v.scale) : pts[i * C + d];   // b[:,0:3] = a / scale (:160-162)
  }
  o[3] = b;
  for (int c = 3; c < C; ++c) f[c] = pts[i * C + c];
}

}  // namespace scn


using namespace scn;

extern "C" {

int scn_input_layer_forward(scn_metadata_t *m, const float *in, float *out, int64_t C,
                            void *stream) {
  SCN_CHECK(m && m->input.built, "input layer not prepared");
  InputRules &ir = m->input;
  const long long total = ir.n_active * C;
  if (total == 0) return 0;
  SCN_CHECK(in && out, "null feature pointer");
  SCN_LAUNCH(k_input_fwd, cdiv(total, 256), 256, 0, (cudaStream_t)stream, in, out, ir.csr_off, ir.members, ir.n_active, (int)C, ir.mode == 0 ? 3 : ir.mode);
  SCN_LAUNCHED();
  return 0;
}

int scn_input_layer_backward(scn_metadata_t *m, float *d_in, const float *d_out, int64_t C,
                             void *stream) {
  SCN_CHECK(m && m->input.built, "input layer not prepared");
  InputRules &ir = m->input;
  const long long total = ir.n_points * C;
  if (total == 0) return 0;
  SCN_CHECK(d_in && d_out, "null feature pointer");
  SCN_LAUNCH(k_input_bwd, cdiv(total, 256), 256, 0, (cudaStream_t)stream, d_in, d_out, ir.point_row, ir.csr_off, ir.members, ir.n_points, (int)C,
      ir.mode == 0 ? 3 : ir.mode, ir.mode == 4);
  SCN_LAUNCHED();
  return 0;
}

// OutputLayer forward = InputLayer backward without averaging (CPU/IOLayers.cpp:91-109)
int scn_output_layer_forward(scn_metadata_t *m, const float *in, float *out, int64_t C,
                             void *stream) {
  SCN_CHECK(m && m->input.built, "input layer not prepared");
  InputRules &ir = m->input;
  const long long total = ir.n_points * C;
  if (total == 0) return 0;
  SCN_CHECK(in && out, "null feature pointer");
  SCN_LAUNCH(k_input_bwd, cdiv(total, 256), 256, 0, (cudaStream_t)stream, out, in, ir.point_row, ir.csr_off, ir.members, ir.n_points, (int)C,
      ir.mode == 0 ? 3 : ir.mode, 0);
  SCN_LAUNCHED();
  return 0;
}

// OutputLayer backward = InputLayer forward without averaging (CPU/IOLayers.cpp:111-131)
int scn_output_layer_backward(scn_metadata_t *m, float *d_in, const float *d_out, int64_t C,
                              void *stream) {
  SCN_CHECK(m && m->input.built, "input layer not prepared");
  InputRules &ir = m->input;
  const long long total = ir.n_active * C;
  if (total == 0) return 0;
  SCN_CHECK(d_in && d_out, "null feature pointer");
  SCN_LAUNCH(k_output_bwd, cdiv(total, 256), 256, 0, (cudaStream_t)stream, d_in, d_out, ir.csr_off, ir.members, ir.n_active, (int)C, ir.mode == 0 ? 3 : ir.mode);
  SCN_LAUNCHED();
  return 0;
}

int scn_sparse_to_dense_forward(scn_metadata_t *m, const int64_t *ss, const float *in, float *out,
                                int64_t C, int64_t batch, void *stream) {
  SCN_CHECK(m && ss && out, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const long long V = ss[0] * ss[1] * ss[2];
  SCN_CUDA(cudaMemsetAsync(out, 0, (size_t)batch * C * V * 4, s));
  Grid *g = find_grid(m, ss);
  if (!g || g->n_active == 0) return 0;
  SCN_CHECK(in, "null feature pointer");
  const long long total = g->n_active * C;
  SCN_LAUNCH(k_s2d_fwd, cdiv(total, 256), 256, 0, s, in, out, g->coords, g->n_active, (int)C, ss[1], ss[2], V);
  SCN_LAUNCHED();
  return 0;
}

// the same into the CROPPED volume [batch, C, ext0, ext1, ext2] = dense[:, :, :ext0, :ext1, :ext2] of the full tensor
// (what tools_3d_2d.py:28 slices out after densifying all of [X, Y, Z]: 1.07 GB zero-filled for a [256,256,32] ROI map
// of which the benchmark building occupies 16 MB)
int scn_sparse_to_dense_cropped_forward(scn_metadata_t *m, const int64_t *ss, const int64_t *ext, const float *in,
                                        float *out, int64_t C, int64_t batch, void *stream) {
  SCN_CHECK(m && ss && ext && out, "null argument");
  SCN_CHECK(ext[0] >= 0 && ext[1] >= 0 && ext[2] >= 0 && ext[0] <= ss[0] && ext[1] <= ss[1] && ext[2] <= ss[2],
            "SparseToDense: extent [%lld,%lld,%lld] outside the spatial size", (long long)ext[0], (long long)ext[1],
            (long long)ext[2]);
  cudaStream_t s = (cudaStream_t)stream;
  const long long V = ext[0] * ext[1] * ext[2];
  if (batch * C * V == 0) return 0;
  SCN_CUDA(cudaMemsetAsync(out, 0, (size_t)batch * C * V * 4, s));
  Grid *g = find_grid(m, ss);
  if (!g || g->n_active == 0) return 0;
  SCN_CHECK(in, "null feature pointer");
  const long long total = g->n_active * C;
  SCN_LAUNCH(k_s2d_fwd, cdiv(total, 256), 256, 0, s, in, out, g->coords, g->n_active, (int)C, ext[1], ext[2], V);
  SCN_LAUNCHED();
  return 0;
}

int scn_sparse_to_dense_cropped_backward(scn_metadata_t *m, const int64_t *ss, const int64_t *ext, float *d_in,
                                         const float *d_out, int64_t C, int64_t batch, void *stream) {
  SCN_CHECK(m && ss && ext, "null argument");
  (void)batch;
  Grid *g = find_grid(m, ss);
  if (!g || g->n_active == 0) return 0;
  SCN_CHECK(d_in && d_out, "null feature pointer");
  const long long V = ext[0] * ext[1] * ext[2];
  const long long total = g->n_active * C;
  SCN_LAUNCH(k_s2d_bwd, cdiv(total, 256), 256, 0, (cudaStream_t)stream, d_in, d_out, g->coords, g->n_active,
             (int)C, ext[1], ext[2], V);
  SCN_LAUNCHED();
  return 0;
}

int scn_sparse_to_dense_backward(scn_metadata_t *m, const int64_t *ss, float *d_in,
                                 const float *d_out, int64_t C, int64_t batch, void *stream) {
  SCN_CHECK(m && ss, "null argument");
  (void)batch;
  Grid *g = find_grid(m, ss);
  if (!g || g->n_active == 0) return 0;
  SCN_CHECK(d_in && d_out, "null feature pointer");
  const long long V = ss[0] * ss[1] * ss[2];
  const long long total = g->n_active * C;
  SCN_LAUNCH(k_s2d_bwd, cdiv(total, 256), 256, 0, (cudaStream_t)stream, d_in, d_out, g->coords, g->n_active,
                                                              (int)C, ss[1], ss[2], V);
  SCN_LAUNCHED();
  return 0;
}

int scn_quantize_points(const double *xyz, int64_t n, double scale, const int64_t *full_scale,
                        int64_t batch_idx, int64_t *coords_out, uint8_t *keep_out, int64_t *n_kept,
                        void *stream) {
  SCN_CHECK(full_scale && n_kept, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  *n_kept = 0;
  if (n == 0) return 0;
  SCN_CHECK(xyz && coords_out && keep_out, "null pointer");
  double *mn = nullptr;
  int32_t *pos = nullptr;
  SCN_TRY(dev_alloc_t(&mn, 4, s));
  SCN_TRY(dev_alloc_t(&pos, (size_t)n + 1, s));
  const double big[3] = {1e300, 1e300, 1e300};
  SCN_CUDA(cudaMemcpyAsync(mn, big, 24, cudaMemcpyHostToDevice, s));
  int gb = cdiv(n, 256);
  if (gb > num_sms() * 4) gb = num_sms() * 4;
  SCN_LAUNCH(k_quant_min, gb, 256, 0, s, xyz, n, scale, mn);
  SCN_LAUNCHED();
  Full3 f;
  for (int d = 0; d < 3; ++d) f.v[d] = full_scale[d];
  SCN_LAUNCH(k_quant_flag, cdiv(n, 256), 256, 0, s, xyz, n, scale, mn, f, keep_out, pos);
  SCN_LAUNCHED();
  SCN_TRY(exclusive_scan_i32(pos, pos, n, s));
  SCN_LAUNCH(k_quant_emit, cdiv(n, 256), 256, 0, s, xyz, n, scale, mn, pos, batch_idx, coords_out);
  SCN_LAUNCHED();
  int32_t *h32 = (int32_t *)host_scratch(16);
  SCN_CUDA(cudaMemcpyAsync(h32, pos + n, 4, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));
  *n_kept = h32[0];
  dev_free(mn, s);
  dev_free(pos, s);
  return 0;
}

int scn_voxelize_batch(const float *points, int64_t n, int64_t n_cols, const int64_t *first_dev, int64_t n_buildings,
                       const double *matrix, double scale, const int64_t *full_scale, int xyz_feature,
                       int64_t *coords_out, float *feats_out, int64_t *n_kept, void *stream) {
  SCN_CHECK(full_scale && n_kept && matrix, "null argument");
  SCN_CHECK(n_cols >= 3, "voxelize: points need at least the 3 xyz columns, got %lld", (long long)n_cols);
  SCN_CHECK(n < (1LL << 31), "voxelize: %lld points do not fit the int32 row space", (long long)n);
  cudaStream_t s = (cudaStream_t)stream;
  *n_kept = 0;
  if (n == 0 || n_buildings == 0) return 0;
  SCN_CHECK(points && first_dev && coords_out && feats_out, "null pointer");
  prof_begin(PROF_IO, s);
  Vox v;
  for (int i = 0; i < 9; ++i) v.m[i] = matrix[i];
  for (int d = 0; d < 3; ++d) v.full[d] = full_scale[d];
  v.scale = scale;
  double *mn = nullptr;
  int32_t *pos = nullptr;
  SCN_TRY(dev_alloc_t(&mn, (size_t)n_buildings * 3, s));
  SCN_TRY(dev_alloc_t(&pos, (size_t)n + 1, s));
  SCN_LAUNCH(k_fill_double, cdiv(n_buildings * 3, 256), 256, 0, s, mn, 1e300, n_buildings * 3);
  SCN_LAUNCHED();
  const int C = (int)n_cols, B = (int)n_buildings;
  SCN_LAUNCH(k_vox_min, cdiv(n, VOX_PER_BLOCK), 256, 0, s, points, C, n, first_dev, B, v, mn);
  SCN_LAUNCHED();
  SCN_LAUNCH(k_vox_flag, cdiv(n, 256), 256, 0, s, points, C, n, first_dev, B, v, mn, pos);
  SCN_LAUNCHED();
  SCN_TRY(exclusive_scan_i32(pos, pos, n, s));
  SCN_LAUNCH(k_vox_emit, cdiv(n, 256), 256, 0, s, points, C, n, first_dev, B, v, mn, pos, coords_out, feats_out, xyz_feature);
  SCN_LAUNCHED();
  int32_t *h32 = (int32_t *)host_scratch(16);
  SCN_CUDA(cudaMemcpyAsync(h32, pos + n, 4, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));    // documented read-back: the batch's row count (tensor shapes)
  *n_kept = h32[0];
  prof_end(PROF_IO, s, (double)n * n_cols * 4.0 + (double)*n_kept * (32.0 + n_cols * 4.0), 0);
  dev_free(mn, s);
  dev_free(pos, s);
  return 0;
}

}  // extern "C"
