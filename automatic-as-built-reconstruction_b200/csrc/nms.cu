// nms.cu - rotated (BEV) box IoU and rotated NMS on the device.
//
// Reference call path: maskrcnn_benchmark/structures/boxlist_ops_3d.py:13-70 boxlist_nms_3d ->
// second/pytorch/core/box_torch_ops.py:557-582 rotate_nms_3d (top-k by score, pre_max_size) ->
// second/core/non_max_suppression/nms_cpu.py:32-44 rotate_nms_3d_cc: the N x N rotated IoU of the BEV boxes
// (utils3d/rotate_nms_3d_torch.py:22-84 boxes_iou_3d -> nms_gpu.py:667-703 rotate_iou_gpu_eval, numba-CUDA kernels
// JIT-compiled at first use, boxes host -> device and the IoU matrix device -> host every call), then the greedy
// pass on the HOST in spconv.utils.rotate_non_max_suppression_cpu (spconv 1.x, un-vendored third-party: boxes in
// score order; a kept box i suppresses every later j whose pre-filter IoU is > 0 and whose polygon overlap -
// boost::geometry intersection / union of the two corner quadrilaterals - is >= thresh).
//
// Here everything stays on the device: scores are sorted by the library's radix sort, one kernel fills the
// 64-box-block suppression bit matrix from the rotated IoU (the same quadrilateral clipping as the reference's
// numba device functions, nms_gpu.py:166-404: corners -> contained vertices + edge intersections -> angular sort
// -> triangle fan area, float32 with the area accumulated in double), and a single-CTA kernel walks the matrix
// greedily.  Only the number of kept boxes returns to the host.  The polygon overlap the reference's host pass
// recomputes with boost::geometry is the same quantity as this IoU up to rounding, so keep lists agree unless an
// IoU lies within rounding of the threshold.
#include "common.cuh"
#include "../../include/scn_b200.h"
#include <math.h>

namespace scn {

// ---- rotated IoU of two BEV boxes (x, y, size_x, size_y, yaw) -----------------------------------------------
__device__ __forceinline__ void box_corners(const float *b, float *c /* 8 */) {      // nms_gpu.py:355-378
  const float ca = cosf(b[4]), sa = sinf(b[4]);
  const float hx = b[2] / 2.f, hy = b[3] / 2.f;
  const float px[4] = {-hx, -hx, hx, hx}, py[4] = {-hy, hy, hy, -hy};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c[2 * i] = ca * px[i] + sa * py[i] + b[0];
    c[2 * i + 1] = -sa * px[i] + ca * py[i] + b[1];
  }
}
__device__ __forceinline__ bool point_in_quad(float x, float y, const float *c) {    // :310-328
  const float ab0 = c[2] - c[0], ab1 = c[3] - c[1], ad0 = c[6] - c[0], ad1 = c[7] - c[1];
  const float ap0 = x - c[0], ap1 = y - c[1];
  const float abab = ab0 * ab0 + ab1 * ab1, abap = ab0 * ap0 + ab1 * ap1;
  const float adad = ad0 * ad0 + ad1 * ad1, adap = ad0 * ap0 + ad1 * ap1;
  return abab >= abap && abap >= 0.f && adad >= adap && adap >= 0.f;
}
// intersection of edge i of quadrilateral p with edge j of quadrilateral q (:222-265)
__device__ __forceinline__ bool edge_intersection(const float *p, const float *q, int i, int j, float *out) {
  const float A0 = p[2 * i], A1 = p[2 * i + 1], B0 = p[2 * ((i + 1) & 3)], B1 = p[2 * ((i + 1) & 3) + 1];
  const float C0 = q[2 * j], C1 = q[2 * j + 1], D0 = q[2 * ((j + 1) & 3)], D1 = q[2 * ((j + 1) & 3) + 1];
  const float BA0 = B0 - A0, BA1 = B1 - A1, DA0 = D0 - A0, CA0 = C0 - A0, DA1 = D1 - A1, CA1 = C1 - A1;
  const bool acd = DA1 * CA0 > CA1 * DA0;
  const bool bcd = (D1 - B1) * (C0 - B0) > (C1 - B1) * (D0 - B0);
  if (acd == bcd) return false;
  const bool abc = CA1 * BA0 > BA1 * CA0, abd = DA1 * BA0 > BA1 * DA0;
  if (abc == abd) return false;
  const float DC0 = D0 - C0, DC1 = D1 - C1;
  const float ABBA = A0 * B1 - B0 * A1, CDDC = C0 * D1 - D0 * C1;
  const float DH = BA1 * DC0 - BA0 * DC1;
  out[0] = (ABBA * DC0 - BA0 * CDDC) / DH;
  out[1] = (ABBA * DC1 - BA1 * CDDC) / DH;
  return true;
}
// area of the intersection polygon of the two boxes (:331-395): up to 8 contained vertices + 16 edge crossings
// can be reported for degenerate (duplicate / touching) boxes; the reference's scratch holds 8 points, the first 8
// found are what it keeps in bounds - here the list is capped at 8 the same way
__device__ float intersection_area(const float *b1, const float *b2) {
  float c1[8], c2[8], pts[16];
  box_corners(b1, c1);
  box_corners(b2, c2);
  int n = 0;
  auto push = [&](float x, float y) { if (n < 8) { pts[2 * n] = x; pts[2 * n + 1] = y; } ++n; };
  for (int i = 0; i < 4; ++i) {
    if (point_in_quad(c1[2 * i], c1[2 * i + 1], c2)) push(c1[2 * i], c1[2 * i + 1]);
    if (point_in_quad(c2[2 * i], c2[2 * i + 1], c1)) push(c2[2 * i], c2[2 * i + 1]);
  }
  float t[2];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (edge_intersection(c1, c2, i, j, t)) push(t[0], t[1]);
  if (n > 8) n = 8;
  if (n > 0) {   // order the vertices by angle around their centroid (:182-219): key in [-3, 1], insertion sort
    float cx = 0.f, cy = 0.f, key[8];
    for (int i = 0; i < n; ++i) { cx += pts[2 * i]; cy += pts[2 * i + 1]; }
    cx /= (float)n; cy /= (float)n;
    for (int i = 0; i < n; ++i) {
      float vx = pts[2 * i] - cx, vy = pts[2 * i + 1] - cy;
      const float d = sqrtf(vx * vx + vy * vy);
      vx /= d; vy /= d;
      if (vy < 0.f) vx = -2.f - vx;
      key[i] = vx;
    }
    for (int i = 1; i < n; ++i)
      if (key[i - 1] > key[i]) {
        const float k = key[i], tx = pts[2 * i], ty = pts[2 * i + 1];
        int j = i;
        while (j > 0 && key[j - 1] > k) {
          key[j] = key[j - 1]; pts[2 * j] = pts[2 * j - 2]; pts[2 * j + 1] = pts[2 * j - 1];
          --j;
        }
        key[j] = k; pts[2 * j] = tx; pts[2 * j + 1] = ty;
      }
  }
  double area = 0.0;                                   // (:172-179) triangle fan from vertex 0
  for (int i = 0; i < n - 2; ++i) {
    const float *a = pts, *b = pts + 2 * i + 2, *c = pts + 2 * i + 4;
    area += fabs((double)((a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0])) / 2.0);
  }
  return (float)area;
}

// devRotateIoUEval (:552-623): r1 = the query box ("anchor"), r2 = the box ("target")
__device__ float rotate_iou_eval(const float *r1, const float *r2, int criterion) {
  const float area1 = r1[2] * r1[3], area2 = r2[2] * r2[3];
  const float inter = intersection_area(r1, r2);
  const float dc = sqrtf((r1[0] - r2[0]) * (r1[0] - r2[0]) + (r1[1] - r2[1]) * (r1[1] - r2[1]));
  switch (criterion) {
    case -1: return inter / (area1 + area2 - inter);
    case 0: return inter / area1;
    case 1: return inter / area2;
    case 2: {
      const bool thin = fminf(r2[2], r2[3]) / fmaxf(r2[2], r2[3]) < 0.25f;
      return thin ? inter / (area2 + fmaxf(0.f, area1 * 0.5f - inter)) : inter / (area1 + area2 - inter);
    }
    case 3: {   // 0 * IoU + 0.1 * DIoU + 0.1 * AIoU with the reference's (sic) diagonal length
      const float diag = dc + sqrtf(r1[2] * r1[2] + r2[0] * r2[0]) * 0.5f + sqrtf(r2[2] * r2[2] + r2[0] * r2[0]) * 0.5f;
      const float da = atanf(r1[2] / r1[3]) - atanf(r2[2] / r2[3]);
      const float iou = inter / (area1 + area2 - inter);
      return iou * 0.f + (1.f - dc * dc / (diag * diag)) * 0.1f + (1.f - (4.f / (float)(M_PI * M_PI)) * da * da) * 0.1f;
    }
    case 4: {   // midpoint distances
      const float dl = sqrtf((r1[0] + r1[2] * 0.5f - r2[0] - r2[2] * 0.5f) * (r1[0] + r1[2] * 0.5f - r2[0] - r2[2] * 0.5f) +
                             (r1[1] - r2[1]) * (r1[1] - r2[1]));
      const float dw = sqrtf((r1[0] - r2[0]) * (r1[0] - r2[0]) +
                             (r1[1] + r1[3] * 0.5f - r2[1] - r2[3] * 0.5f) * (r1[1] + r1[3] * 0.5f - r2[1] - r2[3] * 0.5f));
      return 2.f - (dl + dw + 1.5f * dc);
    }
    case 5: {
      const float l1 = fmaxf(r1[2], r1[3]), l2 = fmaxf(r2[2], r2[3]);
      const float da = atanf(r1[2] / r1[3]) - atanf(r2[2] / r2[3]);
      return 1.f - (fabsf(l1 - l2) + dc) / 0.5f + 0.2f * (4.f / (float)(M_PI * M_PI)) * da * da;
    }
    case 6: return 1.f - (fabsf(r1[2] - r2[2]) + fabsf(r1[3] - r2[3]) + dc) / 0.7f;
    default: return inter;
  }
}

// iou[n * K + k] = rotate_iou_eval(query[k], boxes[n])   (rotate_iou_kernel_eval, :626-664)
__global__ void k_rotate_iou(const float *__restrict__ boxes, const float *__restrict__ query, long long N, long long K,
                             int criterion, float *__restrict__ iou) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const long long n = i / K, k = i - n * K;
  float b[5], q[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) { b[j] = boxes[n * 5 + j]; q[j] = query[k * 5 + j]; }
  iou[i] = rotate_iou_eval(q, b, criterion);
}

// ---- NMS ---------------------------------------------------------------------------------------------------
// key = score as an order-preserving uint32, inverted so that an ascending sort is descending in score
__global__ void k_score_keys(const float *__restrict__ scores, long long n, uint32_t *__restrict__ key, int32_t *__restrict__ idx) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t u = __float_as_uint(scores[i]);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  key[i] = ~u;
  idx[i] = (int)i;
}
__global__ void k_gather_boxes(const float *__restrict__ boxes, const int32_t *__restrict__ order, long long m,
                               float *__restrict__ sorted) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * 5) return;
  sorted[i] = boxes[(long long)order[i / 5] * 5 + i % 5];
}
// mask[i * col_blocks + cb] bit t: sorted box i suppresses sorted box cb * 64 + t (t after i)
__global__ void __launch_bounds__(64)
k_nms_mask(const float *__restrict__ sorted, int m, float thresh, unsigned long long *__restrict__ mask) {
  pdl_sync();
  __shared__ float cb[64 * 5];
  const int row0 = blockIdx.y * 64, col0 = blockIdx.x * 64, tx = threadIdx.x;
  if (col0 < row0) return;                                 // only the upper triangle is ever read
  const int row_n = min(m - row0, 64), col_n = min(m - col0, 64);
  if (tx < col_n)
#pragma unroll
    for (int j = 0; j < 5; ++j) cb[tx * 5 + j] = sorted[(long long)(col0 + tx) * 5 + j];
  __syncthreads();
  if (tx >= row_n) return;
  float b[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) b[j] = sorted[(long long)(row0 + tx) * 5 + j];
  unsigned long long t = 0;
  for (int i = (row0 == col0 ? tx + 1 : 0); i < col_n; ++i) {
    const float iou = rotate_iou_eval(b, cb + i * 5, -1);
    if (iou > 0.f && iou >= thresh) t |= 1ull << i;
  }
  mask[(long long)(row0 + tx) * gridDim.x + blockIdx.x] = t;
}
// greedy walk in score order (nms_postprocess, :109-126): one warp, lane l owns removal words l, l + 32, ...
__global__ void __launch_bounds__(32)
k_nms_scan(const unsigned long long *__restrict__ mask, const int32_t *__restrict__ order, int m, int col_blocks,
           int post_max, int64_t *__restrict__ keep, int32_t *__restrict__ n_keep) {
  pdl_sync();
  extern __shared__ unsigned long long remv[];              // [col_blocks]
  const int lane = threadIdx.x;
  for (int j = lane; j < col_blocks; j += 32) remv[j] = 0ull;
  __syncwarp();
  int kept = 0;
  for (int i = 0; i < m && kept < post_max; ++i) {
    const int blk = i >> 6;
    if (remv[blk] & (1ull << (i & 63))) continue;           // (warp-uniform: shared memory)
    if (lane == 0) keep[kept] = order[i];
    ++kept;
    for (int j = blk + lane; j < col_blocks; j += 32) remv[j] |= mask[(long long)i * col_blocks + j];
    __syncwarp();
  }
  if (lane == 0) *n_keep = kept;
}

}  // namespace scn

using namespace scn;

extern "C" {

int scn_rotate_iou(const float *boxes, int64_t n_boxes, const float *query, int64_t n_query, int criterion,
                   float *iou_out, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_boxes == 0 || n_query == 0) return 0;
  SCN_CHECK(boxes && query && iou_out, "null pointer");
  SCN_LAUNCH(k_rotate_iou, cdiv(n_boxes * n_query, 128), 128, 0, s, boxes, query, n_boxes, n_query, criterion, iou_out);
  SCN_LAUNCHED();
  return 0;
}

int scn_rotate_nms(const float *boxes, const float *scores, int64_t n, float iou_threshold, int64_t pre_max_size,
                   int64_t post_max_size, int64_t *keep_out, int64_t *n_keep, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  SCN_CHECK(n_keep, "null argument");
  *n_keep = 0;
  if (n == 0) return 0;
  SCN_CHECK(boxes && scores && keep_out, "null pointer");
  SCN_CHECK(n < (1LL << 31), "too many boxes");
  const int m = (int)((pre_max_size > 0 && pre_max_size < n) ? pre_max_size : n);
  const int post = (int)((post_max_size > 0 && post_max_size < m) ? post_max_size : m);
  uint32_t *key = nullptr;
  int32_t *order = nullptr, *cnt = nullptr;
  float *sorted = nullptr;
  unsigned long long *mask = nullptr;
  const int col_blocks = cdiv(m, 64);
  SCN_TRY(dev_alloc_t(&key, (size_t)n, s));
  SCN_TRY(dev_alloc_t(&order, (size_t)n, s));
  SCN_TRY(dev_alloc_t(&cnt, 4, s));
  SCN_TRY(dev_alloc_t(&sorted, (size_t)m * 5, s));
  SCN_TRY(dev_alloc_t(&mask, (size_t)m * col_blocks, s));
  SCN_LAUNCH(k_score_keys, cdiv(n, 256), 256, 0, s, scores, n, key, order);
  SCN_LAUNCHED();
  SCN_TRY(radix_sort_pairs(key, order, n, 32, s, false));         // stable: equal scores keep their input order
  SCN_LAUNCH(k_gather_boxes, cdiv((long long)m * 5, 256), 256, 0, s, boxes, order, m, sorted);
  SCN_LAUNCHED();
  SCN_CUDA(cudaMemsetAsync(mask, 0, (size_t)m * col_blocks * 8, s));
  SCN_LAUNCH(k_nms_mask, dim3(col_blocks, col_blocks), 64, 0, s, sorted, m, iou_threshold, mask);
  SCN_LAUNCHED();
  SCN_LAUNCH(k_nms_scan, 1, 32, (size_t)col_blocks * 8, s, mask, order, m, col_blocks, post, keep_out, cnt);
  SCN_LAUNCHED();
  int32_t *h32 = (int32_t *)host_scratch(16);
  SCN_CUDA(cudaMemcpyAsync(h32, cnt, 4, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));      // documented read-back: the number of kept boxes (output shape)
  *n_keep = h32[0];
  dev_free(key, s); dev_free(order, s); dev_free(cnt, s); dev_free(sorted, s); dev_free(mask, s);
  return 0;
}

}  // extern "C"
