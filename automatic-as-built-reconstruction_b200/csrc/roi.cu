// roi.cu - ROIAlignRotated3D sampled STRAIGHT FROM THE SPARSE MAP.
//
// Reference: maskrcnn_benchmark/layers/roi_align_rotated_3d.py:69-85 densifies the sparse ROI map
// (SparseConvNet/sparseconvnet/tools_3d_2d.py:7-26: SparseToDense into [B,C,X,Y,Z] - 1.07 GB for a
// [1,128,256,256,32] map - cropped to max active coordinate + 1) and then runs
// maskrcnn_benchmark/csrc/cuda/ROIAlignRotated3D_cuda.cu:90-178 (forward) / :238-355 (backward), one thread
// per OUTPUT ELEMENT (n, c, ph, pw, pz): every thread recomputes the sample geometry of its bin and reads its 8
// trilinear corners of channel c - a channel-strided walk over the dense volume.
//
// Here the dense volume never exists.  A corner (yi, xi, zi) of a sample is looked up in the scale's hash grid
// (metadata.cuh): active -> its feature ROW (all channels contiguous), inactive -> contributes 0, exactly what
// the zero-filled dense tensor holds.  One warp owns a bin: the sample geometry is computed once per bin, lanes
// 0-7 probe the 8 corners in parallel, and all 32 lanes then stream the found rows' channels with 128-bit
// loads (C * 4 contiguous bytes per active corner).  The CTA's 32 bins are transposed through shared memory so
// that the output [n, C, PH, PW, PZ] is written in 128-byte runs along the bins.  Backward: the same walk,
// red.global.add of w * dOut / count into the active rows of the feature gradient (inactive corners receive
// nothing, as SparseToDense's backward drops them; the reference's atomicAdd order is equally unspecified).
//
// Arithmetic: float32 in the order of the CUDA reference (cosf / sinf, no rounding of roi extents); quirk kept:
// the forward bounds test never rejects z > zsize (`zsize > zsize`, :28) while the backward one does (:190).
#include "metadata.cuh"
#include "../../include/scn_b200.h"

namespace scn {

constexpr int ROI_BINS = 32;       // bins per CTA
constexpr int ROI_WARPS = 8;

// ext[d] = max active coordinate + 1 per axis (tools_3d_2d.py:16-18), ext[3] = max batch index + 1
__global__ void k_coord_extent(const int32_t *__restrict__ coords, long long n, int32_t *__restrict__ ext) {
  pdl_sync();
  int4 m = make_int4(0, 0, 0, 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int4 c = reinterpret_cast<const int4 *>(coords)[i];
    m.x = max(m.x, c.x + 1); m.y = max(m.y, c.y + 1); m.z = max(m.z, c.z + 1); m.w = max(m.w, c.w + 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m.x = max(m.x, __shfl_xor_sync(0xffffffffu, m.x, o)); m.y = max(m.y, __shfl_xor_sync(0xffffffffu, m.y, o));
    m.z = max(m.z, __shfl_xor_sync(0xffffffffu, m.z, o)); m.w = max(m.w, __shfl_xor_sync(0xffffffffu, m.w, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&ext[0], m.x); atomicMax(&ext[1], m.y); atomicMax(&ext[2], m.z); atomicMax(&ext[3], m.w);
  }
}

struct RoiGeom {          // per-roi constants (ROIAlignRotated3D_cuda.cu:110-146)
  float cw, ch, cz, bin_h, bin_w, bin_z, start_h, start_w, start_z, cosT, sinT;
  int gh, gw, gz, batch;
};

__device__ __forceinline__ RoiGeom roi_geom(const float *__restrict__ r, float scale, int PH, int PW, int PZ, int sampling) {
  RoiGeom g;
  g.batch = (int)r[0];
  g.cw = r[1] * scale; g.ch = r[2] * scale; g.cz = r[3] * scale;
  float rw = r[4] * scale, rh = r[5] * scale, rz = r[6] * scale;
  const float theta = (float)((double)r[7] * M_PI / 180.0);
  rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); rz = fmaxf(rz, 1.f);      // "force malformed ROIs to be 1x1"
  g.bin_h = rh / (float)PH; g.bin_w = rw / (float)PW; g.bin_z = rz / (float)PZ;
  g.gh = sampling > 0 ? sampling : (int)ceilf(rh / (float)PH);
  g.gw = sampling > 0 ? sampling : (int)ceilf(rw / (float)PW);
  g.gz = sampling > 0 ? sampling : (int)ceilf(rz / (float)PZ);
  g.start_h = -rh / 2.0f; g.start_w = -rw / 2.0f; g.start_z = -rz / 2.0f;
  g.cosT = cosf(theta); g.sinT = sinf(theta);
  return g;
}

// corner j (0..7, order w1..w8 of the reference) of the sample at (y, x, z): its dense index triple and weight.
// Returns false when the sample lies outside the volume (contributes nothing).
struct Corners { int yl, yh, xl, xh, zl, zh; float ly, lx, lz; };
template <bool BWD>
__device__ __forceinline__ bool sample_corners(float y, float x, float z, int H, int W, int Z, Corners &c) {
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W || z < -1.0f || (BWD && z > (float)Z)) return false;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  if (z <= 0.f) z = 0.f;
  c.yl = (int)y; c.xl = (int)x; c.zl = (int)z;
  if (c.yl >= H - 1) { c.yh = c.yl = H - 1; y = (float)c.yl; } else c.yh = c.yl + 1;
  if (c.xl >= W - 1) { c.xh = c.xl = W - 1; x = (float)c.xl; } else c.xh = c.xl + 1;
  if (c.zl >= Z - 1) { c.zh = c.zl = Z - 1; z = (float)c.zl; } else c.zh = c.zl + 1;
  c.ly = y - (float)c.yl; c.lx = x - (float)c.xl; c.lz = z - (float)c.zl;
  return true;
}
__device__ __forceinline__ float corner_weight(const Corners &c, int j) {
  const float wy = (j & 2) ? c.ly : 1.f - c.ly, wx = (j & 1) ? c.lx : 1.f - c.lx, wz = (j & 4) ? c.lz : 1.f - c.lz;
  return wy * wx * wz;
}

struct RoiArgs {
  const float *feats;              // [n_rows, C]
  const float *rois;               // [n_rois, 8]
  const int32_t *ext;              // device: H, W, Z (crop extents), batch
  const uint64_t *hk; const int32_t *hv; uint32_t hmask;
  float scale;
  int C, PH, PW, PZ, sampling, n_rois;
};

// FWD: out[n][c][bin] = mean over the bin's samples of the trilinear value; BWD: dfeat[row][c] += w dOut / count
template <bool BWD>
__global__ void __launch_bounds__(ROI_WARPS * 32)
k_roi_align(RoiArgs a, float *__restrict__ out /* FWD: output; BWD: feature gradient */,
            const float *__restrict__ dout) {
  pdl_sync();
  extern __shared__ float tile[];                 // [ROI_BINS][C + 1]
  const int n = blockIdx.x, bin0 = blockIdx.y * ROI_BINS;
  const int bins = a.PH * a.PW * a.PZ, C = a.C, ldt = C + 1;
  const int nb = min(ROI_BINS, bins - bin0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = a.ext[0], W = a.ext[1], Z = a.ext[2];
  if (BWD) {   // dOut[n][c][bin0 .. bin0+nb) -> tile[b][c] (reads along the bins)
    for (int i = threadIdx.x; i < C * nb; i += ROI_WARPS * 32) {
      const int c = i / nb, b = i - c * nb;
      tile[b * ldt + c] = dout[((long long)n * C + c) * bins + bin0 + b];
    }
    __syncthreads();
  }
  const RoiGeom g = roi_geom(a.rois + (long long)n * 8, a.scale, a.PH, a.PW, a.PZ, a.sampling);
  const float inv_count = 1.f / (float)(g.gh * g.gw * g.gz);
  const float count = (float)(g.gh * g.gw * g.gz);
  const bool empty = H <= 0 || W <= 0 || Z <= 0;
  const bool vec = (C & 127) == 0;
  for (int b = warp; b < nb; b += ROI_WARPS) {
    const int bin = bin0 + b;
    const int pz = bin % a.PZ, pw = (bin / a.PZ) % a.PW, ph = bin / (a.PZ * a.PW);
    float *trow = tile + b * ldt;
    // forward: the bin's channels accumulate in registers, 512 channels per pass over the samples (one pass for
    // every width the backbone has): lane l owns channels c0 + 128 q + 4 l .. + 3 (128-bit loads) or c0 + 32 q + l
    for (int c0 = 0; c0 < (BWD ? 1 : C); c0 += 512) {
      float acc[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = 0.f;
      if (!empty)
        for (int iy = 0; iy < g.gh; ++iy) {
          const float yy = g.start_h + ph * g.bin_h + ((float)iy + .5f) * g.bin_h / (float)g.gh;
          for (int ix = 0; ix < g.gw; ++ix) {
            const float xx = g.start_w + pw * g.bin_w + ((float)ix + .5f) * g.bin_w / (float)g.gw;
            for (int iz = 0; iz < g.gz; ++iz) {
              const float zz = g.start_z + pz * g.bin_z + ((float)iz + .5f) * g.bin_z / (float)g.gz;
              const float x = xx * g.cosT + yy * g.sinT + g.cw;
              const float y = yy * g.cosT - xx * g.sinT + g.ch;
              const float z = zz + g.cz;
              Corners cr;
              if (!sample_corners<BWD>(y, x, z, H, W, Z, cr)) continue;       // (warp-uniform)
              // lanes 0-7 probe one corner each: dense (H, W, Z) index = sparse (x, y, z) coordinate
              int row = -1;
              if (lane < 8) {
                const int yi = (lane & 2) ? cr.yh : cr.yl, xi = (lane & 1) ? cr.xh : cr.xl, zi = (lane & 4) ? cr.zh : cr.zl;
                if (coord_ok(yi, xi, zi)) row = hash_find(a.hk, a.hv, a.hmask, pack_key(yi, xi, zi, g.batch));
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int rj = __shfl_sync(0xffffffffu, row, j);
                if (rj < 0) continue;
                const float w = corner_weight(cr, j);
                if (BWD) {
                  float *grow = out + (long long)rj * C;
                  for (int c = lane; c < C; c += 32) atomicAdd(grow + c, trow[c] * w / count);
                } else {
                  const float *frow = a.feats + (long long)rj * C + c0;
                  if (vec) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                      if (c0 + q * 128 < C) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(frow + q * 128 + lane * 4));
                        acc[4 * q] += w * v.x; acc[4 * q + 1] += w * v.y; acc[4 * q + 2] += w * v.z; acc[4 * q + 3] += w * v.w;
                      }
                  } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                      if (c0 + q * 32 + lane < C) acc[q] += w * __ldg(frow + q * 32 + lane);
                  }
                }
              }
            }
          }
        }
      if (!BWD) {
        if (vec) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (c0 + q * 128 < C) {
#pragma unroll
              for (int e = 0; e < 4; ++e) trow[c0 + q * 128 + lane * 4 + e] = acc[4 * q + e];
            }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (c0 + q * 32 + lane < C) trow[c0 + q * 32 + lane] = acc[q];
        }
      }
    }
  }
  if (!BWD) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * nb; i += ROI_WARPS * 32) {
      const int c = i / nb, b = i - c * nb;
      out[((long long)n * C + c) * bins + bin0 + b] = tile[b * ldt + c] * inv_count;
    }
  }
}

}  // namespace scn

using namespace scn;

extern "C" {

static int roi_setup(scn_metadata_t *m, const int64_t *ss, const float *rois, int64_t n_rois, int64_t C,
                     float spatial_scale, const int64_t *pooled, int sampling_ratio, cudaStream_t s, RoiArgs *a,
                     Grid **grid, int32_t **ext) {
  SCN_CHECK(m && ss && pooled, "null argument");
  SCN_CHECK(C > 0 && C <= 4096, "ROIAlignRotated3D: %lld planes not in 1..4096", (long long)C);
  SCN_CHECK(pooled[0] > 0 && pooled[1] > 0 && pooled[2] > 0, "ROIAlignRotated3D: bad output size");
  Grid *g = find_grid(m, ss);
  SCN_CHECK(g, "no active sites at spatial size [%lld,%lld,%lld]", (long long)ss[0], (long long)ss[1], (long long)ss[2]);
  *grid = g;
  SCN_TRY(dev_alloc_t(ext, 4, s));
  SCN_CUDA(cudaMemsetAsync(*ext, 0, 16, s));
  if (g->n_active > 0) {
    int blocks = cdiv(g->n_active, 256);
    if (blocks > num_sms() * 4) blocks = num_sms() * 4;
    SCN_LAUNCH(k_coord_extent, blocks, 256, 0, s, g->coords, g->n_active, *ext);
    SCN_LAUNCHED();
  }
  a->rois = rois; a->ext = *ext; a->hk = g->hkeys; a->hv = g->hvals; a->hmask = g->hcap - 1;
  a->scale = spatial_scale; a->C = (int)C; a->PH = (int)pooled[0]; a->PW = (int)pooled[1]; a->PZ = (int)pooled[2];
  a->sampling = sampling_ratio; a->n_rois = (int)n_rois;
  return 0;
}

// ext[0..2] = max active coordinate + 1 per axis, ext[3] = max batch index + 1 of the grid at `ss` - what
// tools_3d_2d.py:16-18 derives on the host from a copy of every location; here one 16-byte read-back
int scn_grid_extent(scn_metadata_t *m, const int64_t *ss, int64_t ext_out[4], void *stream) {
  SCN_CHECK(m && ss && ext_out, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  ext_out[0] = ext_out[1] = ext_out[2] = ext_out[3] = 0;
  Grid *g = find_grid(m, ss);
  if (!g || g->n_active == 0) return 0;
  int32_t *ext = nullptr;
  SCN_TRY(dev_alloc_t(&ext, 4, s));
  SCN_CUDA(cudaMemsetAsync(ext, 0, 16, s));
  int blocks = cdiv(g->n_active, 256);
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  SCN_LAUNCH(k_coord_extent, blocks, 256, 0, s, g->coords, g->n_active, ext);
  SCN_LAUNCHED();
  int32_t *h = (int32_t *)host_scratch(2);
  SCN_CHECK(h, "pinned host scratch unavailable");
  SCN_CUDA(cudaMemcpyAsync(h, ext, 16, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < 4; ++i) ext_out[i] = h[i];
  dev_free(ext, s);
  return 0;
}

int scn_roi_align_rotated_3d_forward(scn_metadata_t *m, const int64_t *ss, const float *feats, int64_t n_planes,
                                     const float *rois, int64_t n_rois, float spatial_scale, const int64_t *pooled,
                                     int sampling_ratio, float *out, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_rois == 0) return 0;
  SCN_CHECK(rois && out, "null argument");
  RoiArgs a;
  Grid *g = nullptr;
  int32_t *ext = nullptr;
  SCN_TRY(roi_setup(m, ss, rois, n_rois, n_planes, spatial_scale, pooled, sampling_ratio, s, &a, &g, &ext));
  SCN_CHECK(feats || g->n_active == 0, "null feature pointer");
  a.feats = feats;
  const int bins = a.PH * a.PW * a.PZ;
  const size_t sm = (size_t)ROI_BINS * (a.C + 1) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    SCN_CUDA(cudaFuncSetAttribute(k_roi_align<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SCN_CUDA(cudaFuncSetAttribute(k_roi_align<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  prof_begin(PROF_IO, s);
  SCN_LAUNCH((k_roi_align<false>), dim3((unsigned)n_rois, cdiv(bins, ROI_BINS)), ROI_WARPS * 32, sm, s, a, out, nullptr);
  SCN_LAUNCHED();
  prof_end(PROF_IO, s, 4.0 * ((double)g->n_active * a.C + (double)n_rois * a.C * bins) + 32.0 * n_rois, 0);
  dev_free(ext, s);
  return 0;
}

int scn_roi_align_rotated_3d_backward(scn_metadata_t *m, const int64_t *ss, const float *d_out, int64_t n_planes,
                                      const float *rois, int64_t n_rois, float spatial_scale, const int64_t *pooled,
                                      int sampling_ratio, float *d_feats, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  SCN_CHECK(m && ss, "null argument");
  Grid *g0 = find_grid(m, ss);
  SCN_CHECK(g0, "no active sites at spatial size [%lld,%lld,%lld]", (long long)ss[0], (long long)ss[1], (long long)ss[2]);
  if (g0->n_active > 0) {
    SCN_CHECK(d_feats, "null gradient pointer");
    SCN_CUDA(cudaMemsetAsync(d_feats, 0, (size_t)g0->n_active * n_planes * 4, s));
  }
  if (n_rois == 0 || g0->n_active == 0) return 0;
  SCN_CHECK(rois && d_out, "null argument");
  RoiArgs a;
  Grid *g = nullptr;
  int32_t *ext = nullptr;
  SCN_TRY(roi_setup(m, ss, rois, n_rois, n_planes, spatial_scale, pooled, sampling_ratio, s, &a, &g, &ext));
  a.feats = nullptr;
  const int bins = a.PH * a.PW * a.PZ;
  const size_t sm = (size_t)ROI_BINS * (a.C + 1) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    SCN_CUDA(cudaFuncSetAttribute(k_roi_align<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  prof_begin(PROF_IO, s);
  SCN_LAUNCH((k_roi_align<true>), dim3((unsigned)n_rois, cdiv(bins, ROI_BINS)), ROI_WARPS * 32, sm, s, a, d_feats, d_out);
  SCN_LAUNCHED();
  prof_end(PROF_IO, s, 4.0 * ((double)g->n_active * a.C + (double)n_rois * a.C * bins) + 32.0 * n_rois, 0);
  dev_free(ext, s);
  return 0;
}

}  // extern "C"
