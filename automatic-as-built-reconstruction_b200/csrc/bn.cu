// bn.cu - BatchNormalization fused with (Leaky)ReLU over the active rows.
//
// Reference: SparseConvNet/sparseconvnet/SCN/CPU/BatchNormalization.cpp:13-107 (math) and
// CUDA/BatchNormalization.cu:14-184 (which runs on <= 16 thread blocks).  Here: a full-grid
// statistics pass with 128-bit coalesced loads and per-thread fp32 / cross-thread fp64
// accumulation folded into 2C fp64 sums, and a vectorised elementwise apply whose blocks derive
// the per-channel coefficients themselves (no separate finalize launch).
#include "common.cuh"
#include "../../include/scn_b200.h"
#include <cstdlib>
#include <mutex>
#include <vector>

namespace scn {

constexpr int BN_T = 256;
// The 2C fp64 column sums exist in G = 8 replicas, [G][2C]: block b of a statistics kernel adds into replica
// b % G and every apply block sums the replicas (fixed order) when it derives the coefficients.  With one copy, the
// <= 4 blocks per SM x 2C atomics of a launch all land on the 2C x 8 bytes of a few 128-byte lines and drain through
// the L2 slices that own them one at a time: ncu showed the SMs of a statistics kernel idle for 7 us (C = 32) to 11 us
// (C = 64) of a 14 - 42 us launch, growing with C and with the number of blocks (tools/bn_probe.py).
constexpr int BN_G_MAX = 32;
static int bn_replicas() {            // SCN_B200_BN_REPLICAS (1 ... 32, default 8: one unrolled batch of loads in bn_acc)
  static const int g = [] {
    const char *e = getenv("SCN_B200_BN_REPLICAS");
    int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > BN_G_MAX ? BN_G_MAX : v);
  }();
  return g;
}

// y = lrelu(fmaf(x, w, b)) with w = invstd * gamma, b = fma(-mean, w, beta): ONE definition for the forward apply and
// for the backward pass, which recomputes the sign of the pre-activation from x instead of reading y back (the
// same fp32 operations on the same saved mean / invstd give the same bits, hence the same ReLU mask)
__device__ __forceinline__ void bn_affine(float mean, float invstd, float gamma, float beta, float &w, float &b) {
  w = invstd * gamma;
  b = __fmaf_rn(-mean, w, beta);
}

// ---- statistics: q0 = sum a, q1 = sum b over rows, per channel ---------------------------
// forward:  a = x,  b = x*x
// backward: a = d,  b = (x-mean)*d   with d = dy * (y > 0 ? 1 : leak)
template <bool BWD>
__device__ __forceinline__ void bn_terms(float x, float y, float dy, float mean, float leak,
                                         float &a, float &b) {
  if (BWD) {
    const float d = dy * (y > 0.f ? 1.f : leak);   // y: the BN output, or the recomputed pre-activation (same sign)
    a = d;
    b = (x - mean) * d;
  } else {
    a = x;
    b = x * x;
  }
}

// vector path: C % 4 == 0 and 1024 % C == 0; thread owns one channel quad for all its rows
// Yo == nullptr (backward): the ReLU mask comes from the recomputed pre-activation fmaf(x, w, b) with
// (w, b) = bn_affine(mean, invstd, gamma, beta) - one tensor less to read
template <bool BWD>
__global__ void __launch_bounds__(BN_T)
k_bn_stats_vec(const float *__restrict__ X, const float *__restrict__ Yo,
               const float *__restrict__ dY, const float *__restrict__ mean, float leak,
               long long n, int C, double *__restrict__ acc, int G, const float *__restrict__ invstd = nullptr,
               const float *__restrict__ gamma = nullptr, const float *__restrict__ beta = nullptr) {
  pdl_sync();
  __shared__ double red[2][BN_T * 4];
  const int qpr = C >> 2;               // quads per row
  const int rpb = BN_T / qpr;           // rows per block iteration (>= 1 since C <= 1024)
  const int cq = threadIdx.x % qpr, rs = threadIdx.x / qpr;
  float4 s0 = make_float4(0, 0, 0, 0), s1 = make_float4(0, 0, 0, 0);
  float4 mu = make_float4(0, 0, 0, 0), aw = mu, ab = mu;
  if (BWD) mu = reinterpret_cast<const float4 *>(mean)[cq];
  const bool recompute = BWD && Yo == nullptr;
  if (recompute) {
    const float4 is = reinterpret_cast<const float4 *>(invstd)[cq];
    const int c = cq * 4;
    bn_affine(mu.x, is.x, gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, aw.x, ab.x);
    bn_affine(mu.y, is.y, gamma ? gamma[c + 1] : 1.f, beta ? beta[c + 1] : 0.f, aw.y, ab.y);
    bn_affine(mu.z, is.z, gamma ? gamma[c + 2] : 1.f, beta ? beta[c + 2] : 0.f, aw.z, ab.z);
    bn_affine(mu.w, is.w, gamma ? gamma[c + 3] : 1.f, beta ? beta[c + 3] : 0.f, aw.w, ab.w);
  }
  if (rs < rpb) {
#pragma unroll 4   // independent 128-bit loads of 4 row slots in flight per thread
    for (long long r = (long long)blockIdx.x * rpb + rs; r < n; r += (long long)gridDim.x * rpb) {
      const long long o = r * qpr + cq;
      const float4 x = __ldg(reinterpret_cast<const float4 *>(X) + o);
      float4 y = x, d = x;
      if (BWD) {
        if (recompute)
          y = make_float4(fmaf(x.x, aw.x, ab.x), fmaf(x.y, aw.y, ab.y), fmaf(x.z, aw.z, ab.z), fmaf(x.w, aw.w, ab.w));
        else
          y = __ldg(reinterpret_cast<const float4 *>(Yo) + o);
        d = __ldg(reinterpret_cast<const float4 *>(dY) + o);
      }
      float a, b;
      bn_terms<BWD>(x.x, y.x, d.x, mu.x, leak, a, b); s0.x += a; s1.x += b;
      bn_terms<BWD>(x.y, y.y, d.y, mu.y, leak, a, b); s0.y += a; s1.y += b;
      bn_terms<BWD>(x.z, y.z, d.z, mu.z, leak, a, b); s0.z += a; s1.z += b;
      bn_terms<BWD>(x.w, y.w, d.w, mu.w, leak, a, b); s0.w += a; s1.w += b;
    }
  }
  const int t4 = threadIdx.x * 4;
  red[0][t4] = s0.x; red[0][t4 + 1] = s0.y; red[0][t4 + 2] = s0.z; red[0][t4 + 3] = s0.w;
  red[1][t4] = s1.x; red[1][t4 + 1] = s1.y; red[1][t4 + 2] = s1.z; red[1][t4 + 3] = s1.w;
  __syncthreads();
  // thread c (< C) sums the rpb row-slots of channel c: element index = rs*C + c
  double *__restrict__ mine = acc + (size_t)(blockIdx.x % G) * 2 * C;
  for (int c = threadIdx.x; c < C; c += BN_T) {
    double a = 0, b = 0;
    for (int j = 0; j < rpb; ++j) { a += red[0][j * C + c]; b += red[1][j * C + c]; }
    atomicAdd(&mine[c], a);        // fp64: the order of the block sums inside a replica is immaterial
    atomicAdd(&mine[C + c], b);
  }
}

// generic path: block (32,8); x = channel, y = row slot
template <bool BWD>
__global__ void __launch_bounds__(BN_T)
k_bn_stats_gen(const float *__restrict__ X, const float *__restrict__ Yo,
               const float *__restrict__ dY, const float *__restrict__ mean, float leak,
               long long n, int C, double *__restrict__ acc, int G) {
  pdl_sync();
  __shared__ double red[2][8][33];
  const int c = blockIdx.y * 32 + threadIdx.x;
  double s0 = 0, s1 = 0;
  if (c < C) {
    const float mu = BWD ? mean[c] : 0.f;
    for (long long r = (long long)blockIdx.x * 8 + threadIdx.y; r < n; r += (long long)gridDim.x * 8) {
      const float x = X[r * C + c];
      float a, b;
      bn_terms<BWD>(x, BWD ? Yo[r * C + c] : 0.f, BWD ? dY[r * C + c] : 0.f, mu, leak, a, b);
      s0 += a;
      s1 += b;
    }
  }
  red[0][threadIdx.y][threadIdx.x] = s0;
  red[1][threadIdx.y][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double a = 0, b = 0;
    for (int j = 0; j < 8; ++j) { a += red[0][j][threadIdx.x]; b += red[1][j][threadIdx.x]; }
    double *__restrict__ mine = acc + (size_t)(blockIdx.x % G) * 2 * C;
    atomicAdd(&mine[c], a);        // fp64: the order of the block sums inside a replica is immaterial
    atomicAdd(&mine[C + c], b);
  }
}

// column sum c (s = 0) / second sum (s = 1) over the replicas, in replica order
__device__ __forceinline__ double bn_acc(const double *__restrict__ acc, int G, int C, int c, int s) {
  double v = 0;
#pragma unroll 8
  for (int g = 0; g < G; ++g) v += acc[(size_t)g * 2 * C + s * C + c];
  return v;
}

// ---- coefficients -------------------------------------------------------------------------
// Every apply block derives the per-channel coefficients from the fp64 sums itself (2C loads);
// block 0 also publishes the saved / running statistics or the parameter gradients.
// forward train (CPU/BatchNormalization.cpp:19-40): coef = [scale, shift]
__device__ __forceinline__ void bn_fwd_coef(const double *__restrict__ acc, int G, long long n, int C, int c,
                                            float *save_mean, float *save_invstd, float *running_mean,
                                            float *running_var, const float *weight, const float *bias,
                                            float eps, float momentum, int train, bool publish,
                                            float *coef) {
  float mean, invstd;
  if (train) {
    const double sum = bn_acc(acc, G, C, c, 0), sq = bn_acc(acc, G, C, c, 1);
    const double dmean = sum / (double)n;
    const double s = sq - dmean * dmean * (double)n;  // sum of squared deviations
    mean = (float)dmean;
    // train == 2: evaluation with the statistics of the CURRENT batch and the unbiased variance - what the
    // reference's Python layer does in eval mode with track_running_stats=False (batchNormalization.py:51-56:
    // features.mean(0) / features.var(0) handed to the C++ eval path); the running buffers are left alone
    invstd = powf((float)(s / (double)(train == 2 ? n - 1 : n)) + eps, -0.5f);
    if (publish && train == 1) {
      running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * mean;
      running_var[c] = momentum * running_var[c] + (1.f - momentum) * (float)(s / (double)(n - 1));
    }
  } else {  // :41-46 statistics are the buffers as passed
    mean = running_mean[c];
    invstd = powf(running_var[c] + eps, -0.5f);
  }
  if (publish) {
    save_mean[c] = mean;
    save_invstd[c] = invstd;
  }
  float w, b;
  bn_affine(mean, invstd, weight ? weight[c] : 1.f, bias ? bias[c] : 0.f, w, b);
  coef[c] = w;
  coef[C + c] = b;
}

// backward (:86-106): coef = [gradMean, k, invstd*gamma]
__device__ __forceinline__ void bn_bwd_coef(const double *__restrict__ acc, int G, long long n, int C, int c,
                                            const float *save_invstd, const float *weight,
                                            float *d_weight, float *d_bias, bool publish, float *coef) {
  const double gsum = bn_acc(acc, G, C, c, 0), dotp = bn_acc(acc, G, C, c, 1);
  const float invstd = save_invstd[c];
  if (publish) {
    if (d_bias) d_bias[c] = (float)gsum;
    if (d_weight) d_weight[c] = (float)dotp * invstd;
  }
  coef[c] = (float)(gsum / (double)n);
  coef[C + c] = (float)dotp * invstd * invstd / (float)n;
  coef[2 * C + c] = invstd * (weight ? weight[c] : 1.f);
}

// ---- apply -----------------------------------------------------------------------------
__device__ __forceinline__ float lrelu(float v, float leak) { return v > 0.f ? v : v * leak; }

struct BnFwdArgs {
  double *zero_buf;   // the OTHER ping-pong sum buffer: block 0 clears it for the next BN call
  int zero_n;
  const double *acc;
  int replicas;
  float *save_mean, *save_invstd, *running_mean, *running_var;
  const float *weight, *bias;
  float eps, momentum;
  int train;
};

template <bool VEC>
__global__ void __launch_bounds__(BN_T)
k_bn_fwd_apply(const float *__restrict__ X, float *__restrict__ Y, BnFwdArgs a, float leak,
               long long n, int C) {
  pdl_sync();
  extern __shared__ float coef[];  // [2C]
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < a.zero_n; i += BN_T) a.zero_buf[i] = 0.0;
  for (int c = threadIdx.x; c < C; c += BN_T)
    bn_fwd_coef(a.acc, a.replicas, n, C, c, a.save_mean, a.save_invstd, a.running_mean, a.running_var, a.weight,
                a.bias, a.eps, a.momentum, a.train, blockIdx.x == 0, coef);
  __syncthreads();
  const long long total = n * C;
  const long long st = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    const long long nq = total >> 2;
    auto apply4 = [&](long long q, const float4 x) {
      const int c = (int)((q << 2) % C);
      const float4 w = *reinterpret_cast<const float4 *>(coef + c);
      const float4 b = *reinterpret_cast<const float4 *>(coef + C + c);
      float4 y;
      y.x = lrelu(fmaf(x.x, w.x, b.x), leak);
      y.y = lrelu(fmaf(x.y, w.y, b.y), leak);
      y.z = lrelu(fmaf(x.z, w.z, b.z), leak);
      y.w = lrelu(fmaf(x.w, w.w, b.w), leak);
      reinterpret_cast<float4 *>(Y)[q] = y;
    };
    // four independent 128-bit loads in flight per thread (one per iteration left a full grid at 32 KB in flight per SM)
    for (; i + 3 * st < nq; i += 4 * st) {
      const float4 x0 = __ldg(reinterpret_cast<const float4 *>(X) + i);
      const float4 x1 = __ldg(reinterpret_cast<const float4 *>(X) + i + st);
      const float4 x2 = __ldg(reinterpret_cast<const float4 *>(X) + i + 2 * st);
      const float4 x3 = __ldg(reinterpret_cast<const float4 *>(X) + i + 3 * st);
      apply4(i, x0); apply4(i + st, x1); apply4(i + 2 * st, x2); apply4(i + 3 * st, x3);
    }
    for (; i < nq; i += st) apply4(i, __ldg(reinterpret_cast<const float4 *>(X) + i));
  } else {
    for (; i < total; i += st) {
      const int c = (int)(i % C);
      Y[i] = lrelu(fmaf(X[i], coef[c], coef[C + c]), leak);
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(BN_T)
k_bn_bwd_apply(const float *__restrict__ X, const float *__restrict__ Yo,
               const float *__restrict__ dY, float *dX /* may alias R */,
               const float *__restrict__ mean, const double *__restrict__ acc,
               const float *__restrict__ save_invstd, const float *__restrict__ weight,
               float *d_weight, float *d_bias, float leak, long long n, int C, double *zero_buf, int zero_n,
               const float *R, const float *__restrict__ beta, int G) {
  pdl_sync();
  extern __shared__ float coef[];  // [3C] then mean [C], then (Yo == nullptr) the forward's [w, b] [2C]
  float *smean = coef + 3 * C;
  float *aff = coef + 4 * C;
  const bool recompute = Yo == nullptr;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < zero_n; i += BN_T) zero_buf[i] = 0.0;
  for (int c = threadIdx.x; c < C; c += BN_T) {
    bn_bwd_coef(acc, G, n, C, c, save_invstd, weight, d_weight, d_bias, blockIdx.x == 0, coef);
    smean[c] = mean[c];
    if (recompute) bn_affine(mean[c], save_invstd[c], weight ? weight[c] : 1.f, beta ? beta[c] : 0.f, aff[c], aff[C + c]);
  }
  __syncthreads();
  const long long total = n * C;
  const long long st = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    const long long nq = total >> 2;
    auto bwd4 = [&](long long q, const float4 x, const float4 yo, const float4 d, const float4 r) {
      const int c = (int)((q << 2) % C);
      float4 y = yo;
      if (recompute) {
        const float4 w = *reinterpret_cast<const float4 *>(aff + c);
        const float4 b = *reinterpret_cast<const float4 *>(aff + C + c);
        y = make_float4(fmaf(x.x, w.x, b.x), fmaf(x.y, w.y, b.y), fmaf(x.z, w.z, b.z), fmaf(x.w, w.w, b.w));
      }
      const float4 mu = *reinterpret_cast<const float4 *>(smean + c);
      const float4 gm = *reinterpret_cast<const float4 *>(coef + c);
      const float4 kk = *reinterpret_cast<const float4 *>(coef + C + c);
      const float4 sc = *reinterpret_cast<const float4 *>(coef + 2 * C + c);
      float4 o;
      o.x = (d.x * (y.x > 0.f ? 1.f : leak) - gm.x - (x.x - mu.x) * kk.x) * sc.x;
      o.y = (d.y * (y.y > 0.f ? 1.f : leak) - gm.y - (x.y - mu.y) * kk.y) * sc.y;
      o.z = (d.z * (y.z > 0.f ? 1.f : leak) - gm.z - (x.z - mu.z) * kk.z) * sc.z;
      o.w = (d.w * (y.w > 0.f ? 1.f : leak) - gm.w - (x.w - mu.w) * kk.w) * sc.w;
      if (R) { o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }   // second gradient of the same value (skip connection), may alias dX
      reinterpret_cast<float4 *>(dX)[q] = o;
    };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ld = [&](const float *p, long long q) { return p ? __ldg(reinterpret_cast<const float4 *>(p) + q) : z4; };
    auto ldr = [&](long long q) { return R ? reinterpret_cast<const float4 *>(R)[q] : z4; };   // (R may alias dX: plain load)
    // two elements per iteration: 4 - 8 independent 128-bit loads in flight per thread
    for (; i + st < nq; i += 2 * st) {
      const float4 xa = ld(X, i), xb = ld(X, i + st);
      const float4 ya = ld(recompute ? nullptr : Yo, i), yb = ld(recompute ? nullptr : Yo, i + st);
      const float4 da = ld(dY, i), db = ld(dY, i + st);
      const float4 ra = ldr(i), rb = ldr(i + st);
      bwd4(i, xa, ya, da, ra);
      bwd4(i + st, xb, yb, db, rb);
    }
    for (; i < nq; i += st) bwd4(i, ld(X, i), ld(recompute ? nullptr : Yo, i), ld(dY, i), ldr(i));
  } else {
    for (; i < total; i += st) {
      const int c = (int)(i % C);
      const float yv = recompute ? fmaf(X[i], aff[c], aff[C + c]) : Yo[i];
      const float d = dY[i] * (yv > 0.f ? 1.f : leak);
      dX[i] = (d - coef[c] - (X[i] - smean[c]) * coef[C + c]) * coef[2 * C + c] + (R ? R[i] : 0.f);
    }
  }
}

static bool vec_ok(long long n, int C, const void *a, const void *b, const void *c, const void *d) {
  auto al = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
  return C % 4 == 0 && C <= 1024 && 1024 % C == 0 && al(a) && al(b) && al(c) && al(d);
}

static int stats_grid(long long n, int rows_per_iter) {
  long long g = (n + (long long)rows_per_iter * 4 - 1) / ((long long)rows_per_iter * 4);
  const long long cap = (long long)num_sms() * 4;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Per-stream ping-pong pair of fp64 sum buffers.  A BN call accumulates into one and its apply kernel
// clears the other (last read by the previous call's apply kernel, complete by stream order), so no
// memset is launched per call.
struct BnState { double *buf[2]; int used[2]; int k; };
static std::mutex g_bn_mu;
static std::vector<std::pair<StreamKey, BnState *>> g_bn;
constexpr int BN_MAX_C = 4096;

static int bn_buffers(cudaStream_t s, int C, double **cur, double **other, int *other_used) {
  std::lock_guard<std::mutex> lk(g_bn_mu);
  BnState *st = nullptr;
  for (auto &e : g_bn)
    if (e.first == stream_key(s)) { st = e.second; break; }
  if (!st) {
    st = new BnState();
    for (int i = 0; i < 2; ++i) {
      SCN_CUDA(cudaMalloc((void **)&st->buf[i], (size_t)BN_G_MAX * 2 * BN_MAX_C * sizeof(double)));
      SCN_CUDA(cudaMemset(st->buf[i], 0, (size_t)BN_G_MAX * 2 * BN_MAX_C * sizeof(double)));
      st->used[i] = 0;
    }
    st->k = 0;
    g_bn.emplace_back(stream_key(s), st);
  }
  const int a = st->k & 1, b = a ^ 1;
  *cur = st->buf[a];
  *other = st->buf[b];
  *other_used = st->used[b];
  st->used[a] = bn_replicas() * 2 * C;
  st->used[b] = 0;
  st->k++;
  return 0;
}

static int apply_grid(long long work) {
  long long g = (work + BN_T - 1) / BN_T;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace scn

using namespace scn;

extern "C" {

int scn_batchnorm_forward(const float *in, float *out, float *save_mean, float *save_invstd,
                          float *running_mean, float *running_var, const float *weight,
                          const float *bias, float eps, float momentum, int train, float leakiness,
                          int64_t n, int64_t C64, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int C = (int)C64;
  SCN_CHECK(C > 0 && C <= 4096 && save_mean && save_invstd && running_mean && running_var, "bad BN arguments");
  SCN_CHECK(train >= 0 && train <= 2, "BN mode %d not in {0 eval, 1 train, 2 eval with batch statistics}", train);
  if (n == 0) return 0;
  SCN_CHECK(in && out, "null feature pointer");
  const bool vec = vec_ok(n, C, in, out, in, out);
  prof_begin(PROF_BN, s);
  double *acc = nullptr, *other = nullptr;
  int other_used = 0;
  if (train) {
    SCN_TRY(bn_buffers(s, C, &acc, &other, &other_used));
    if (vec)
      SCN_LAUNCH((k_bn_stats_vec<false>), stats_grid(n, BN_T / (C / 4)), BN_T, 0, s, in, nullptr, nullptr, nullptr, 0.f, n, C, acc,
                 bn_replicas(), nullptr, nullptr, nullptr);
    else
      SCN_LAUNCH((k_bn_stats_gen<false>), dim3(stats_grid(n, 8), cdiv(C, 32)), dim3(32, 8), 0, s, in, nullptr, nullptr, nullptr,
                                                                                     0.f, n, C, acc, bn_replicas());
    SCN_LAUNCHED();
  }
  BnFwdArgs a{other, other_used, acc, bn_replicas(), save_mean, save_invstd, running_mean, running_var, weight, bias, eps, momentum, train};
  const long long total = (long long)n * C;
  const size_t sm = (size_t)2 * C * sizeof(float);
  if (vec) SCN_LAUNCH((k_bn_fwd_apply<true>), apply_grid(total / 4), BN_T, sm, s, in, out, a, leakiness, n, C);
  else SCN_LAUNCH((k_bn_fwd_apply<false>), apply_grid(total), BN_T, sm, s, in, out, a, leakiness, n, C);
  SCN_LAUNCHED();
  prof_end(PROF_BN, s, (train ? 3.0 : 2.0) * 4.0 * (double)n * C, 0);  // 3 n C s
  return 0;
}

int scn_batchnorm_backward(const float *in, float *d_in, const float *out, const float *d_out,
                           const float *save_mean, const float *save_invstd, const float *weight,
                           float *d_weight, float *d_bias, float leakiness, int64_t n, int64_t C64,
                           void *stream) {
  return scn_batchnorm_backward_add(in, d_in, out, d_out, save_mean, save_invstd, weight, d_weight, d_bias,
                                    leakiness, n, C64, nullptr, stream);
}

int scn_batchnorm_backward_add(const float *in, float *d_in, const float *out, const float *d_out,
                               const float *save_mean, const float *save_invstd, const float *weight,
                               float *d_weight, float *d_bias, float leakiness, int64_t n, int64_t C64,
                               const float *residual, void *stream) {
  SCN_CHECK(out || n == 0, "null feature pointer");
  return scn_batchnorm_backward_fused(in, d_in, out, d_out, save_mean, save_invstd, weight, nullptr, 0, d_weight, d_bias,
                                      leakiness, n, C64, residual, stream);
}

int scn_batchnorm_backward_fused(const float *in, float *d_in, const float *out, const float *d_out,
                                 const float *save_mean, const float *save_invstd, const float *weight,
                                 const float *bias, int recompute_mask, float *d_weight, float *d_bias,
                                 float leakiness, int64_t n, int64_t C64, const float *residual, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int C = (int)C64;
  SCN_CHECK(C > 0 && C <= 4096 && save_mean && save_invstd, "bad BN arguments");
  if (n == 0) {
    if (d_weight) SCN_CUDA(cudaMemsetAsync(d_weight, 0, (size_t)C * 4, s));
    if (d_bias) SCN_CUDA(cudaMemsetAsync(d_bias, 0, (size_t)C * 4, s));
    return 0;
  }
  SCN_CHECK(in && d_in && d_out, "null feature pointer");
  const bool vec = vec_ok(n, C, in, d_in, out, d_out) && (((uintptr_t)save_mean & 15) == 0) &&
                   (((uintptr_t)save_invstd & 15) == 0) && (((uintptr_t)residual & 15) == 0);
  // recompute_mask: `weight` / `bias` are the forward's gamma / beta (NULL = 1 / 0), so the sign of the
  // pre-activation can be recomputed from `in` and `out` need not be read (vector path only)
  if (!(recompute_mask && vec)) SCN_CHECK(out, "null feature pointer");
  else out = nullptr;
  prof_begin(PROF_BN, s);
  double *acc = nullptr, *other = nullptr;
  int other_used = 0;
  SCN_TRY(bn_buffers(s, C, &acc, &other, &other_used));
  if (vec)
    SCN_LAUNCH((k_bn_stats_vec<true>), stats_grid(n, BN_T / (C / 4)), BN_T, 0, s, in, out, d_out, save_mean, leakiness, n, C, acc,
                                                                        bn_replicas(), save_invstd, weight, bias);
  else
    SCN_LAUNCH((k_bn_stats_gen<true>), dim3(stats_grid(n, 8), cdiv(C, 32)), dim3(32, 8), 0, s, in, out, d_out, save_mean,
                                                                                   leakiness, n, C, acc, bn_replicas());
  SCN_LAUNCHED();
  const long long total = (long long)n * C;
  const size_t sm = (size_t)6 * C * sizeof(float);
  if (vec)
    SCN_LAUNCH((k_bn_bwd_apply<true>), apply_grid(total / 4), BN_T, sm, s, in, out, d_out, d_in, save_mean, acc, save_invstd,
                                                              weight, d_weight, d_bias, leakiness, n, C, other, other_used,
                                                              residual, bias, bn_replicas());
  else
    SCN_LAUNCH((k_bn_bwd_apply<false>), apply_grid(total), BN_T, sm, s, in, out, d_out, d_in, save_mean, acc, save_invstd,
                                                           weight, d_weight, d_bias, leakiness, n, C, other, other_used,
                                                           residual, bias, bn_replicas());
  SCN_LAUNCHED();
  prof_end(PROF_BN, s, 5.0 * 4.0 * (double)n * C, 0);  // 5 n C s
  return 0;
}

}  // extern "C"
