// conv.cu - sparse convolution compute: output-stationary gather-GEMM (forward and dX), the
// per-offset weight-gradient contraction, bias helpers, NetworkInNetwork.
//
// Reference math (SparseConvNet/sparseconvnet/SCN/CPU/Convolution.cpp:46-185,
// CPU/Deconvolution.cpp:8-77): Y = bias; for k: Y[out_k] += X[in_k] @ W[k];
// dX = 0; for k: dX[in_k] += dY[out_k] @ W[k]^T; dW[k] = X[in_k]^T @ dY[out_k].
// The reference runs one gather / GEMM / scatter-add per offset; here every stationary row is
// produced once: a tile of 128 rows loops over the offsets active in the tile, gathers the
// partner rows and accumulates in registers (this file, exact fp32) or TMEM (conv_tc.cu).
#include "metadata.cuh"
#include "conv.cuh"
#include "../../include/scn_b200.h"

namespace scn {

// ---------------------------------------------------------------------------------------
// FFMA output-stationary gather-GEMM: tile 128 rows x BN cols, BK=16, 256 threads, 8xTN/thread
// ---------------------------------------------------------------------------------------
constexpr int BK = 16;
constexpr int LDA = TILE_M + 4;

template <int BN, bool VA, bool VB>
__global__ void __launch_bounds__(256)
k_osgemm_ffma(const float *__restrict__ X, const float *__restrict__ W,
              const float *__restrict__ bias, float *__restrict__ Y, int Cin, int Cout,
              long long n_rows, TileView tb) {
  pdl_sync();
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][LDA];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ int32_t sIdx[MAX_K][TILE_M];
  __shared__ int32_t sPerm[TILE_M];
  __shared__ int8_t sK[MAX_K];

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;

  int nE;
  if (tb.identity) {
    nE = 1;
    if (tid < TILE_M) {
      long long r = (long long)tile * TILE_M + tid;
      int v = r < n_rows ? (int)r : -1;
      sPerm[tid] = v;
      sIdx[0][tid] = v;
    }
    if (tid == 0) sK[0] = 0;
  } else {
    const uint32_t mask = tb.tile_mask[tile];
    const int e0 = tb.tile_off[tile];
    nE = __popc(mask);
    if (tid < TILE_M) sPerm[tid] = tb.perm[(long long)tile * TILE_M + tid];
    for (int i = tid; i < nE * TILE_M; i += 256)
      sIdx[i / TILE_M][i % TILE_M] = tb.entries[(long long)e0 * TILE_M + i];
    if (tid < 32) {
      // j-th set bit of the mask -> kernel offset of entry j
      if (mask & (1u << tid)) sK[__popc(mask & ((1u << tid) - 1u))] = (int8_t)tid;
    }
  }
  __syncthreads();

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int kchunks = (Cin + BK - 1) / BK;
  const int steps = nE * kchunks;

  float4 ra[2];
  float ras[8];
  float4 rb4;
  float rbs[TN];

  auto load_step = [&](int st) {
    const int e = st / kchunks, kc = (st - e * kchunks) * BK;
    const int k = sK[e];
    if (VA) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int f = tid + j * 256;
        const int row = f >> 2, c4 = f & 3;
        const int idx = sIdx[e][row];
        const int col = kc + c4 * 4;
        ra[j] = (idx >= 0 && col < Cin)
                    ? __ldg(reinterpret_cast<const float4 *>(X + (long long)idx * Cin + col))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int f = tid + j * 256;
        const int row = f >> 4, kk = f & 15;
        const int idx = sIdx[e][row];
        ras[j] = (idx >= 0 && kc + kk < Cin) ? __ldg(X + (long long)idx * Cin + kc + kk) : 0.f;
      }
    }
    const float *Wk = W + (long long)(tb.k_flip >= 0 ? tb.k_flip - (tb.k_base + k) : tb.k_base + k) * Cin * Cout;
    if (VB) {
      if (tid < BK * BN / 4) {
        const int kk = tid / (BN / 4), n4 = tid % (BN / 4);
        const int col = n0 + n4 * 4;
        rb4 = (kc + kk < Cin && col < Cout)
                  ? __ldg(reinterpret_cast<const float4 *>(Wk + (long long)(kc + kk) * Cout + col))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int f = tid + j * 256;
        const int kk = f / BN, nn = f % BN;
        rbs[j] = (kc + kk < Cin && n0 + nn < Cout) ? __ldg(Wk + (long long)(kc + kk) * Cout + n0 + nn)
                                                    : 0.f;
      }
    }
  };
  auto store_step = [&]() {
    if (VA) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int f = tid + j * 256;
        const int row = f >> 2, c4 = f & 3;
        As[c4 * 4 + 0][row] = ra[j].x;
        As[c4 * 4 + 1][row] = ra[j].y;
        As[c4 * 4 + 2][row] = ra[j].z;
        As[c4 * 4 + 3][row] = ra[j].w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int f = tid + j * 256;
        As[f & 15][f >> 4] = ras[j];
      }
    }
    if (VB) {
      if (tid < BK * BN / 4) {
        const int kk = tid / (BN / 4), n4 = tid % (BN / 4);
        *reinterpret_cast<float4 *>(&Bs[kk][n4 * 4]) = rb4;
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int f = tid + j * 256;
        Bs[f / BN][f % BN] = rbs[j];
      }
    }
  };

  if (steps > 0) {
    load_step(0);
    store_step();
  }
  __syncthreads();
  for (int st = 0; st < steps; ++st) {
    if (st + 1 < steps) load_step(st + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      if (TN == 4) {
        const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
        b[0] = b4.x; b[1] = b4.y; b[2] = b4.z; b[3] = b4.w;
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (st + 1 < steps) store_step();
    __syncthreads();
  }

  // epilogue: every stationary row is written exactly once (no atomics, no zero-fill pass)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int orow = sPerm[ty * 8 + i];
    if (orow < 0) continue;
    float *yp = Y + (long long)orow * Cout + n0 + tx * TN;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + tx * TN + j;
      if (col < Cout) yp[j] = acc[i][j] + (bias ? bias[col] : 0.f);
    }
  }
}

template <int BN>
static int launch_osgemm_ffma(const float *X, const float *W, const float *bias, float *Y, int Cin,
                              int Cout, long long n_rows, const TileView &tv, cudaStream_t s) {
  dim3 grid(tv.n_tiles, cdiv(Cout, BN));
  const bool va = (Cin % 4 == 0) && ((uintptr_t)X % 16 == 0);
  const bool vb = (Cout % 4 == 0) && ((uintptr_t)W % 16 == 0);
  if (va && vb) SCN_LAUNCH((k_osgemm_ffma<BN, true, true>), grid, 256, 0, s, X, W, bias, Y, Cin, Cout, n_rows, tv);
  else if (!va && vb) SCN_LAUNCH((k_osgemm_ffma<BN, false, true>), grid, 256, 0, s, X, W, bias, Y, Cin, Cout, n_rows, tv);
  else SCN_LAUNCH((k_osgemm_ffma<BN, false, false>), grid, 256, 0, s, X, W, bias, Y, Cin, Cout, n_rows, tv);
  SCN_LAUNCHED();
  return 0;
}

TileView make_view(const TileBook &tb, int k_flip) {
  TileView v;
  v.identity = tb.identity ? 1 : 0;
  v.n_tiles = tb.n_tiles;
  v.k_flip = k_flip;
  v.k_base = tb.k_base;
  v.perm = tb.perm;
  v.tile_mask = tb.tile_mask;
  v.tile_off = tb.tile_off;
  v.order = tb.identity ? nullptr : reinterpret_cast<const int4 *>(tb.order);
  v.entries = tb.entries;
  return v;
}

// dst[r][0..cp) = src[r][0..c), zero beyond: narrow feature rows (the 9-channel stem input) padded to one
// 32-channel slice so that they take the tensor-core kernels
__global__ void k_pad_cols(const float *__restrict__ src, float *__restrict__ dst, long long rows, int c, int cp) {
  pdl_sync();
  const long long total = rows * cp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cp;
    const int j = (int)(i - r * cp);
    dst[i] = j < c ? src[r * c + j] : 0.f;
  }
}
// Wp[k][0..cp)[co] = W[k][0..c)[co], zero rows beyond
__global__ void k_pad_w_rows(const float *__restrict__ W, float *__restrict__ Wp, int K, int c, int cp, int cout) {
  pdl_sync();
  const int total = K * cp * cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / (cp * cout), r = i - k * cp * cout, ci = r / cout, co = r - ci * cout;
    Wp[i] = ci < c ? W[((long long)k * c + ci) * cout + co] : 0.f;
  }
}
constexpr int PAD_C = 32;
static bool pad_ok(int Cin, int Cout, int precision) {
  return precision != SCN_PRECISION_FP32 && Cin < PAD_C && Cout >= 16 && Cout % 16 == 0 && Cout <= 256;
}
static int pad_rows(const float *X, long long rows, int Cin, float **out, cudaStream_t s) {
  SCN_TRY(workspace_t(out, WS_PAD_X, (size_t)rows * PAD_C, s));
  long long blocks = (rows * PAD_C + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  SCN_LAUNCH(k_pad_cols, (int)(blocks < 1 ? 1 : blocks), 256, 0, s, X, *out, rows, Cin, PAD_C);
  SCN_LAUNCHED();
  return 0;
}

__global__ void k_accumulate(float *__restrict__ y, const float *__restrict__ t, long long n) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] += t[i];
}

// one tile book: Y[stationary rows] = bias + sum over the book's offsets of X[partner_k] @ W[k]; K = offsets of
// the whole filter (W: [K,Cin,Cout] row-major)
static int osgemm_book(const float *X, const float *W, const float *bias, float *Y, int Cin, int Cout,
                       const TileBook &tb, int K, double bytes, double flops, int precision, int transpose_w,
                       cudaStream_t s, int k_flip, const int64_t *weight_tag) {
  const TileView tv = make_view(tb, k_flip);
  int r = 1;
  if (!transpose_w && !tb.identity && pad_ok(Cin, Cout, precision)) {
    float *xp = nullptr, *wp = nullptr;
    SCN_TRY(pad_rows(X, tb.n_partner, Cin, &xp, s));
    SCN_TRY(workspace_t(&wp, WS_PAD_W, (size_t)K * PAD_C * Cout, s));
    SCN_LAUNCH(k_pad_w_rows, cdiv((long long)K * PAD_C * Cout, 256), 256, 0, s, W, wp, K, Cin, PAD_C, Cout);
    SCN_LAUNCHED();
    r = osgemm_tc(xp, wp, bias, Y, PAD_C, Cout, tb.n_rows, tv, K, tb.K, precision, 0, s, bytes, flops, nullptr);
  } else if (precision != SCN_PRECISION_FP32)
    r = osgemm_tc(X, W, bias, Y, Cin, Cout, tb.n_rows, tv, K, tb.K, precision, transpose_w, s, bytes, flops, weight_tag);
  if (r > 0) {  // 0 = done, negative = -(error); positive = shape outside the tensor path (e.g. Cin = 9)
    float *wt = nullptr;
    r = 0;
    if (transpose_w) {
      r = workspace_t(&wt, WS_WT, (size_t)K * Cin * Cout, s);
      if (!r) r = transpose_weights(W, wt, K, Cout, Cin, s);   // W is [K][Cout=N][Cin=Kd] -> [K][Kd][N]
    }
    const float *w = transpose_w ? wt : W;
    prof_begin(PROF_GEMM, s);
    if (!r)
      r = Cout <= 32 ? launch_osgemm_ffma<32>(X, w, bias, Y, Cin, Cout, tb.n_rows, tv, s)
                     : launch_osgemm_ffma<64>(X, w, bias, Y, Cin, Cout, tb.n_rows, tv, s);
    prof_end(PROF_GEMM, s, bytes, flops);
  } else {
    r = -r;
  }
  return r;
}

// Y[stationary rows] = bias + sum_k X[partner_k] @ W[k]   (W: [K,Cin,Cout] row-major).  Filters of more than
// MAX_K offsets come as a chain of tile books (metadata.cuh): the first book writes Y, every further one its
// own partial into a scratch buffer that is then added (rare: FPN_Net's [1,1,64] z-collapse, 4^3 / 5^3 filters).
int osgemm(const float *X, const float *W, const float *bias, float *Y, int Cin, int Cout,
           const TileBook &tb, int precision, int transpose_w, cudaStream_t s, int k_flip,
           const int64_t *weight_tag) {
  if (tb.n_tiles == 0) return 0;
  int K = 0;
  for (const TileBook *b = &tb; b; b = b->next) K = b->k_base + b->K;
  // algorithmic bytes: every feature row once, the weights once, 8 B per pair.  The timed
  // region is the gather-GEMM kernel itself (weight packing / split-K reduce are outside it).
  const double bytes = 4.0 * ((double)tb.n_partner * Cin + (double)tb.n_rows * Cout) +
                       4.0 * K * Cin * Cout + (tb.identity ? 0.0 : 8.0 * tb.n_pairs);
  const double flops = 2.0 * tb.n_pairs * Cin * Cout;
  SCN_TRY(osgemm_book(X, W, bias, Y, Cin, Cout, tb, K, bytes, flops, precision, transpose_w, s, k_flip, weight_tag));
  for (const TileBook *b = tb.next; b; b = b->next) {
    float *part = nullptr;
    const long long n = (long long)tb.n_rows * Cout;
    SCN_TRY(workspace_t(&part, WS_CHAIN, (size_t)n, s));
    SCN_TRY(osgemm_book(X, W, nullptr, part, Cin, Cout, *b, K, 0.0, 0.0, precision, transpose_w, s, k_flip, weight_tag));
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    SCN_LAUNCH(k_accumulate, (int)blocks, 256, 0, s, Y, part, n);
    SCN_LAUNCHED();
  }
  return 0;
}

// Wt[k][co][ci] = W[k][ci][co]
__global__ void k_transpose_w(const float *__restrict__ W, float *__restrict__ Wt, int K, int Cin,
                              int Cout) {
  pdl_sync();
  __shared__ float t[32][33];
  const int k = blockIdx.z;
  const float *w = W + (long long)k * Cin * Cout;
  float *wt = Wt + (long long)k * Cin * Cout;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ci = ci0 + i, co = co0 + threadIdx.x;
    t[i][threadIdx.x] = (ci < Cin && co < Cout) ? w[(long long)ci * Cout + co] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int co = co0 + i, ci = ci0 + threadIdx.x;
    if (co < Cout && ci < Cin) wt[(long long)co * Cin + ci] = t[threadIdx.x][i];
  }
}

int transpose_weights(const float *W, float *Wt, int K, int Cin, int Cout, cudaStream_t s) {
  dim3 grid(cdiv(Cout, 32), cdiv(Cin, 32), K), block(32, 8);
  SCN_LAUNCH(k_transpose_w, grid, block, 0, s, W, Wt, K, Cin, Cout);
  SCN_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------
// weight gradient: dW[k] = sum over pairs of offset k of  x_row^T (outer) dy_row
// phase 1: one CTA per (work item, Cin block, Cout block) -> partial; phase 2: fixed-order sum
// ---------------------------------------------------------------------------------------
constexpr int DW_P = 16;  // pairs per smem step

template <int TI, int TJ>
__global__ void __launch_bounds__(256)
k_dw_partial(const float *__restrict__ X, const float *__restrict__ dY,
             const int32_t *__restrict__ pairs, const DwWork *__restrict__ work,
             float *__restrict__ partial, int Cin, int Cout, int xcol, int ycol,
             long long ident_n, int ident_chunk) {
  pdl_sync();
  constexpr int CI = 16 * TI, CO = 16 * TJ;
  __shared__ __align__(16) float Xs[DW_P][CI];
  __shared__ __align__(16) float Ys[DW_P][CO];
  __shared__ int32_t sx[DW_P], sy[DW_P];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ci0 = blockIdx.y * CI, co0 = blockIdx.z * CO;
  long long start;
  int len, out_slot = blockIdx.x;
  if (work) {
    const DwWork w = work[blockIdx.x];
    start = w.start;
    len = w.len;
    out_slot = w.slot;
  } else {
    start = (long long)blockIdx.x * ident_chunk;
    len = (int)min((long long)ident_chunk, ident_n - start);
  }
  float acc[TI][TJ];
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TJ; ++j) acc[i][j] = 0.f;

  for (int p0 = 0; p0 < len; p0 += DW_P) {
    if (tid < DW_P) {
      int xi = -1, yi = -1;
      if (p0 + tid < len) {
        if (pairs) {
          const int2 pr = reinterpret_cast<const int2 *>(pairs)[start + p0 + tid];
          xi = xcol ? pr.y : pr.x;
          yi = ycol ? pr.y : pr.x;
        } else {
          xi = yi = (int)(start + p0 + tid);
        }
      }
      sx[tid] = xi;
      sy[tid] = yi;
    }
    __syncthreads();
    for (int f = tid; f < DW_P * CI; f += 256) {
      const int p = f / CI, c = f % CI;
      const int r = sx[p];
      Xs[p][c] = (r >= 0 && ci0 + c < Cin) ? __ldg(X + (long long)r * Cin + ci0 + c) : 0.f;
    }
    for (int f = tid; f < DW_P * CO; f += 256) {
      const int p = f / CO, c = f % CO;
      const int r = sy[p];
      Ys[p][c] = (r >= 0 && co0 + c < Cout) ? __ldg(dY + (long long)r * Cout + co0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < DW_P; ++p) {
      float a[TI], b[TJ];
#pragma unroll
      for (int i = 0; i < TI; ++i) a[i] = Xs[p][ty * TI + i];
#pragma unroll
      for (int j = 0; j < TJ; ++j) b[j] = Ys[p][tx * TJ + j];
#pragma unroll
      for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float *out = partial + (long long)out_slot * Cin * Cout;
#pragma unroll
  for (int i = 0; i < TI; ++i) {
    const int ci = ci0 + ty * TI + i;
    if (ci >= Cin) continue;
#pragma unroll
    for (int j = 0; j < TJ; ++j) {
      const int co = co0 + tx * TJ + j;
      if (co < Cout) out[(long long)ci * Cout + co] = acc[i][j];
    }
  }
}

struct KFirst { int v[MAX_KT + 1]; };

// cc = Cin*Cout elements of dW[k]; ccp = elements of one partial (Cin padded up for narrow inputs: the
// first Cin rows of a partial are the real ones)
__global__ void k_dw_reduce(const float *__restrict__ partial, float *__restrict__ dW, KFirst first,
                            int cc /* Cin*Cout */, int ccp) {
  pdl_sync();
  const int k = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cc) return;
  float s = 0.f;
  for (int w = first.v[k]; w < first.v[k + 1]; ++w) s += partial[(long long)w * ccp + e];
  dW[(long long)k * cc + e] = s;
}

template <int TI, int TJ>
static int launch_dw_partial(const float *X, const float *dY, const int32_t *pairs,
                             const DwWork *work, float *partial, int Cin, int Cout, int xcol,
                             int ycol, int n_work, long long ident_n, int ident_chunk,
                             cudaStream_t s) {
  dim3 grid(n_work, cdiv(Cin, 16 * TI), cdiv(Cout, 16 * TJ));
  SCN_LAUNCH((k_dw_partial<TI, TJ>), grid, 256, 0, s, X, dY, pairs, work, partial, Cin, Cout, xcol, ycol,
                                            ident_n, ident_chunk);
  SCN_LAUNCHED();
  return 0;
}

// dW[k] (fully written for all k) from the rulebook's pair lists.  xcol/ycol select which pair
// column indexes X and dY (convolution: 0,1; deconvolution: 1,0).
int weight_grad(const float *X, const float *dY, float *dW, int Cin, int Cout, RuleBook *rb,
                int xcol, int ycol, int precision, cudaStream_t s) {
  const int K = rb->K;
  const int cc = Cin * Cout;
  KFirst first;
  int n_work;
  const int32_t *pairs = nullptr;
  const DwWork *work = nullptr;
  long long ident_n = 0;
  int ident_chunk = 0;
  if (rb->identity) {
    ident_n = rb->total_pairs;
    long long chunk = (ident_n + 2LL * num_sms() - 1) / (2LL * num_sms());
    chunk = (chunk + 63) / 64 * 64;
    if (chunk < 512) chunk = 512;
    if (chunk > 16384) chunk = 16384;
    ident_chunk = (int)chunk;
    n_work = cdiv(ident_n, chunk);
    first.v[0] = 0;
    first.v[1] = n_work;
  } else {
    SCN_TRY(ensure_dw_work(rb, s));
    n_work = rb->n_dw_work;
    pairs = rb->pairs;
    work = rb->dw_work;
    int w = 0;
    const long long chunk = rb->dw_chunk > 0 ? rb->dw_chunk : 1;
    for (int k = 0; k < K; ++k) {
      first.v[k] = w;
      w += (int)((rb->counts[k] + chunk - 1) / chunk);
    }
    first.v[K] = w;
  }
  prof_begin(PROF_DW, s);
  const double dw_bytes = 4.0 * ((double)(xcol ? rb->n_out : rb->n_in) * Cin + (double)(ycol ? rb->n_out : rb->n_in) * Cout) +
                          4.0 * K * cc + (rb->identity ? 0.0 : 8.0 * rb->total_pairs);
  const double dw_flops = 2.0 * rb->total_pairs * cc;
  float *partial = nullptr;
  int ccp = cc;
  if (n_work > 0) {
    const int cmax = Cin < Cout ? Cout : Cin;
    int r = 1;
    // narrow inputs (the 9-channel stem): X padded to 32 channels, tensor-core partials of 32 x Cout
    if (!rb->identity && pad_ok(Cin, Cout, precision) && Cout % 32 == 0) {
      float *xp = nullptr;
      const long long xrows = xcol ? rb->n_out : rb->n_in;
      SCN_TRY(pad_rows(X, xrows, Cin, &xp, s));
      ccp = PAD_C * Cout;
      SCN_TRY(workspace_t(&partial, WS_DW_PARTIAL, (size_t)n_work * ccp, s));
      r = dw_partial_tc(xp, dY, pairs, work, partial, PAD_C, Cout, xcol, ycol, n_work, ident_n, ident_chunk, precision, s);
      if (r > 0) ccp = cc;
    }
    if (r > 0) SCN_TRY(workspace_t(&partial, WS_DW_PARTIAL, (size_t)n_work * cc, s));
    if (r > 0 && precision != SCN_PRECISION_FP32)
      r = dw_partial_tc(X, dY, pairs, work, partial, Cin, Cout, xcol, ycol, n_work, ident_n, ident_chunk, precision, s);
    if (r < 0) r = 1000;   // error already recorded
    else if (r == 0) r = 0;
    else if (cmax <= 32) r = launch_dw_partial<2, 2>(X, dY, pairs, work, partial, Cin, Cout, xcol, ycol, n_work, ident_n, ident_chunk, s);
    else if (cmax <= 64) r = launch_dw_partial<4, 4>(X, dY, pairs, work, partial, Cin, Cout, xcol, ycol, n_work, ident_n, ident_chunk, s);
    else r = launch_dw_partial<8, 8>(X, dY, pairs, work, partial, Cin, Cout, xcol, ycol, n_work, ident_n, ident_chunk, s);
    if (r) return 1;
  }
  dim3 grid(cdiv(cc, 256), K);
  SCN_LAUNCH(k_dw_reduce, grid, 256, 0, s, partial, dW, first, cc, ccp);
  SCN_LAUNCHED();
  prof_end(PROF_DW, s, dw_bytes, dw_flops);
  return 0;
}

// d_bias[c] = sum_rows d_out[row][c]  (CPU/Convolution.cpp:99-100)
__global__ void k_colsum(const float *__restrict__ A, float *__restrict__ out, long long n, int C) {
  pdl_sync();
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C)
    for (long long r = threadIdx.y; r < n; r += 8) s += A[r * C + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[c] = t;
  }
}

int bias_grad(const float *d_out, float *d_bias, long long n, int C, cudaStream_t s) {
  if (!d_bias) return 0;
  SCN_LAUNCH(k_colsum, cdiv(C, 32), dim3(32, 8), 0, s, d_out, d_bias, n, C);
  SCN_LAUNCHED();
  return 0;
}

extern thread_local bool g_defer_dw_join, g_dw_join_pending;
extern bool g_dw_companion;

static double macs_of(const RuleBook *rb, int64_t cin, int64_t cout) {
  return (double)rb->total_pairs * (double)cin * (double)cout;
}

// shared backward: dX through `tb_dx` with transposed weights, dW from the pair lists
static int conv_backward_common(RuleBook *rb, bool dx_stationary_out, const float *in, float *d_in,
                                const float *d_out, const float *weight, float *d_weight,
                                float *d_bias, int Cin, int Cout, int xcol, int ycol,
                                long long n_dout_rows, int precision, cudaStream_t s,
                                const int64_t *weight_tag = nullptr) {
  // submanifold rulebooks with odd filters are their own mirror image: the out-row that in-row i
  // feeds at offset k is the site at i - delta_k = t_out[K-1-k][i], so dX runs on the forward lists
  // with the weight index flipped and no second tile book is ever built
  const bool mirror = rb->kind == 0 && !rb->identity && (rb->filter[0] & rb->filter[1] & rb->filter[2] & 1);
  if (!mirror) SCN_TRY(ensure_tilebook(rb, dx_stationary_out, s));
  TileBook &tb = (dx_stationary_out || mirror) ? rb->tb_out : rb->tb_in;
  // dX and dW are independent: the weight gradient runs on the companion stream, so the many small
  // (latency-bound, few-CTA) launches of the coarse scales overlap instead of queueing
  SideStream *ss = nullptr;
  const bool fork = d_in && d_weight && g_dw_companion;
  if (fork) {
    SCN_TRY(side_stream(s, &ss));
    SCN_TRY(side_fork(s, ss));
  }
  // the caller's stream must never run ahead of a forked companion stream, also when a launch below fails: the guard
  // joins on every exit path unless the join was done (or handed to the sweep's marks) explicitly
  struct JoinGuard {
    cudaStream_t s;
    SideStream *ss;
    bool armed;
    ~JoinGuard() { if (armed) side_join(s, ss); }
  } join_guard{s, ss, fork};
  if (d_in)
    SCN_TRY(osgemm(d_out, weight, nullptr, d_in, Cout, Cin, tb, precision, /*transpose_w=*/1, s, mirror ? rb->K - 1 : -1,
                   weight_tag));
  if (d_weight) SCN_TRY(weight_grad(in, d_out, d_weight, Cin, Cout, rb, xcol, ycol, precision, fork ? ss->stream : s));
  // (layer-graph reverse sweep: the join is deferred to the sweep's progress marks / its end, so the weight
  // gradients of the latency-bound coarse scales run under the following layers instead of holding them up)
  if (fork && !g_defer_dw_join) {
    join_guard.armed = false;
    SCN_TRY(side_join(s, ss));
  }
  if (fork && g_defer_dw_join) {
    join_guard.armed = false;
    g_dw_join_pending = true;
  }
  SCN_TRY(bias_grad(d_out, d_bias, n_dout_rows, Cout, s));
  return 0;
}

bool g_dw_companion = true;              // false: weight gradients on the caller's stream (scn_set_graph_overlap(0))
thread_local bool g_defer_dw_join = false;
thread_local bool g_dw_join_pending = false;

int dw_join_pending(cudaStream_t s) {
  if (!g_dw_join_pending) return 0;
  SideStream *ss = nullptr;
  SCN_TRY(side_stream(s, &ss));
  SCN_TRY(side_join(s, ss));
  g_dw_join_pending = false;
  return 0;
}

}  // namespace scn

using namespace scn;

extern "C" {

int scn_submanifold_conv_forward(scn_metadata_t *m, const int64_t *ss, const int64_t *filter,
                                 const float *in, float *out, const float *weight,
                                 const float *bias, int64_t cin, int64_t cout, int precision,
                                 void *stream, double *macs, const int64_t *weight_tag) {
  SCN_CHECK(m && ss && filter && weight, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  RuleBook *rb = nullptr;
  SCN_TRY(get_submanifold_rulebook(m, ss, filter, s, &rb));
  if (macs) *macs = macs_of(rb, cin, cout);
  if (rb->n_out == 0) return 0;
  SCN_CHECK(in && out, "null feature pointer");
  return osgemm(in, weight, bias, out, (int)cin, (int)cout, rb->tb_out, precision, 0, s, -1, weight_tag);
}

int scn_submanifold_conv_backward(scn_metadata_t *m, const int64_t *ss, const int64_t *filter,
                                  const float *in, float *d_in, const float *d_out,
                                  const float *weight, float *d_weight, float *d_bias, int64_t cin,
                                  int64_t cout, int precision, void *stream, const int64_t *weight_tag) {
  SCN_CHECK(m && ss && filter && weight, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  RuleBook *rb = nullptr;
  SCN_TRY(get_submanifold_rulebook(m, ss, filter, s, &rb));
  if (rb->n_out == 0) {
    if (d_weight) SCN_CUDA(cudaMemsetAsync(d_weight, 0, (size_t)rb->K * cin * cout * 4, s));
    return 0;
  }
  return conv_backward_common(rb, /*dx over in rows*/ false, in, d_in, d_out, weight, d_weight,
                              d_bias, (int)cin, (int)cout, 0, 1, rb->n_out, precision, s, weight_tag);
}

int scn_conv_forward(scn_metadata_t *m, const int64_t *in_ss, const int64_t *out_ss,
                     const int64_t *filter, const int64_t *stride, const float *in, float *out,
                     const float *weight, const float *bias, int64_t cin, int64_t cout,
                     int precision, void *stream, double *macs, const int64_t *weight_tag) {
  SCN_CHECK(m && in_ss && out_ss && filter && stride && weight, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  RuleBook *rb = nullptr;
  SCN_TRY(get_conv_rulebook(m, in_ss, out_ss, filter, stride, s, &rb));
  if (macs) *macs = macs_of(rb, cin, cout);
  if (rb->n_out == 0) return 0;
  return osgemm(in, weight, bias, out, (int)cin, (int)cout, rb->tb_out, precision, 0, s, -1, weight_tag);
}

int scn_conv_backward(scn_metadata_t *m, const int64_t *in_ss, const int64_t *out_ss,
                      const int64_t *filter, const int64_t *stride, const float *in, float *d_in,
                      const float *d_out, const float *weight, float *d_weight, float *d_bias,
                      int64_t cin, int64_t cout, int precision, void *stream, const int64_t *weight_tag) {
  SCN_CHECK(m && in_ss && out_ss && filter && stride && weight, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  RuleBook *rb = nullptr;
  SCN_TRY(get_conv_rulebook(m, in_ss, out_ss, filter, stride, s, &rb));
  if (rb->n_in == 0) {
    if (d_weight) SCN_CUDA(cudaMemsetAsync(d_weight, 0, (size_t)rb->K * cin * cout * 4, s));
    return 0;
  }
  return conv_backward_common(rb, false, in, d_in, d_out, weight, d_weight, d_bias, (int)cin,
                              (int)cout, 0, 1, rb->n_out, precision, s, weight_tag);
}

// Deconvolution: rulebook of the matching down-convolution with roles swapped
// (CPU/Deconvolution.cpp:15-16,34-37): in = coarse scale, out = fine scale.
int scn_deconv_forward(scn_metadata_t *m, const int64_t *in_ss, const int64_t *out_ss,
                       const int64_t *filter, const int64_t *stride, const float *in, float *out,
                       const float *weight, const float *bias, int64_t cin, int64_t cout,
                       int precision, void *stream, double *macs, const int64_t *weight_tag) {
  SCN_CHECK(m && in_ss && out_ss && filter && stride && weight, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  RuleBook *rb = nullptr;
  SCN_TRY(get_conv_rulebook(m, out_ss, in_ss, filter, stride, s, &rb));
  if (macs) *macs = macs_of(rb, cin, cout);
  if (rb->n_in == 0) return 0;
  SCN_TRY(ensure_tilebook(rb, false, s));
  return osgemm(in, weight, bias, out, (int)cin, (int)cout, rb->tb_in, precision, 0, s, -1, weight_tag);
}

int scn_deconv_backward(scn_metadata_t *m, const int64_t *in_ss, const int64_t *out_ss,
                        const int64_t *filter, const int64_t *stride, const float *in, float *d_in,
                        const float *d_out, const float *weight, float *d_weight, float *d_bias,
                        int64_t cin, int64_t cout, int precision, void *stream, const int64_t *weight_tag) {
  SCN_CHECK(m && in_ss && out_ss && filter && stride && weight, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  RuleBook *rb = nullptr;
  SCN_TRY(get_conv_rulebook(m, out_ss, in_ss, filter, stride, s, &rb));
  if (rb->n_out == 0) {
    if (d_weight) SCN_CUDA(cudaMemsetAsync(d_weight, 0, (size_t)rb->K * cin * cout * 4, s));
    return 0;
  }
  // d_in lives on the coarse scale = the rulebook's "out" side
  return conv_backward_common(rb, true, in, d_in, d_out, weight, d_weight, d_bias, (int)cin,
                              (int)cout, 1, 0, rb->n_in, precision, s, weight_tag);
}

int scn_nin_forward(const float *in, float *out, const float *weight, const float *bias,
                    int64_t n_rows, int64_t cin, int64_t cout, int precision, void *stream,
                    double *macs) {
  if (macs) *macs = (double)n_rows * (double)cin * (double)cout;
  if (n_rows == 0) return 0;
  TileBook tb;
  tb.identity = true; tb.built = true; tb.K = 1;
  tb.n_rows = tb.n_partner = n_rows;
  tb.n_tiles = cdiv(n_rows, TILE_M);
  return osgemm(in, weight, bias, out, (int)cin, (int)cout, tb, precision, 0, (cudaStream_t)stream);
}

int scn_nin_backward(const float *in, float *d_in, const float *d_out, const float *weight,
                     float *d_weight, float *d_bias, int64_t n_rows, int64_t cin, int64_t cout,
                     int precision, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_rows == 0) {
    if (d_weight) SCN_CUDA(cudaMemsetAsync(d_weight, 0, (size_t)cin * cout * 4, s));
    return 0;
  }
  RuleBook rb;
  rb.kind = 0; rb.K = 1; rb.identity = true;
  rb.n_in = rb.n_out = n_rows;
  memset(rb.counts, 0, sizeof(rb.counts));
  memset(rb.pair_off, 0, sizeof(rb.pair_off));
  rb.counts[0] = n_rows; rb.pair_off[1] = n_rows; rb.total_pairs = n_rows;
  return conv_backward_common(&rb, false, in, d_in, d_out, weight, d_weight, d_bias, (int)cin,
                              (int)cout, 0, 1, n_rows, precision, s);
}

}  // extern "C"
