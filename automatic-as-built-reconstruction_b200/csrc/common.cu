// common.cu - error state, stream-ordered allocation, device-wide scan and radix sort.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <mutex>
#include <vector>
#include "../../include/scn_b200.h"

namespace scn {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- profiling regions ----------------------------------------------------------------------
bool g_prof_on = false;
// programmatic dependent launch (common.cuh launch_k): 1 = every kernel (default), 2 = all but the tensor-core GEMMs, 0 = off
int g_pdl = getenv("SCN_B200_PDL") ? atoi(getenv("SCN_B200_PDL")) : 1;
struct ProfSeg { int cls; cudaEvent_t a, b; double bytes, flops; };
static std::vector<ProfSeg> g_prof_segs;
static std::vector<cudaEvent_t> g_prof_open[PROF_N];
static std::mutex g_prof_mu;

void prof_begin_(int cls, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, s);
  g_prof_open[cls].push_back(e);
}
void prof_end_(int cls, cudaStream_t s, double bytes, double flops) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_open[cls].empty()) return;
  ProfSeg sg;
  sg.cls = cls;
  sg.a = g_prof_open[cls].back();
  g_prof_open[cls].pop_back();
  cudaEventCreate(&sg.b);
  cudaEventRecord(sg.b, s);
  sg.bytes = bytes;
  sg.flops = flops;
  g_prof_segs.push_back(sg);
}

static int g_num_sms = 0;

static void init_pool() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ULL;  // never trim: the per-step Metadata rebuild reuses it
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    // Two streams allocate from this pool (feature kernels / the prefetcher's metadata builds).  Never
    // let an allocation on one stream wait for the other stream's queue to reach a free: take a
    // block whose free has already completed, or fresh memory.
    int off = 0;
    cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
    // Pre-size the pool (it never trims): without this the first ~25 steps each extend it through the
    // driver (hundreds of microseconds per extension) until it covers the window of steps in flight.
    // SCN_B200_POOL_MB overrides the 6 GB default (0 = grow on demand).
    const char *env = getenv("SCN_B200_POOL_MB");
    const size_t mb = env ? (size_t)atoll(env) : 6144;
    if (mb > 0) {
      void *p = nullptr;
      if (cudaMallocAsync(&p, mb << 20, 0) == cudaSuccess) {
        cudaFreeAsync(p, 0);
        cudaStreamSynchronize(0);
      } else {
        cudaGetLastError();   // not enough free memory: fall back to growing on demand
      }
    }
  }
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
}

// once per DEVICE (a process may drive several): the pool attributes and the pre-sizing belong to the current device's pool
static std::once_flag g_pool_once_dev[64];
static void ensure_pool() {
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(g_pool_once_dev[dev & 63], init_pool);
}

int num_sms() {
  ensure_pool();
  return g_num_sms > 0 ? g_num_sms : 148;
}

int dev_alloc(void **p, size_t bytes, cudaStream_t s) {
  ensure_pool();
  *p = nullptr;
  SCN_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, s));
  return 0;
}

void dev_free(void *p, cudaStream_t s) {
  if (p) cudaFreeAsync(p, s);
}

struct WsBuf { void *p = nullptr; size_t cap = 0; };
static std::mutex g_ws_mu;
static std::vector<std::pair<StreamKey, WsBuf *>> g_ws;   // a handful of streams: linear search

int workspace(void **p, int slot, size_t bytes, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  WsBuf *bufs = nullptr;
  for (auto &e : g_ws)
    if (e.first == stream_key(s)) { bufs = e.second; break; }
  if (!bufs) {
    bufs = new WsBuf[WS_SLOTS];
    g_ws.emplace_back(stream_key(s), bufs);
  }
  WsBuf &b = bufs[slot];
  if (b.cap < bytes) {
    if (b.p) cudaFreeAsync(b.p, s);
    b.p = nullptr;
    b.cap = 0;
    size_t cap = bytes + bytes / 4;
    cap = (cap + 255) & ~(size_t)255;
    SCN_TRY(dev_alloc(&b.p, cap, s));
    b.cap = cap;
  }
  *p = b.p;
  return 0;
}

static std::vector<std::pair<StreamKey, int *>> g_sched;

int sched_counters(int **p, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  for (auto &e : g_sched)
    if (e.first == stream_key(s)) { *p = e.second; return 0; }
  int *d = nullptr;
  SCN_CUDA(cudaMalloc((void **)&d, 64));
  SCN_CUDA(cudaMemset(d, 0, 64));          // synchronous: visible to every stream that follows
  g_sched.emplace_back(stream_key(s), d);
  *p = d;
  return 0;
}

static std::mutex g_side_mu;
static std::vector<std::pair<StreamKey, SideStream *>> g_side;

int side_stream(cudaStream_t s, SideStream **out) {
  std::lock_guard<std::mutex> lk(g_side_mu);
  for (auto &e : g_side)
    if (e.first == stream_key(s)) { *out = e.second; return 0; }
  SideStream *ss = new SideStream();
  SCN_CUDA(cudaStreamCreateWithFlags(&ss->stream, cudaStreamNonBlocking));
  SCN_CUDA(cudaEventCreateWithFlags(&ss->fork, cudaEventDisableTiming));
  SCN_CUDA(cudaEventCreateWithFlags(&ss->join, cudaEventDisableTiming));
  g_side.emplace_back(stream_key(s), ss);
  *out = ss;
  return 0;
}
int side_fork(cudaStream_t s, SideStream *ss) {
  SCN_CUDA(cudaEventRecord(ss->fork, s));
  SCN_CUDA(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
  return 0;
}
int side_join(cudaStream_t s, SideStream *ss) {
  SCN_CUDA(cudaEventRecord(ss->join, ss->stream));
  SCN_CUDA(cudaStreamWaitEvent(s, ss->join, 0));
  return 0;
}

int64_t *host_scratch(size_t n) {
  static thread_local int64_t *buf = nullptr;
  static thread_local size_t cap = 0;
  if (n > cap) {
    if (buf) cudaFreeHost(buf);
    cap = n < 4096 ? 4096 : n;
    if (cudaMallocHost((void **)&buf, cap * sizeof(int64_t)) != cudaSuccess) {
      buf = nullptr;
      cap = 0;
    }
  }
  return buf;
}

// ---------------------------------------------------------------------------------------
// exclusive scan in ONE pass: 512 threads x 8 items per tile, tiles chained by decoupled look-back
// (Merrill & Garland).  A tile takes its index from a per-stream ticket counter (so every predecessor has
// started), publishes its aggregate, then walks back over its predecessors' descriptors until it meets an
// inclusive prefix.  Descriptor = epoch << 34 | state << 32 | value in one 64-bit word: a word of another epoch
// reads as "not ready", so neither the descriptors nor the ticket counter are cleared between scans (the host
// keeps the per-stream epoch and the running ticket base).  FLAGS: the scanned value is (in[i] >= 0), which
// fuses the rulebook's "partner present" flag pass into the scan.
// ---------------------------------------------------------------------------------------
constexpr int SCAN_T = 512, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;

template <bool FLAGS>
__global__ void __launch_bounds__(SCAN_T)
k_scan_lookback(const int32_t *in, int32_t *out, long long n_in, long long n_out,
                volatile unsigned long long *desc, unsigned int *ticket, unsigned int ticket_base,
                unsigned long long epoch) {
  pdl_sync();
  __shared__ int32_t warp_tot[SCAN_T / 32];
  __shared__ unsigned int s_tile;
  __shared__ int32_t s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u) - ticket_base;
  __syncthreads();
  const unsigned int tile = s_tile;
  const long long base = (long long)tile * SCAN_B + (long long)threadIdx.x * SCAN_I;
  int32_t v[SCAN_I];
  int32_t tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_I; ++i) {
    long long g = base + i;
    int32_t x = (g < n_in) ? in[g] : (FLAGS ? -1 : 0);
    v[i] = FLAGS ? (x >= 0 ? 1 : 0) : x;
    tsum += v[i];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int32_t inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int32_t w = (lane < SCAN_T / 32) ? warp_tot[lane] : 0;
    int32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < SCAN_T / 32) warp_tot[lane] = wi - w;  // exclusive warp offsets
    const int32_t agg = __shfl_sync(0xffffffffu, wi, SCAN_T / 32 - 1);
    const unsigned long long tag = epoch << 34;
    int32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) desc[0] = tag | (2ull << 32) | (unsigned int)agg;
    } else {
      if (lane == 0) desc[tile] = tag | (1ull << 32) | (unsigned int)agg;
      // look back 32 predecessors at a time: lane l reads tile - 1 - l (of the current window)
      long long first = (long long)tile - 1;
      for (;;) {
        const long long t = first - lane;
        unsigned long long d = 0;
        bool ready = t < 0;                         // before tile 0: nothing to add, counts as "prefix"
        unsigned int st = 2;
        int32_t val = 0;
        if (t >= 0) {
          d = desc[t];
          ready = (d >> 34) == epoch && ((d >> 32) & 3ull) != 0;
          st = (unsigned int)((d >> 32) & 3ull);
          val = (int32_t)(unsigned int)d;
        }
        if (!__all_sync(0xffffffffu, ready)) continue;          // spin until the whole window is published
        const unsigned int pref = __ballot_sync(0xffffffffu, st == 2);
        const int stop = pref ? __ffs(pref) - 1 : 32;           // nearest predecessor holding a prefix
        int32_t part = lane <= stop ? val : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        excl += part;
        if (pref) break;
        first -= 32;
      }
      if (lane == 0) desc[tile] = tag | (2ull << 32) | (unsigned int)(excl + agg);
    }
    if (lane == 0) s_prefix = excl;
  }
  __syncthreads();
  int32_t run = s_prefix + warp_tot[wid] + inc - tsum;
#pragma unroll
  for (int i = 0; i < SCAN_I; ++i) {
    long long g = base + i;
    if (g < n_out) out[g] = run;
    run += v[i];
  }
}

struct ScanState {
  unsigned long long *desc = nullptr;
  unsigned int *ticket = nullptr;
  size_t cap = 0;
  unsigned int base = 0;
  unsigned long long epoch = 0;
};
static std::mutex g_scan_mu;
static std::vector<std::pair<StreamKey, ScanState *>> g_scan;

static int scan_launch(const int32_t *in, int32_t *out, long long n, bool flags, cudaStream_t s) {
  const long long n_out = n + 1;
  const int nb = cdiv(n_out, SCAN_B);
  ScanState *st = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_scan_mu);
    for (auto &e : g_scan)
      if (e.first == stream_key(s)) { st = e.second; break; }
    if (!st) {
      st = new ScanState();
      g_scan.emplace_back(stream_key(s), st);
    }
  }
  // (a stream's calls are made by one host thread at a time: the state below is only touched by that thread)
  if (!st->ticket) {
    SCN_TRY(dev_alloc_t(&st->ticket, 4, s));
    SCN_CUDA(cudaMemsetAsync(st->ticket, 0, 16, s));
    st->base = 0;
  }
  if (st->cap < (size_t)nb) {
    if (st->desc) dev_free(st->desc, s);
    st->desc = nullptr;
    size_t cap = (size_t)nb * 2;
    if (cap < 4096) cap = 4096;
    SCN_TRY(dev_alloc_t(&st->desc, cap, s));
    SCN_CUDA(cudaMemsetAsync(st->desc, 0, cap * 8, s));     // epoch 0 is never used by a scan
    st->cap = cap;
  }
  st->epoch = (st->epoch + 1) & ((1ull << 30) - 1);
  if (st->epoch == 0) {                                      // wrapped (2^30 scans): restart from clean words
    SCN_CUDA(cudaMemsetAsync(st->desc, 0, st->cap * 8, s));
    st->epoch = 1;
  }
  if (flags)
    SCN_LAUNCH((k_scan_lookback<true>), nb, SCAN_T, 0, s, in, out, n, n_out, st->desc, st->ticket, st->base, st->epoch);
  else
    SCN_LAUNCH((k_scan_lookback<false>), nb, SCAN_T, 0, s, in, out, n, n_out, st->desc, st->ticket, st->base, st->epoch);
  st->base += (unsigned int)nb;
  SCN_LAUNCHED();
  return 0;
}

int exclusive_scan_i32(const int32_t *in, int32_t *out, long long n, cudaStream_t s) {
  return scan_launch(in, out, n, false, s);
}
int exclusive_scan_flags_i32(const int32_t *in, int32_t *out, long long n, cudaStream_t s) {
  return scan_launch(in, out, n, true, s);
}

// ---------------------------------------------------------------------------------------
// stable LSD radix sort, 9-bit digits (27-bit neighbour masks = 3 passes); every warp owns one contiguous chunk
// ---------------------------------------------------------------------------------------
// 4 warps x 512 keys per block: a 278k-row scale gives 136 blocks (chunks of 2048 keys left 131 of
// the 148 SMs idle; chunks of 128 keys - 8 warps per block - halve the scatter's time but quadruple the
// bin-offset table the three kernels of a pass move: rulebook build 3.24 -> 3.42 ms, measured and reverted)
constexpr int RS_WARPS = 4, RS_CHUNK = 512, RS_BITS = 9, RS_BINS = 1 << RS_BITS;

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rs_hist(const uint32_t *__restrict__ keys, int32_t *__restrict__ H, long long n, int shift,
          int n_chunks) {
  pdl_sync();
  __shared__ int32_t hist[RS_WARPS][RS_BINS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int chunk = blockIdx.x * RS_WARPS + w;
  for (int d = lane; d < RS_BINS; d += 32) hist[w][d] = 0;
  __syncwarp();
  if (chunk < n_chunks) {
    // all 16 keys of the lane first (independent coalesced loads: the kernel runs ~4 warps per SM, so a load inside
    // the loop was a full memory latency per iteration), then the shared-memory atomics
    const long long b = (long long)chunk * RS_CHUNK;
    uint32_t kreg[RS_CHUNK / 32];
#pragma unroll
    for (int i = 0; i < RS_CHUNK / 32; ++i) {
      const long long g = b + i * 32 + lane;
      kreg[i] = g < n ? keys[g] : 0u;
    }
#pragma unroll
    for (int i = 0; i < RS_CHUNK / 32; ++i)
      if (b + i * 32 + lane < n) atomicAdd(&hist[w][(kreg[i] >> shift) & (RS_BINS - 1)], 1);
  }
  __syncwarp();
  if (chunk < n_chunks)
    for (int d = lane; d < RS_BINS; d += 32) H[(long long)d * n_chunks + chunk] = hist[w][d];
}

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rs_scatter(const uint32_t *__restrict__ keys, const int32_t *__restrict__ vals,
             uint32_t *__restrict__ keys_out, int32_t *__restrict__ vals_out,
             const int32_t *__restrict__ Hs, long long n, int shift, int n_chunks) {
  pdl_sync();
  __shared__ int32_t cnt[RS_WARPS][RS_BINS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int chunk = blockIdx.x * RS_WARPS + w;
  if (chunk >= n_chunks) return;
  {   // the chunk's 512 bin offsets: 16 strided loads per lane, all in flight before the first shared-memory store
      // (ncu: one dependent load per iteration was 46 % of the kernel's stall samples)
    int32_t off[RS_BINS / 32];
#pragma unroll
    for (int i = 0; i < RS_BINS / 32; ++i) off[i] = __ldg(Hs + (long long)(i * 32 + lane) * n_chunks + chunk);
#pragma unroll
    for (int i = 0; i < RS_BINS / 32; ++i) cnt[w][i * 32 + lane] = off[i];
  }
  __syncwarp();
  const long long b = (long long)chunk * RS_CHUNK;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t kreg[RS_CHUNK / 32];
  int32_t vreg[RS_CHUNK / 32];
#pragma unroll
  for (int i = 0; i < RS_CHUNK / 32; ++i) {               // (see k_rs_hist: loads hoisted out of the ranking loop)
    const long long g = b + i * 32 + lane;
    kreg[i] = g < n ? keys[g] : 0u;
    vreg[i] = g < n ? vals[g] : 0;
  }
#pragma unroll
  for (int i = 0; i < RS_CHUNK / 32; ++i) {
    const long long g = b + i * 32 + lane;
    const bool valid = g < n;
    const uint32_t key = kreg[i];
    const int32_t val = vreg[i];
    const int d = valid ? (int)((key >> shift) & (RS_BINS - 1)) : 0;
    // lanes holding the same digit, by one ballot per digit bit: __match_any_sync serialises over the DISTINCT values
    // in the warp (~70 cycles each; 9-bit digits are mostly distinct), which made a round cost ~2300 cycles and every
    // launch of this kernel ~20 us whatever its size (ncu launch list, round 2)
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < RS_BITS; ++b) {
      const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1);
      peers &= ((d >> b) & 1) ? bal : ~bal;
    }
    const int rank = __popc(peers & lt);
    if (valid) {
      const int pos = cnt[w][d] + rank;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncwarp();
    if (valid && rank == 0) cnt[w][d] += __popc(peers);
    __syncwarp();
  }
}

int radix_sort_pairs(uint32_t *keys, int32_t *vals, long long n, int bits, cudaStream_t s, bool keys_out) {
  if (n <= 1 || bits <= 0) return 0;
  const int passes = (bits + RS_BITS - 1) / RS_BITS;
  const int n_chunks = cdiv(n, RS_CHUNK);
  const int nb = cdiv(n_chunks, RS_WARPS);
  uint32_t *k2 = nullptr;
  int32_t *v2 = nullptr, *H = nullptr;
  SCN_TRY(dev_alloc_t(&k2, (size_t)n, s));
  SCN_TRY(dev_alloc_t(&v2, (size_t)n, s));
  SCN_TRY(dev_alloc_t(&H, (size_t)RS_BINS * n_chunks + 1, s));
  uint32_t *ka = keys, *kb = k2;
  int32_t *va = vals, *vb = v2;
  for (int p = 0; p < passes; ++p) {
    SCN_LAUNCH(k_rs_hist, nb, RS_WARPS * 32, 0, s, ka, H, n, p * RS_BITS, n_chunks);
    SCN_LAUNCHED();
    SCN_TRY(exclusive_scan_i32(H, H, (long long)RS_BINS * n_chunks, s));
    SCN_LAUNCH(k_rs_scatter, nb, RS_WARPS * 32, 0, s, ka, va, kb, vb, H, n, p * RS_BITS, n_chunks);
    SCN_LAUNCHED();
    uint32_t *tk = ka; ka = kb; kb = tk;
    int32_t *tv = va; va = vb; vb = tv;
  }
  if (ka != keys) {
    if (keys_out) SCN_CUDA(cudaMemcpyAsync(keys, ka, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    SCN_CUDA(cudaMemcpyAsync(vals, va, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
  }
  dev_free(k2, s);
  dev_free(v2, s);
  dev_free(H, s);
  return 0;
}

__global__ void k_scale(float *y, float a, long long n) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long st = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += st) y[i] *= a;
}

}  // namespace scn

extern "C" {
const char *scn_last_error(void) { return scn::g_err; }
int scn_version(void) { return 100; }
int scn_n_rulebook_bits(void) { return 32; }
int64_t scn_launch_count(void) { return scn::g_launches.load(); }
int scn_set_pdl(int mode) {
  scn::g_pdl = mode;
  return 0;
}
int scn_prof_enable(int on) {
  scn::g_prof_on = on != 0;
  return 0;
}
int scn_prof_read(int cls, double out[4]) {
  using namespace scn;
  SCN_CHECK(cls >= 0 && cls < PROF_N && out, "bad profile class %d", cls);
  SCN_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  out[0] = out[1] = out[2] = out[3] = 0;
  std::vector<ProfSeg> keep;
  for (auto &sg : g_prof_segs) {
    if (sg.cls != cls) { keep.push_back(sg); continue; }
    float ms = 0;
    cudaEventElapsedTime(&ms, sg.a, sg.b);
    out[0] += 1; out[1] += ms; out[2] += sg.bytes; out[3] += sg.flops;
    cudaEventDestroy(sg.a);
    cudaEventDestroy(sg.b);
  }
  g_prof_segs.swap(keep);
  return 0;
}
int scn_scale_inplace(float *y, float alpha, int64_t n, void *stream) {
  if (n <= 0) return 0;
  int nb = scn::cdiv(n, 256);
  int cap = scn::num_sms() * 8;
  if (nb > cap) nb = cap;
  SCN_LAUNCH(scn::k_scale, nb, 256, 0, (cudaStream_t)stream, y, alpha, n);
  SCN_LAUNCHED();
  return 0;
}
}
