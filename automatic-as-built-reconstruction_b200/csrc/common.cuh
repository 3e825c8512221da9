// common.cuh - shared plumbing of libscn_b200: error reporting, stream-ordered device
// memory, launch counting, and the small device primitives (scan, warp reductions) every
// kernel file uses.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>
#include <utility>

namespace scn {

// ---- errors ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define SCN_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      scn::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), \
                     cudaGetErrorString(_e));                                           \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

#define SCN_CHECK(cond, ...)                                                            \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      scn::set_error(__VA_ARGS__);                                                      \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

#define SCN_TRY(expr)                                                                   \
  do {                                                                                  \
    int _r = (expr);                                                                    \
    if (_r) return _r;                                                                  \
  } while (0)

// count + check a kernel launch
#define SCN_LAUNCHED()                                                                  \
  do {                                                                                  \
    scn::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    SCN_CUDA(cudaGetLastError());                                                       \
  } while (0)

// ---- stream-ordered device memory -------------------------------------------------------
// cudaMallocAsync on the device's default pool with the release threshold raised, so the
// rebuild-every-step Metadata never goes back to the driver after the first iteration.
int dev_alloc(void **p, size_t bytes, cudaStream_t s);
void dev_free(void *p, cudaStream_t s);

template <typename T>
inline int dev_alloc_t(T **p, size_t n, cudaStream_t s) {
  return dev_alloc((void **)p, (n ? n : 1) * sizeof(T), s);
}

// Per-op scratch (packed weights, split-K partials, BN sums, dW partials): one grow-only buffer per
// (stream, slot).  Ops on a stream are serialised, so the next op may overwrite the previous op's
// scratch; this replaces a cudaMallocAsync/cudaFreeAsync pair per op (~3 us of host time each).
enum WsSlot { WS_PACKED_W = 0, WS_SPLITK = 1, WS_BN = 2, WS_DW_PARTIAL = 3, WS_WT = 4, WS_PAD_X = 5, WS_PAD_W = 6, WS_CHAIN = 7, WS_SLOTS = 8 };
int workspace(void **p, int slot, size_t bytes, cudaStream_t s);
template <typename T>
inline int workspace_t(T **p, int slot, size_t n, cudaStream_t s) {
  return workspace((void **)p, slot, (n ? n : 1) * sizeof(T), s);
}

// Two zero-initialised ints per stream for the persistent kernels' dynamic work distribution (the
// kernels leave them zero on exit).
int sched_counters(int **p, cudaStream_t s);

// A companion stream of `s` for work that is independent of what `s` does next (the weight gradient
// of a layer next to its input gradient).  fork: side waits for everything queued on s so far;
// join: s waits for the side stream.
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; };
int side_stream(cudaStream_t s, SideStream **out);
int side_fork(cudaStream_t s, SideStream *ss);
int side_join(cudaStream_t s, SideStream *ss);

// Key of all per-stream state (workspaces, scheduler counters, companion streams, scan descriptors, BN sum buffers):
// the stream handle AND the current device - the default-stream handle is the same value on every device, so two
// devices driven from one process must not share (or free) each other's buffers.
struct StreamKey {
  int dev;
  cudaStream_t s;
  bool operator==(const StreamKey &o) const { return dev == o.dev && s == o.s; }
};
inline StreamKey stream_key(cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  return StreamKey{dev, s};
}

// pinned host scratch for count read-backs (per thread)
int64_t *host_scratch(size_t n_int64);

// ---- optional per-kernel-class timing (bench.py roofline; off by default) --------------------
// A profiled region is bracketed by CUDA events on the launching stream; scn_prof_read sums the
// elapsed times and the ALGORITHMIC bytes / flops the callers attribute to each region.
enum ProfClass { PROF_GEMM = 0, PROF_DW = 1, PROF_BN = 2, PROF_RULES = 3, PROF_IO = 4, PROF_N = 5 };
extern bool g_prof_on;
void prof_begin_(int cls, cudaStream_t s);
void prof_end_(int cls, cudaStream_t s, double bytes, double flops);
static inline void prof_begin(int cls, cudaStream_t s) { if (g_prof_on) prof_begin_(cls, s); }
static inline void prof_end(int cls, cudaStream_t s, double bytes, double flops) {
  if (g_prof_on) prof_end_(cls, s, bytes, flops);
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- kernel launches: programmatic dependent launch ---------------------------------------------
// A backbone step is ~1000 small DEPENDENT kernels; between two of them on one stream the GPU idles for the launch
// latency (measured, tools/pdl_probe.cu: 4.0 us per trivial kernel, 2.5 us with programmatic stream serialization).
// Every kernel of this library therefore starts with pdl_sync(): griddepcontrol.launch_dependents (the NEXT kernel of
// the stream may be scheduled as soon as every CTA of this one has started) followed by griddepcontrol.wait (block
// until the PREVIOUS kernel has completed and its writes are visible).  No kernel touches global memory before its
// wait, and every thread executes it, so "kernel N complete" still implies "kernel N-1 complete": events, memcpys
// and foreign (torch) kernels that follow in the stream see exactly the ordinary stream order.
// g_pdl (SCN_B200_PDL): 0 = plain launches, 1 = every kernel, 2 = every kernel except the persistent tensor-core
// GEMMs (whose early-scheduled CTAs would hold the SMs of a finishing GEMM against the companion streams).
extern int g_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(int pdl_class, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (g_pdl == 1 || (g_pdl == 2 && pdl_class == 0)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// SCN_LAUNCH(kernel, grid, block, smem, stream, args...): class 0 (ordinary kernel)
#define SCN_LAUNCH(kernel, ...) scn::launch_k(0, kernel, __VA_ARGS__)
#define SCN_LAUNCH_GEMM(kernel, ...) scn::launch_k(1, kernel, __VA_ARGS__)
int num_sms();

// ---- device-wide exclusive scan of int32 ----------------------------------------------
// out[i] = sum_{j<i} in[j]; out has n+1 entries (out[n] = total). in may alias out.
int exclusive_scan_i32(const int32_t *in, int32_t *out, long long n, cudaStream_t s);
// same with in[j] replaced by (in[j] >= 0 ? 1 : 0): positions of the present partners of a neighbour table
int exclusive_scan_flags_i32(const int32_t *in, int32_t *out, long long n, cudaStream_t s);

// ---- stable LSD radix sort of (key uint32, value int32), keys limited to `bits` ----------
// keys_out = false: only the permuted values are needed, the key buffer is left in an unspecified order
int radix_sort_pairs(uint32_t *keys, int32_t *vals, long long n, int bits, cudaStream_t s, bool keys_out = true);

// ---- device helpers ------------------------------------------------------------------------
#ifdef __CUDACC__
// first statement of every kernel (see launch_k); a no-op for a kernel launched without the attribute
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint64_t mix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}
#endif

}  // namespace scn
