// conv_tc.cu - tcgen05 / TMEM gather-GEMM for the sparse convolutions (sm_100a only).
//
// Same contraction as conv.cu (reference: CPU/Convolution.cpp:46-185, CPU/Deconvolution.cpp:8-77)
//   Y[stationary rows] = bias + sum_k X[partner_k rows] @ W[k]
// executed on the 5th-generation tensor cores:
//   * one CTA owns a tile of 128 stationary rows; the fp32 accumulator [128 x N] lives in TMEM;
//   * for every kernel offset active in the tile and every 32-wide slice of the reduction dim, the
//     128 partner rows are GATHERED straight from global memory into shared memory with 16-byte
//     cp.async (zero-fill where a row has no partner) in the canonical K-major no-swizzle UMMA
//     layout, while the matching weight slice - pre-packed into the same layout by k_pack_weights -
//     arrives as one cp.async.bulk (TMA, mbarrier complete_tx);
//   * one thread issues tcgen05.mma.kind::tf32 (M=128, N=Cout, K=8) on the staged slices and
//     tcgen05.commit releases the stage; a 4-stage ring keeps the gathers of steps s+1, s+2 in flight
//     while the tensor core works on step s;
//   * epilogue: tcgen05.ld TMEM -> registers -> (+bias) -> each stationary row written exactly once.
// fp32 features are fed unconverted (kind::tf32 reads the upper 19 bits); weights are rounded to
// tf32 when packed.  Accumulation is fp32.
#include "conv.cuh"
#include "../../include/scn_b200.h"
#include <cuda_bf16.h>
#include <algorithm>
#include <mutex>
#include <vector>

namespace scn {
namespace tc {

constexpr int KC = 32;                      // reduction elements per pipeline step
constexpr int NCORE = KC / 4;               // 16-byte pieces (4 tf32) per row and step
constexpr int A_STAGE = TILE_M * 128;       // 128 rows x 128 bytes, K-major SWIZZLE_128B

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
// arrive on `bar` once all cp.async issued so far by this thread have landed (pending count +1 now, -1
// then: pair it with a plain mbar_arrive, as cutlass::PipelineAsync's cp.async producers do)
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// the same for roles that wait long and are not on the critical path (epilogue warps waiting for a finished
// accumulator, the loaders waiting for a free slot): back off between probes - their spin loops were a quarter of
// the instructions the gather-GEMM issued (ncu source view) and compete with the producers for issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) break;
    __nanosleep(64);
  }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
// A operand from tensor memory (row i of A in TMEM lane i, the K = 8 values in 8 consecutive columns)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor): start>>4 [0,14),
// leading byte offset>>4 [16,30), stride byte offset>>4 [32,46), version=1 [46,48), layout [61,64)=0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major (or MN-major
// when the flags are set), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---------------------------------------------------------------------------------------------
// weight packing: W fp32 [K][Cin][Cout] -> Wp[k][chunk c][n][32], the exact shared-memory image of
// the K-major SWIZZLE_128B B operand: row n of a 32-wide reduction slice is 128 bytes whose 16-byte
// chunks are XOR-ed with n%8.  transpose=0: reduction dim = Cin, n = Cout (forward);
// transpose=1: reduction dim = Cout, n = Cin (dX = dY @ W[k]^T).  Values are rounded to tf32.
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// the gather-GEMM kernel
// ---------------------------------------------------------------------------------------------
#if defined(SCN_EXPERIMENT_TRACE) || defined(SCN_EXPERIMENT_TRACE_DW)
// per-role clock64 trace of CTA 0 (developer experiment, tools/gemm_trace.py): role r appends to g_trace[r]
__device__ unsigned long long g_trace[5][8192];
__device__ int g_trace_n[5];
#define SCN_TRACE_(role, tag)                                                                     \
  do {                                                                                            \
    if (blockIdx.x == 0) {                                                                        \
      const int _i = g_trace_n[role];                                                             \
      if (_i < 8190) { g_trace[role][_i] = ((unsigned long long)(tag) << 56) | (clock64() & 0xffffffffffffffull); g_trace_n[role] = _i + 1; } \
    }                                                                                             \
  } while (0)
#endif
#ifdef SCN_EXPERIMENT_TRACE
#define SCN_TRACE(role, tag) SCN_TRACE_(role, tag)
#else
#define SCN_TRACE(role, tag) do { } while (0)
#endif
#ifdef SCN_EXPERIMENT_TRACE_DW
#define SCN_TRACE_DW(role, tag) SCN_TRACE_(role, tag)
#else
#define SCN_TRACE_DW(role, tag) do { } while (0)
#endif
constexpr int MS = 3;                        // tile-metadata slots, at most (producers may run ~2 tiles ahead of the epilogue;
                                             // heavy layers run with 2 when the third would cost a gathered-row stage)
constexpr int NSA_MAX = 8, NSB_MAX = 4;      // ring depths: A (gathered rows) / B (weight slices)
// warp roles: 0-7 gather (one warp issues its cp.async chain at ~105 cycles per copy, so the ISSUE rate of
// four warps bounded the step - tools/gemm_trace.py, dw_trace.py), 8 MMA, 9 weight loader, 10 metadata
// loader, 11 idle, 12-15 epilogue (TMEM lanes 32*(warp%4)), 16-19 converters (3xTF32 only)
constexpr int GP_W = 8, MMA_W = 8, WL_W = 9, META_W = 10, EPI_W = 12, CONV_W = 16;
constexpr int NT_P = 16 * 32;
constexpr int NT_P3 = 20 * 32;
__host__ __device__ constexpr int gemm_threads(int mode) { return mode ? NT_P3 : NT_P; }
constexpr int NLO = 2;                       // stages of low-order halves (3xTF32 mode, weight-gradient kernel: shared memory)
// 3xTF32 gather-GEMM: the low-order halves of the gathered rows live in TENSOR MEMORY (row r in lane r, the
// 32 values of a step in 32 columns; tools/tmem_a_probe.cu) and feed tcgen05.mma as its A operand from there.
// Against a shared-memory side ring this frees 32 KB (a third weight stage: the weight ring's depth bounded
// the step) and a quarter of the step's shared-memory traffic.
#ifndef SCN_X3_LO_SMEM
constexpr bool LO_TMEM = true;
#else
constexpr bool LO_TMEM = false;
#endif
constexpr int LO_COLS = 128;                 // TMEM columns reserved for the low-order stages (4 x 32)
// 3xTF32: the converters also copy the gathered rows THEMSELVES into tensor memory (another 4 x 32 columns), so all
// three MMAs of a K = 8 slice take their A operand from TMEM.  The kernel is shared-memory-bandwidth bound in this
// mode (per step: 80 KB of operand reads by the MMAs + 48 KB written by the copies + 16 KB read by the converters
// = 1125 cycles at 128 B/clk against 768 cycles of MMAs); two of the three A reads (32 KB per step) go away.
#ifndef SCN_X3_HI_SMEM
constexpr bool HI_TMEM = LO_TMEM;
#else
constexpr bool HI_TMEM = false;
#endif

struct Smem {
  // offsets (bytes) into the dynamic shared memory block, computed identically on host and device
  int a, alo, b, stage, meta, meta_bytes, bars, tmem_slot, total;
  // mode 0: tf32, 1: 3xTF32 (a weight stage = hi slice + lo slice), 2: bf16 (a weight stage = N x 32 bf16)
  __host__ __device__ static int b_stage_bytes(int N, int mode) { return mode == 1 ? 2 * NCORE * N * 16 : (mode == 2 ? N * 64 : NCORE * N * 16); }
  __host__ __device__ Smem(int N, int K, int nsa, int nsb, int mode, int ms) {
    a = 0;
    alo = a + nsa * A_STAGE;                // 3xTF32 with the low-order halves in shared memory: NLO stages
    b = alo + ((mode == 1 && !LO_TMEM) ? NLO * A_STAGE : 0);
    stage = b + nsb * b_stage_bytes(N, mode);
    meta = stage + 4 * 4096;                // 4 epilogue warps x 4 KB transpose tiles
    meta_bytes = K * TILE_M * 4 + TILE_M * 4 + 64;      // sIdx[K][128], sPerm[128], {nE, pad, sK[32]}
    bars = meta + ms * meta_bytes;
    tmem_slot = bars + (2 * NSA_MAX + 2 * NSB_MAX + 2 * MS + 4 + 4) * 8;
    total = tmem_slot + 16;
  }
};

// one lane of the (fully converged) warp: elect.sync
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_128B operand (rows of 128 bytes, 16-byte chunks XOR-ed with row%8): 8-row groups
// 1024 B apart, layout_type 2; K advances by adding bytes to the start address (tools/umma_probe.py)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return make_desc(saddr, 16, 1024) | (2ull << 61);
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// Persistent, warp-specialised gather-GEMM.  A CTA walks work items (tile, split) round-robin; five
// roles run decoupled through mbarrier rings, so the gathers of the next tiles, the MMAs of the
// current one and the write-back of the previous one overlap:
//   warp 10      metadata loader: tile's gather lists / row permutation / offset list -> smem slot
//   warps 0-3    A producers: 16-byte cp.async gathers of the partner rows (zero-fill) followed by
//                cp.async.mbarrier.arrive: the stage's barrier completes when the copies have landed,
//                the producer never waits for them and runs up to DEPTH+2 steps ahead of the MMA
//   warp 5       weight loader: one cp.async.bulk (TMA) per step of the packed B slice
//   warp 4       MMA issuer: tcgen05.mma.kind::tf32 into one of two TMEM accumulators
//   warps 6-9    epilogue: tcgen05.ld -> (+bias) -> each stationary row written once
// Rings: fullA/emptyA per gathered-row stage (DEPTH+2 deep: the gathers are latency-bound, so what
// counts is bytes in flight), fullB/emptyB per weight-slice stage (2-3 deep: bulk copies of L2-resident
// weights), meta_full/meta_empty per metadata slot, tmem_full/tmem_empty per accumulator.
//
// X3 (3xTF32, the library's fp32 mode): x = hi + lo with hi = the 19 bits kind::tf32 reads, lo = x - hi
// (exact), w = whi + wlo (both rounded to tf32 when packed); Y += hi*whi + lo*whi + hi*wlo - the
// dropped lo*wlo term is 2^-20 relative.  Warps 11-14 (converters) compute lo from every landed
// stage into a 2-stage side ring; the MMA warp issues three instructions per K = 8 slice.  N is the
// column width of a work item (<= 128 in X3 mode: wider outputs are split over work items), ldn the
// row stride of Y.
//
// MODE 2 (bf16 operands, fp32 accumulation): the converter warps turn every landed stage into bf16 pairs in
// tensor memory (row r -> lane r, 32 values -> 16 columns) and the MMA warp issues two kind::f16 instructions
// (K = 16) per step with that A operand from TMEM; the weight stage is the no-swizzle K-major bf16 image
// (core matrices of 8 rows x 16 B, N x 64 bytes per step) - layouts confirmed by tools/tmem_a_bf16_probe.cu.
// per-role stall accounting of CTA 0 (developer experiment, tools/gemm_stalls.py): cycles every warp's lane 0
// spends inside each kind of mbarrier wait (0 metadata, 1 emptyA, 2 fullA, 3 fullB, 4 converted stage, 5 TMEM
// accumulator) and, in slot 7, from kernel start to the end of its role
#ifdef SCN_EXPERIMENT_STALLS
__device__ long long g_stall[20][8];
__device__ long long g_cta[160][4];      // per CTA (MMA thread): steps, items, globaltimer at start / end of its role
#define SCN_STALL_DECL long long stall_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long stall_t0_ = clock64()
#define MBW(bar, parity, id) do { const long long t0_ = clock64(); mbar_wait(bar, parity); stall_[id] += clock64() - t0_; } while (0)
#define MBWS(bar, parity, id) do { const long long t0_ = clock64(); mbar_wait_relaxed(bar, parity); stall_[id] += clock64() - t0_; } while (0)
#define SCN_STALL_T0 long long ts_ = clock64()
#define SCN_STALL_ADD(id) do { const long long tn_ = clock64(); stall_[id] += tn_ - ts_; ts_ = tn_; } while (0)
#define SCN_STALL_WRITE                                                                  \
  do {                                                                                   \
    stall_[7] = clock64() - stall_t0_;                                                   \
    if (blockIdx.x == 0 && lane == 0)                                                    \
      for (int i_ = 0; i_ < 8; ++i_) g_stall[warp][i_] = stall_[i_];                     \
  } while (0)
#else
#define SCN_STALL_DECL do { } while (0)
#define MBW(bar, parity, id) mbar_wait(bar, parity)
#define MBWS(bar, parity, id) mbar_wait_relaxed(bar, parity)
#define SCN_STALL_T0 do { } while (0)
#define SCN_STALL_ADD(id) do { } while (0)
#define SCN_STALL_WRITE do { } while (0)
#endif
template <int DEPTH, int MODE>
__global__ void __launch_bounds__(gemm_threads(MODE), 1)
k_osgemm_tf32(const float *__restrict__ X, const float *__restrict__ Wp, const float *__restrict__ bias,
              float *__restrict__ Y, int Kd, int N, int ldn, int K, long long n_rows, TileView tb, uint32_t acc_cols,
              float *__restrict__ Ypart, int n_items, int splits, int NSB, int *__restrict__ sched, int ms, int static_first) {
  // DEPTH + 2 stages of gathered rows (16 KB each)
  constexpr int NSA = DEPTH + 2;
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr bool X3 = MODE == 1, BF = MODE == 2, CONV = MODE != 0;
  const Smem L(N, K, NSA, NSB, MODE, ms);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.tmem_slot);
  const uint32_t a_base = smem_u32(smem + L.a), b_base = smem_u32(smem + L.b);
  const uint32_t bar_fullA = smem_u32(smem + L.bars);
  const uint32_t bar_emptyA = bar_fullA + NSA_MAX * 8;
  const uint32_t bar_fullB = bar_emptyA + NSA_MAX * 8;
  const uint32_t bar_emptyB = bar_fullB + NSB_MAX * 8;
  const uint32_t bar_mfull = bar_emptyB + NSB_MAX * 8;
  const uint32_t bar_mempty = bar_mfull + MS * 8;
  const uint32_t bar_tfull = bar_mempty + MS * 8;
  const uint32_t bar_tempty = bar_tfull + 2 * 8;
  const uint32_t bar_fullL = bar_tempty + 2 * 8;
  const uint32_t alo_base = smem_u32(smem + L.alo);
  const int B_SLICE = NCORE * N * 16;               // one packed weight slice (hi or lo)
  const int B_STAGE = Smem::b_stage_bytes(N, MODE);
  const int item_stride = tb.n_tiles * splits;      // work items per column block
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  SCN_STALL_DECL;
  const int n_tiles = tb.n_tiles;
  const int kchunks = Kd / KC;

  auto meta_idx = [&](int slot) { return reinterpret_cast<int32_t(*)[TILE_M]>(smem + L.meta + slot * L.meta_bytes); };
  auto meta_perm = [&](int slot) { return reinterpret_cast<int32_t *>(smem + L.meta + slot * L.meta_bytes + K * TILE_M * 4); };
  auto meta_hdr = [&](int slot) { return reinterpret_cast<int32_t *>(smem + L.meta + slot * L.meta_bytes + (K + 1) * TILE_M * 4); };
  // steps of work item `item` given the tile's entry count
  auto item_steps = [&](int item, int nE) {
    const int split = (item % item_stride) / n_tiles, all_steps = nE * kchunks;
    return all_steps > split ? (all_steps - split + splits - 1) / splits : 0;
  };

  if (tid == 0) {
    for (int i = 0; i < NSA; ++i) {
#ifdef SCN_EXPERIMENT_FEWARRIVE     // timing experiment only (wrong results): one arrival per producer WARP, copies untracked
      mbar_init(bar_fullA + i * 8, GP_W);
#else
      mbar_init(bar_fullA + i * 8, GP_W * 32);
#endif
      mbar_init(bar_emptyA + i * 8, 1);
    }
    for (int i = 0; i < NSB; ++i) {
      mbar_init(bar_fullB + i * 8, 1);
      mbar_init(bar_emptyB + i * 8, 1);
    }
    for (int i = 0; i < ms; ++i) {
      mbar_init(bar_mfull + i * 8, 1);
      mbar_init(bar_mempty + i * 8, GP_W + 6 + (CONV ? 4 : 0));   // producer + 4 epilogue warps + MMA + weight loader (+ 4 converters)
    }
    if (CONV)
      for (int i = 0; i < 4; ++i) mbar_init(bar_fullL + i * 8, 4);     // one arrival per converter warp
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + i * 8, 1);
      mbar_init(bar_tempty + i * 8, 4);      // 4 epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // accumulators: 2 x acc_cols; 3xTF32 with the low-order halves in TMEM: the whole 512 columns
  const uint32_t tmem_cols = ((X3 && LO_TMEM) || BF) ? 512u : 2 * acc_cols;
  constexpr int NLT = (LO_TMEM || BF) ? (NSA >= 4 ? 4 : 2) : NLO;      // low-order stages (never deeper than the A ring)
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barrier init and the TMEM allocation overlap the previous kernel's tail; nothing global before this point
  pdl_wait();

  if (warp == META_W) {
    // ===== metadata loader =====
    // Work items are handed out dynamically (one atomic per item on the launch's counter), heaviest
    // tiles first: rows are sorted by neighbour mask, so the tiles with the most active offsets sit at
    // the end of the tile book - a static round-robin left the SMs idle 35 % of the kernel (ncu:
    // smsp__cycles_active vs sm__cycles_elapsed).  sched = {next item, CTAs done}; the last CTA to
    // finish resets both for the next launch on this stream.
    for (int it = 0;; ++it) {
      const int slot = it % ms, use = it / ms;
      if (use > 0) MBWS(bar_mempty + slot * 8, (use - 1) & 1, 0);
      int q = 0;
      if (lane == 0) SCN_TRACE(0, 1);
      // a CTA's FIRST item is its block index (grid <= n_items; no atomic round trip - 0.5 us - before the kernel's first
      // gather), the following ones come off the counter, which therefore counts from gridDim.x
      if (lane == 0) q = (it == 0 && static_first) ? (int)blockIdx.x : (static_first ? (int)gridDim.x : 0) + atomicAdd(sched, 1);
      q = __shfl_sync(0xffffffffu, q, 0);
      if (lane == 0) SCN_TRACE(0, 2);
      int32_t(*sIdx)[TILE_M] = meta_idx(slot);
      int32_t *sPerm = meta_perm(slot), *hdr = meta_hdr(slot);
      if (q >= n_items) {                     // end marker for the other roles
        if (lane == 0) { hdr[1] = -1; mbar_arrive(bar_mfull + slot * 8); }
        break;
      }
      // heaviest tiles first: by the tile book's order (descending number of active offsets)
      int4 rec = make_int4(n_tiles - 1 - q % n_tiles, 0, 0, -1);
      if (tb.order) rec = tb.order[q % n_tiles];
      const int tile = rec.x;
      if (lane == 0) hdr[1] = q - q % n_tiles + tile;
      if (tb.identity) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long long r = (long long)tile * TILE_M + lane * 4 + i;
          const int v = r < n_rows ? (int)r : -1;
          sPerm[lane * 4 + i] = v;
          sIdx[0][lane * 4 + i] = v;
        }
        if (lane == 0) { hdr[0] = 1; reinterpret_cast<int8_t *>(hdr + 2)[0] = 0; }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_mfull + slot * 8);
      } else {
        // the tile's row permutation (512 B) and its nE gather lists (nE x 512 B, contiguous in the tile
        // book) come in as two bulk copies that complete the slot's barrier - the loader never waits for
        // them (a per-list load loop cost ~2 us per list and paced the whole CTA)
        const uint32_t mask = rec.w < 0 ? tb.tile_mask[tile] : (uint32_t)rec.y;
        const int e0 = rec.w < 0 ? tb.tile_off[tile] : rec.z;
        const int nE = __popc(mask);
        if (mask & (1u << lane)) reinterpret_cast<int8_t *>(hdr + 2)[__popc(mask & ((1u << lane) - 1u))] = (int8_t)lane;
        if (lane == 0) hdr[0] = nE;
        __syncwarp();
        if (lane == 0) {
          mbar_expect_tx(bar_mfull + slot * 8, (uint32_t)(nE + 1) * TILE_M * 4);
          bulk_copy_g2s(smem_u32(sPerm), tb.perm + (long long)tile * TILE_M, TILE_M * 4, bar_mfull + slot * 8);
          if (nE > 0)
            bulk_copy_g2s(smem_u32(&sIdx[0][0]), tb.entries + (long long)e0 * TILE_M, (uint32_t)nE * TILE_M * 4,
                          bar_mfull + slot * 8);
          SCN_TRACE(0, 3);
        }
      }
    }
  } else if (warp < GP_W) {
    // ===== A producers: thread (j = k-core, rows r0 + 32 i) =====
    const int j = tid & 7, r0 = tid >> 3;
    constexpr int RSTEP = GP_W * 4;          // rows covered by one pass of the producer threads
    // ring position / phase and the (offset, chunk) of a step are running counters: a run-time division per step
    // (st / kchunks: I2F + MUFU.RCP + F2I on the quarter-rate pipe, ~40 instructions) was a quarter of this role's
    // step time, and this role bounds the step in every mode (tools/gemm_stalls.py: busy 85 %)
    int stage = 0;
    uint32_t phE = 1;                        // parity to wait for on emptyA[stage]: (use - 1) & 1, no wait during use 0
    bool first_use = true;
    for (int it = 0;; ++it) {
      const int slot = it % ms;
      MBW(bar_mfull + slot * 8, (it / ms) & 1, 0);
      const int item = meta_hdr(slot)[1];
      if (item < 0) break;
      int32_t(*sIdx)[TILE_M] = meta_idx(slot);
      const int split = (item % item_stride) / n_tiles;
      const int steps = item_steps(item, meta_hdr(slot)[0]);
      if (tid == 0) SCN_TRACE(1, 1);
      int e = split / kchunks, c = split - e * kchunks;
      for (int lst = 0; lst < steps; ++lst) {
        if (!first_use) MBW(bar_emptyA + stage * 8, phE, 1);
#ifdef SCN_EXPERIMENT_NO_A
        const int use = first_use ? 0 : 1;
#endif
#ifdef SCN_EXPERIMENT_NO_A          // timing experiment only (wrong results): what if the gathers were free?
        if (use == 0)
#endif
        {
          // row r lives at (r/8)*1024 + (r%8)*128, its piece j at chunk j ^ (r%8): the 8 lanes of a row
          // fill one 128-byte line of shared memory - no bank conflicts, no padding
          const uint32_t dst = a_base + stage * A_STAGE + (r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4);
          const float *colp = X + c * KC + j * 4;
#pragma unroll
          for (int i = 0; i < TILE_M / RSTEP; ++i) {
            const int idx = sIdx[e][r0 + RSTEP * i];  // rows r0 + RSTEP i share r0 % 8
            cp_async_16(dst + i * (RSTEP * 128), colp + (long long)(idx < 0 ? 0 : idx) * Kd, idx < 0 ? 0 : 16);
          }
        }
        // the stage's barrier completes when every producer thread has passed here AND its copies have
        // landed - the thread itself never waits for them
#ifdef SCN_EXPERIMENT_FEWARRIVE
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_fullA + stage * 8);
#else
        cp_async_mbar_arrive(bar_fullA + stage * 8);
        mbar_arrive(bar_fullA + stage * 8);
#endif
        if (++stage == NSA) { stage = 0; phE ^= 1; first_use = false; }
        c += splits;
        while (c >= kchunks) { c -= kchunks; ++e; }
      }
      if (tid == 0) SCN_TRACE(1, 2);
      __syncwarp();                           // the tile's lists are no longer needed by this warp
      if (lane == 0) mbar_arrive(bar_mempty + slot * 8);
    }
  } else if (warp == MMA_W) {
    // ===== MMA issuer: the whole warp walks the loop (warp-uniform control flow: ring positions and descriptors
    // stay in uniform registers), one elected lane issues the tcgen05 instructions.  Under `if (lane == 0)` the
    // compiler wrapped every tcgen05.mma in an election loop and moved each operand through R2UR =====
    {
      const uint32_t idesc = make_idesc(TILE_M, N, 0, 0);
      // kind::f16, bf16 operands: D = f32 (bit 4), A = B = bf16 (1 at bits 7 and 10)
      const uint32_t idesc_bf = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
      // This thread is the kernel's serial bottleneck (tools/gemm_stalls.py: tcgen05.mma issue blocks for the
      // execution time of the previous instruction, so every cycle spent between two steps leaves the tensor
      // pipe idle): ring positions and phases are running counters (NSB is a run-time value: no divisions),
      // barriers are probed with test_wait before falling into the try_wait loop, descriptors are base + offset.
      int accn = 0;
      int stage = 0, stb = 0, stl = 0;
      uint32_t phA = 0, phB = 0, phL = 0;
      const uint64_t desc_hi = (uint64_t)(16u >> 4) << 16 | (uint64_t)(1024u >> 4) << 32 | (1ull << 46) | (2ull << 61);   // make_desc_sw128
      const uint64_t desc_bf = make_desc(0, (uint32_t)N * 16, 128);      // bf16 weight stage: no swizzle, K-major
#ifdef SCN_EXPERIMENT_STALLS
      long long cta_steps_ = 0, cta_items_ = 0, cta_t0_;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(cta_t0_));
#endif
      for (int it = 0, slot = 0, mph = 0;; ++it) {
        MBW(bar_mfull + slot * 8, mph, 0);
        const int item = meta_hdr(slot)[1];
        if (item < 0) break;
        const int steps = item_steps(item, meta_hdr(slot)[0]);
        if (lane == 0) mbar_arrive(bar_mempty + slot * 8);
        if (++slot == ms) { slot = 0; mph ^= 1; }
        if (lane == 0) { SCN_TRACE(2, 1); SCN_TRACE(4, steps); }
        if (steps == 0) continue;
        const int acc = accn & 1;
        if (accn >= 2) MBW(bar_tempty + acc * 8, ((accn >> 1) - 1) & 1, 5);
        tc_fence_after();
        if (lane == 0) SCN_TRACE(2, 2);
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * acc_cols;
        for (int lst = 0; lst < steps; ++lst) {
          // converter modes: fullL implies fullA (the converters wait for the landed stage before they arrive)
#ifdef SCN_EXPERIMENT_STALLS
          MBW(bar_fullB + stb * 8, phB, 3);
          if (CONV) MBW(bar_fullL + stl * 8, phL, 4);
          else MBW(bar_fullA + stage * 8, phA, 2);
#else
          {   // both probes in flight before either is looked at (probing the NEXT step's barriers one step ahead
              // instead was measured: 190 -> 211 us on the 128 -> 128 layer, the probes delay the MMA issue)
            const uint32_t bar2 = CONV ? bar_fullL + stl * 8 : bar_fullA + stage * 8, ph2 = CONV ? phL : phA;
            const bool ok1 = mbar_test(bar_fullB + stb * 8, phB), ok2 = mbar_test(bar2, ph2);
            if (!ok1) mbar_wait(bar_fullB + stb * 8, phB);
            if (!ok2) mbar_wait(bar2, ph2);
          }
#endif
          SCN_STALL_T0;
          tc_fence_after();
          SCN_STALL_ADD(0);            // (MMA role: slot 0 also counts the fences)
          if (lst == 0 && lane == 0) SCN_TRACE(2, 3);
          const uint32_t sa = a_base + stage * A_STAGE, sb = b_base + stb * B_STAGE;
          const uint32_t sl = alo_base + stl * A_STAGE;
          const uint64_t da = desc_hi | (uint64_t)((sa & 0x3FFFFu) >> 4), db = desc_hi | (uint64_t)((sb & 0x3FFFFu) >> 4);
          const uint64_t dl = desc_hi | (uint64_t)((sl & 0x3FFFFu) >> 4);
          const uint64_t db_lo = desc_hi | (uint64_t)(((sb + B_SLICE) & 0x3FFFFu) >> 4);
          if (elect_one()) {
          if (BF) {
            // A (bf16 pairs) from TMEM: 16 columns per step, 8 per K = 16 instruction; B: no-swizzle K-major,
            // core matrices N*16 B apart along K (LBO), 8-row groups 128 B apart (SBO)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              mma_bf16_ts(tmem_d, tmem_base + 2 * acc_cols + (uint32_t)(stl * 16 + kk * 8),
                          desc_bf + (uint64_t)(((sb + kk * 2 * N * 16) & 0x3FFFFu) >> 4), idesc_bf, (lst > 0 || kk > 0) ? 1u : 0u);
          }
#pragma unroll
#ifdef SCN_EXPERIMENT_NO_MMA     // timing experiment only (wrong results): one MMA per step instead of 4 / 12
          if (lst == 0) mma_tf32(tmem_d, make_desc_sw128(sa), make_desc_sw128(sb), idesc, 0u);
          for (int kk = 0; kk < 0; ++kk) {
#else
          for (int kk = 0; kk < (BF ? 0 : KC / 8); ++kk) { // K = 8 per instruction: 32 bytes (2 descriptor units) further along the 128-byte rows
#endif
            if (X3 && HI_TMEM) {
              const uint32_t ta = tmem_base + 2 * acc_cols + (uint32_t)(stl * KC + kk * 8);
              mma_tf32_ts(tmem_d, ta + LO_COLS, db + 2 * kk, idesc, (lst > 0 || kk > 0) ? 1u : 0u);
              mma_tf32_ts(tmem_d, ta, db + 2 * kk, idesc, 1u);
              mma_tf32_ts(tmem_d, ta + LO_COLS, db_lo + 2 * kk, idesc, 1u);
              continue;
            }
            mma_tf32(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (lst > 0 || kk > 0) ? 1u : 0u);
            if (X3) {
              if (LO_TMEM)
                mma_tf32_ts(tmem_d, tmem_base + 2 * acc_cols + (uint32_t)(stl * KC + kk * 8), db + 2 * kk, idesc, 1u);
              else
                mma_tf32(tmem_d, dl + 2 * kk, db + 2 * kk, idesc, 1u);
              mma_tf32(tmem_d, da + 2 * kk, db_lo + 2 * kk, idesc, 1u);
            }
          }
          // ONE commit per step: emptyA[g % NSA] completing means "the MMAs of step g are done"; the
          // weight loader and the converters wait on the same barrier for the step that last used the
          // stage they are about to refill (their rings are no deeper than NSA, so the barrier cannot
          // run two phases ahead of them)
          SCN_STALL_ADD(6);            // descriptor arithmetic + tcgen05.mma issue
#ifdef SCN_EXPERIMENT_XCOMMIT    // timing experiment: what does one more tcgen05.commit per step cost? (emptyB is unused)
          tc_commit(bar_emptyB);
#endif
          tc_commit(bar_emptyA + stage * 8);
          if (lst + 1 == steps) tc_commit(bar_tfull + acc * 8);     // the tile's last step: accumulator complete
          }
          __syncwarp();
          SCN_STALL_ADD(1);            // (MMA role: slot 1 = tcgen05.commit)
          if (++stage == NSA) { stage = 0; phA ^= 1; }
          if (++stb == NSB) { stb = 0; phB ^= 1; }
          if (++stl == NLT) { stl = 0; phL ^= 1; }
        }
        if (lane == 0) SCN_TRACE(2, 4);
        ++accn;
#ifdef SCN_EXPERIMENT_STALLS
        cta_steps_ += steps; ++cta_items_;
#endif
      }
#ifdef SCN_EXPERIMENT_STALLS
      if (blockIdx.x < 160) {
        long long t1_;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1_));
        g_cta[blockIdx.x][0] = cta_steps_; g_cta[blockIdx.x][1] = cta_items_;
        g_cta[blockIdx.x][2] = cta_t0_; g_cta[blockIdx.x][3] = t1_;
      }
#endif
    }
  } else if (warp == WL_W) {
    // ===== weight-slice loader (TMA bulk copies of the packed B operand) =====
    if (lane == 0) {
      int g = 0;
      for (int it = 0;; ++it) {
        const int slot = it % ms;
        MBW(bar_mfull + slot * 8, (it / ms) & 1, 0);
        const int item = meta_hdr(slot)[1];
        if (item < 0) break;
        const int split = (item % item_stride) / n_tiles, col0 = (item / item_stride) * N;
        const int steps = item_steps(item, meta_hdr(slot)[0]);
        const int8_t *sK = reinterpret_cast<const int8_t *>(meta_hdr(slot) + 2);
        for (int lst = 0; lst < steps; ++lst, ++g) {
          const int stage = g % NSB;
          if (g >= NSB) MBWS(bar_emptyA + ((g - NSB) % NSA) * 8, ((g - NSB) / NSA) & 1, 1);
          const int st = split + lst * splits;
          const int e = st / kchunks, c = st - e * kchunks;
          const int kw = tb.k_flip >= 0 ? tb.k_flip - (tb.k_base + sK[e]) : tb.k_base + sK[e];   // weight slice of this table offset
          const uint32_t bytes = (uint32_t)B_STAGE;
#ifdef SCN_EXPERIMENT_NO_B          // timing experiment only (wrong results): what if weight slices were free?
          if (g >= NSB) { mbar_arrive(bar_fullB + stage * 8); continue; }
#endif
          mbar_expect_tx(bar_fullB + stage * 8, bytes);
if (BF) {   // bf16 image [k][c][column block][core j][N][8]: one contiguous slice per (k, c, block)
            const int ncb = ldn / N;
            const __nv_bfloat16 *src = reinterpret_cast<const __nv_bfloat16 *>(Wp) +
                                       (((long long)kw * kchunks + c) * ncb + col0 / N) * (long long)N * KC;
            bulk_copy_g2s(b_base + stage * B_STAGE, src, bytes, bar_fullB + stage * 8);
          } else if (X3) {   // slices [k][c][hi|lo][ldn][32]: rows col0 .. col0+N of the hi and of the lo slice
            const float *src = Wp + (((long long)kw * kchunks + c) * 2 * ldn + col0) * KC;
            bulk_copy_g2s(b_base + stage * B_STAGE, src, (uint32_t)B_SLICE, bar_fullB + stage * 8);
            bulk_copy_g2s(b_base + stage * B_STAGE + B_SLICE, src + (long long)ldn * KC, (uint32_t)B_SLICE,
                          bar_fullB + stage * 8);
          } else {
            bulk_copy_g2s(b_base + stage * B_STAGE, Wp + ((long long)kw * Kd + (long long)c * KC) * N, bytes,
                          bar_fullB + stage * 8);
          }
        }
        mbar_arrive(bar_mempty + slot * 8);
      }
    }
  } else if (warp >= CONV_W) {
    // ===== converters (3xTF32 only): lo = x - (x & ~0x1fff) of every landed stage, same swizzled
    // positions, into the low-order ring; generic-proxy writes are fenced for the tensor core =====
    if (CONV) {
      const int ct = tid - CONV_W * 32;
      int g = 0;
      for (int it = 0;; ++it) {
        const int slot = it % ms;
        MBW(bar_mfull + slot * 8, (it / ms) & 1, 0);
        const int item = meta_hdr(slot)[1];
        if (item < 0) break;
        const int steps = item_steps(item, meta_hdr(slot)[0]);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_mempty + slot * 8);
        for (int lst = 0; lst < steps; ++lst, ++g) {
          const int stage = g % NSA, stl = g % NLT;
          MBW(bar_fullA + stage * 8, (g / NSA) & 1, 2);
          if (g >= NLT) MBW(bar_emptyA + ((g - NLT) % NSA) * 8, ((g - NLT) / NSA) & 1, 1);
          if (BF) {
            // thread = row ct: its 8 chunks (un-swizzled by index) -> 16 bf16 pairs -> 16 columns of TMEM lane ct
            tc_fence_after();
            const float4 *row = reinterpret_cast<const float4 *>(smem + L.a + stage * A_STAGE + (ct >> 3) * 1024 + (ct & 7) * 128);
            uint32_t w[16];
#pragma unroll
            for (int jc = 0; jc < 8; ++jc) {
              const float4 v = row[jc ^ (ct & 7)];
              const __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
              w[2 * jc] = *reinterpret_cast<const uint32_t *>(&p0);
              w[2 * jc + 1] = *reinterpret_cast<const uint32_t *>(&p1);
            }
            tmem_st16(tmem_base + 2 * acc_cols + (uint32_t)(stl * 16) + ((uint32_t)((warp & 3) * 32) << 16), w);
            tmem_st_wait();
            tc_fence_before();
          } else if (LO_TMEM) {
            // thread = row ct: its 8 chunks (un-swizzled by index) -> lo -> 32 columns of TMEM lane ct
            tc_fence_after();
            const float4 *row = reinterpret_cast<const float4 *>(smem + L.a + stage * A_STAGE + (ct >> 3) * 1024 + (ct & 7) * 128);
            uint32_t w[32];
#pragma unroll
            for (int jc = 0; jc < 8; ++jc) {
              const float4 v = row[jc ^ (ct & 7)];
              w[4 * jc + 0] = __float_as_uint(v.x); w[4 * jc + 1] = __float_as_uint(v.y);
              w[4 * jc + 2] = __float_as_uint(v.z); w[4 * jc + 3] = __float_as_uint(v.w);
            }
            // the row itself (kind::tf32 reads its upper 19 bits = hi), then lo = x - hi in place
            if (HI_TMEM)
              tmem_st32(tmem_base + 2 * acc_cols + LO_COLS + (uint32_t)(stl * KC) + ((uint32_t)((warp & 3) * 32) << 16), w);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              w[i] = __float_as_uint(__uint_as_float(w[i]) - __uint_as_float(w[i] & 0xffffe000u));
            tmem_st32(tmem_base + 2 * acc_cols + (uint32_t)(stl * KC) + ((uint32_t)((warp & 3) * 32) << 16), w);
            tmem_st_wait();
            tc_fence_before();
          } else {
            const float4 *src = reinterpret_cast<const float4 *>(smem + L.a + stage * A_STAGE);
            float4 *dst = reinterpret_cast<float4 *>(smem + L.alo + stl * A_STAGE);
#pragma unroll
            for (int u = 0; u < A_STAGE / 16 / 128; ++u) {
              const float4 v = src[ct + u * 128];
              float4 o;
              o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
              o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
              o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
              o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
              dst[ct + u * 128] = o;
            }
            fence_proxy_async();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_fullL + stl * 8);
        }
      }
    }
  } else if (warp >= EPI_W && warp < EPI_W + 4) {
    // ===== epilogue (warps 12-15): TMEM lanes 32*(warp%4).. -> registers -> global =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int accn = 0;
    for (int it = 0;; ++it) {
      const int slot = it % ms;
      MBWS(bar_mfull + slot * 8, (it / ms) & 1, 0);
      const int item = meta_hdr(slot)[1];
      if (item < 0) break;
      const int tile = item % n_tiles, split = (item % item_stride) / n_tiles, col0 = (item / item_stride) * N;
      const int steps = item_steps(item, meta_hdr(slot)[0]);
      int orow = meta_perm(slot)[row];
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_mempty + slot * 8);
      if (warp == EPI_W && lane == 0) SCN_TRACE(3, 1);
      const float *bs = bias ? bias + col0 : nullptr;
      float *yp = Y + (long long)(orow < 0 ? 0 : orow) * ldn + col0;
      if (splits > 1) {                     // partial tile, slot order, no bias
        orow = 0;
        bs = nullptr;
        yp = Ypart + (((long long)split * n_tiles + tile) * TILE_M + row) * ldn + col0;
      }
      const int acc = accn & 1;
      if (steps > 0) {
        MBWS(bar_tfull + acc * 8, (accn >> 1) & 1, 5);
        tc_fence_after();
      }
      if (warp == EPI_W && lane == 0) SCN_TRACE(3, 2);
      const uint32_t taddr = tmem_base + (uint32_t)acc * acc_cols + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        if (steps > 0) {
          if (N - c0 >= 32) {
            tmem_ld32(taddr + (uint32_t)c0, v);
          } else {                          // N = 16 (mod 32): two 8-column loads
            uint32_t w0[8], w1[8];
            tmem_ld8(taddr + (uint32_t)c0, w0);
            tmem_ld8(taddr + (uint32_t)c0 + 8, w1);
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] = w0[i]; v[8 + i] = w1[i]; }
          }
          tmem_ld_wait();
          if (c0 + 32 >= N) {               // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + acc * 8);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (N - c0 >= 32) {
          // Transpose through this warp's 4 KB staging tile so that a store instruction covers whole
          // 128-byte row segments (8 lanes x 16 B per row, 4 rows per instruction).  Writing the
          // registers out directly (thread = row) issues 32 x 16 B requests to 32 different lines per
          // instruction - measured: the kernel was epilogue-bound on exactly that.
          float4 *st = reinterpret_cast<float4 *>(smem + L.stage + q * 4096);
#pragma unroll
          for (int jc = 0; jc < 8; ++jc)      // row `lane`, chunk jc -> slot jc ^ (lane & 7): conflict-free
            st[lane * 8 + (jc ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * jc]), __uint_as_float(v[4 * jc + 1]),
                                                           __uint_as_float(v[4 * jc + 2]), __uint_as_float(v[4 * jc + 3]));
          __syncwarp();
          const int jc = lane & 7;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bs) b4 = *reinterpret_cast<const float4 *>(bs + c0 + 4 * jc);
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const int r = itr * 4 + (lane >> 3);
            const int dst_row = __shfl_sync(0xffffffffu, orow, r);
            float4 o = st[r * 8 + (jc ^ (r & 7))];
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
#ifdef SCN_EXPERIMENT_NO_EPI     // timing experiment only (wrong results): no global stores in the epilogue
            if (dst_row >= 0 && o.x == 12345.678f) {
#else
            if (dst_row >= 0) {
#endif
              float *p = splits > 1 ? Ypart + (((long long)split * n_tiles + tile) * TILE_M + q * 32 + r) * ldn
                                    : Y + (long long)dst_row * ldn;
              *reinterpret_cast<float4 *>(p + col0 + c0 + 4 * jc) = o;
            }
          }
          __syncwarp();
        } else if (orow >= 0) {             // 16-column tail (N = 16 mod 32): thread = row
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                   __uint_as_float(v[i + 3]));
            if (bs) {
              const float4 b4 = *reinterpret_cast<const float4 *>(bs + c0 + i);
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            }
            *reinterpret_cast<float4 *>(yp + c0 + i) = o;
          }
        }
      }
      if (steps > 0) ++accn;
      if (warp == EPI_W && lane == 0) SCN_TRACE(3, 3);
    }
  }
  SCN_STALL_WRITE;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
  if (tid == 0 && atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) {   // last CTA out: rearm the counters
    sched[0] = 0;
    sched[1] = 0;
  }
}

// Y[perm[slot]] = bias + sum_s Ypart[s][slot]   (fixed summation order)
__global__ void k_splitk_reduce(const float *__restrict__ Ypart, const float *__restrict__ bias, float *__restrict__ Y,
                                int N, int splits, long long n_slots, long long n_rows, TileView tb) {
  pdl_sync();
  const int q = N >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots * q) return;
  const long long slot = i / q;
  const int c = (int)(i - slot * q) * 4;
  const long long orow = tb.identity ? (slot < n_rows ? slot : -1) : tb.perm[slot];
  if (orow < 0) return;
  float4 a = bias ? *reinterpret_cast<const float4 *>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4 *>(Ypart + ((long long)s * n_slots + slot) * N + c);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  *reinterpret_cast<float4 *>(Y + orow * N + c) = a;
}

// both operand images of a weight tensor in one launch: [0,total) forward layout, [total,2 total) dX layout.
// x3 = 1 (3xTF32 mode): every (k, chunk) slice is stored twice, [k][c][hi|lo][n][32] with
// hi = tf32(w), lo = tf32(w - hi); an image then holds 2 total floats.
__device__ __forceinline__ void pack_element(const float *__restrict__ W, float *__restrict__ Wf, float *__restrict__ Wb,
                                             int K, int Cin, int Cout, int do_f, int do_b, int x3, long long i2) {
  const long long per_k = (long long)Cin * Cout, total = (long long)K * per_k;
  const int transpose = i2 >= total;
  if (transpose ? !do_b : !do_f) return;
  const long long i = transpose ? i2 - total : i2;
  const int N = transpose ? Cin : Cout;
  const int k = (int)(i / per_k);
  long long r = i - (long long)k * per_k;
  const int c = (int)(r / ((long long)KC * N));
  r -= (long long)c * KC * N;
  const int n = (int)(r >> 5), pos = (int)(r & 31);
  const int kk = c * KC + ((((pos >> 2) ^ (n & 7)) << 2) | (pos & 3));
  const float v = transpose ? W[((long long)k * Cin + n) * Cout + kk] : W[((long long)k * Cin + kk) * Cout + n];
  const float hi = to_tf32(v);
  float *dst = transpose ? Wb : Wf;
  if (x3) {
    const long long slice = (long long)KC * N;                 // floats of one (k, chunk) slice
    const long long base = (i - r) * 2;                        // slices before this one, doubled
    dst[base + r] = hi;
    dst[base + slice + r] = to_tf32(v - hi);
  } else {
    dst[i] = hi;
  }
}

__global__ void k_pack_weights_both(const float *__restrict__ W, float *__restrict__ Wf, float *__restrict__ Wb,
                                    int K, int Cin, int Cout, int do_f, int do_b, int x3) {
  pdl_sync();
  const long long total = (long long)K * Cin * Cout;
  for (long long i2 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i2 < 2 * total;
       i2 += (long long)gridDim.x * blockDim.x)
    pack_element(W, Wf, Wb, K, Cin, Cout, do_f, do_b, x3, i2);
}

// bf16 operand images (mode 2): [k][chunk c][column block][core j = 4][NW rows][8 bf16] - the no-swizzle K-major
// layout of tcgen05 (core matrix = 8 rows x 16 bytes), one contiguous NW x 32 slice per (k, c, column block);
// NW = min(N, 128).  Forward image at Wf, dX image (W[k]^T) at Wb.
__device__ __forceinline__ void pack_element_bf16(const float *__restrict__ W, __nv_bfloat16 *__restrict__ Wf,
                                                  __nv_bfloat16 *__restrict__ Wb, int K, int Cin, int Cout, int do_f, int do_b,
                                                  long long i2) {
  const long long per_k = (long long)Cin * Cout, total = (long long)K * per_k;
  const int transpose = i2 >= total;
  if (transpose ? !do_b : !do_f) return;
  const long long i = transpose ? i2 - total : i2;
  const int k = (int)(i / per_k);
  const long long r = i - (long long)k * per_k;
  const int ci = (int)(r / Cout), co = (int)(r - (long long)ci * Cout);
  const int kd = transpose ? co : ci, n = transpose ? ci : co;          // reduction index, operand row
  const int Kd = transpose ? Cout : Cin, N = transpose ? Cin : Cout;
  const int NW = N > 128 ? 128 : N, ncb = N / NW, kch = Kd / KC;
  const int c = kd / KC, kin = kd % KC, j = kin >> 3, e = kin & 7, cb = n / NW, nb = n % NW;
  const long long idx = (((((long long)k * kch + c) * ncb + cb) * 4 + j) * NW + nb) * 8 + e;
  (transpose ? Wb : Wf)[idx] = __float2bfloat16(W[i]);
}

__global__ void k_pack_weights_bf16(const float *__restrict__ W, __nv_bfloat16 *__restrict__ Wf, __nv_bfloat16 *__restrict__ Wb,
                                    int K, int Cin, int Cout, int do_f, int do_b) {
  pdl_sync();
  const long long total = (long long)K * Cin * Cout;
  for (long long i2 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i2 < 2 * total;
       i2 += (long long)gridDim.x * blockDim.x)
    pack_element_bf16(W, Wf, Wb, K, Cin, Cout, do_f, do_b, i2);
}

// every convolution weight of a layer graph in ONE launch (blockIdx.y = tensor): the per-layer pack launches were
// ~45 x 7 us of a backbone step
struct PackJob { const float *W; float *wf, *wb; int K, Cin, Cout, flags; };     // flags: 1 forward image, 2 dX image
constexpr int PACK_JOBS = 64;
struct PackBatch { PackJob job[PACK_JOBS]; };
// tf32 / 3xTF32 images: one warp per 32 x 32 tile (32 reduction indices x 32 operand rows) of one (k, chunk) slice,
// global reads and writes both in 128-byte lines.  dX image (row n = input channel, reduction = output channels):
// a weight row is contiguous along the reduction, lane l reads element l and writes it to its swizzled position
// of the same 128-byte line.  Forward image (row n = output channel, reduction = input channels): the tile is
// transposed through shared memory.  bf16 images keep the element-wise path.
__global__ void __launch_bounds__(256) k_pack_weights_batch(const __grid_constant__ PackBatch b, int x3) {
  pdl_sync();
  const PackJob &j = b.job[blockIdx.y];
  if (x3 == 2) {
    const long long total = (long long)j.K * j.Cin * j.Cout;
    for (long long i2 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i2 < 2 * total;
         i2 += (long long)gridDim.x * blockDim.x)
      pack_element_bf16(j.W, reinterpret_cast<__nv_bfloat16 *>(j.wf), reinterpret_cast<__nv_bfloat16 *>(j.wb), j.K, j.Cin, j.Cout,
                        j.flags & 1, j.flags & 2, i2);
    return;
  }
  __shared__ float tile[8][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int Cin = j.Cin, Cout = j.Cout;
  // tiles of the dX image: K x (Cout/32 chunks) x (Cin/32 row groups); of the forward image: K x (Cin/32) x (Cout/32)
  const int tiles_b = (j.flags & 2) ? j.K * (Cout >> 5) * (Cin >> 5) : 0;
  const int tiles_f = (j.flags & 1) ? j.K * (Cin >> 5) * (Cout >> 5) : 0;
  const int mul = x3 ? 2 : 1;
  for (int t = blockIdx.x * 8 + w; t < tiles_b + tiles_f; t += gridDim.x * 8) {
    if (t < tiles_b) {
      const int ng = t % (Cin >> 5), c = (t / (Cin >> 5)) % (Cout >> 5), k = t / ((Cin >> 5) * (Cout >> 5));
      // slice (k, c) of the dX image: Cin rows x 32 floats (x 2 in 3xTF32 mode: hi slice then lo slice)
      float *slice = j.wb + ((long long)(k * (Cout >> 5) + c) * Cin * KC) * mul;
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int n = ng * 32 + r;
        const float v = j.W[((long long)k * Cin + n) * Cout + c * KC + lane];
        const int pos = (((lane >> 2) ^ (n & 7)) << 2) | (lane & 3);
        const float hi = to_tf32(v);
        slice[n * KC + pos] = hi;
        if (x3) slice[(long long)Cin * KC + n * KC + pos] = to_tf32(v - hi);
      }
    } else {
      const int tf = t - tiles_b;
      const int ng = tf % (Cout >> 5), c = (tf / (Cout >> 5)) % (Cin >> 5), k = tf / ((Cout >> 5) * (Cin >> 5));
      float *slice = j.wf + ((long long)(k * (Cin >> 5) + c) * Cout * KC) * mul;
      __syncwarp();
#pragma unroll 4
      for (int r = 0; r < 32; ++r)           // row r of the tile = reduction index c*32 + r, lane = operand row
        tile[w][r][lane] = j.W[((long long)k * Cin + c * KC + r) * Cout + ng * 32 + lane];
      __syncwarp();
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {         // operand row n = ng*32 + r, lane = reduction index within the chunk
        const int n = ng * 32 + r;
        const float v = tile[w][lane][r];
        const int pos = (((lane >> 2) ^ (n & 7)) << 2) | (lane & 3);
        const float hi = to_tf32(v);
        slice[n * KC + pos] = hi;
        if (x3) slice[(long long)Cout * KC + n * KC + pos] = to_tf32(v - hi);
      }
    }
  }
}

// Packed-operand cache, keyed by the caller's identity token of the weight tensor.
struct PackEntry {
  int64_t token = 0, version = -1;
  int K = 0, Cin = 0, Cout = 0, x3 = 0;
  bool has_f = false, has_b = false;
  cudaStream_t stream = 0;
  float *wf = nullptr, *wb = nullptr;
  uint64_t last_use = 0;
  uint64_t batch = 0;            // id of the prepack batch that last filled it (see prepack_weights_batch)
};
static std::mutex g_pack_mu;
static std::vector<PackEntry *> g_pack;
static uint64_t g_pack_clock = 0;
static uint64_t g_batch_id = 0;      // current prepack batch (layer-graph forward), 0 = none active
static uint64_t g_batch_seq = 0;

static int gemm_mode(int precision) {
  return precision == SCN_PRECISION_FP32_3XTF32 ? 1 : (precision == SCN_PRECISION_BF16 ? 2 : 0);
}
static bool layout_ok(int Kd, int N) { return Kd >= KC && Kd % KC == 0 && N >= 16 && N % 16 == 0 && N <= 256; }

// returns the packed image for (transpose ? dX : forward), (re)building both when the version moved
// mode: 0 tf32, 1 3xTF32 (hi + lo slices), 2 bf16
// entry of the weight tensor `tag` sized for (K, Cin, Cout, mode); caller holds g_pack_mu
static int pack_entry(const int64_t *tag, int K, int Cin, int Cout, int x3, PackEntry **out) {
  PackEntry *e = nullptr;
  for (PackEntry *p : g_pack)
    if (p->token == tag[0]) { e = p; break; }
  if (!e) {
    if (g_pack.size() >= 1024) {   // evict the least recently used half (weights of modules long gone)
      std::sort(g_pack.begin(), g_pack.end(), [](PackEntry *a, PackEntry *b) { return a->last_use > b->last_use; });
      for (size_t i = 512; i < g_pack.size(); ++i) {
        cudaFree(g_pack[i]->wf);
        delete g_pack[i];
      }
      g_pack.resize(512);
    }
    e = new PackEntry();
    e->token = tag[0];
    g_pack.push_back(e);
  }
  const size_t total = (size_t)K * Cin * Cout;
  const size_t image = x3 == 2 ? (total + 1) / 2 : total * (x3 ? 2 : 1);   // floats: bf16 halves, 3xTF32 doubles
  if (e->K != K || e->Cin != Cin || e->Cout != Cout || e->x3 != x3) {
    cudaFree(e->wf);
    e->wf = nullptr;
    SCN_CUDA(cudaMalloc((void **)&e->wf, 2 * image * sizeof(float)));
    e->wb = e->wf + image;
    e->K = K; e->Cin = Cin; e->Cout = Cout; e->x3 = x3;
    e->version = -1;
  }
  *out = e;
  return 0;
}

// returns the packed image for (transpose ? dX : forward), (re)building both when the version moved
// mode: 0 tf32, 1 3xTF32 (hi + lo slices), 2 bf16
static int cached_pack(const int64_t *tag, const float *W, int K, int Cin, int Cout, int transpose, int x3,
                       cudaStream_t s, float **out) {
  std::lock_guard<std::mutex> lk(g_pack_mu);
  PackEntry *e = nullptr;
  SCN_TRY(pack_entry(tag, K, Cin, Cout, x3, &e));
  const size_t total = (size_t)K * Cin * Cout;
  // the forward pass (transpose == 0) always repacks: in-place edits through `.data` do not move the
  // version counter, and the forward weights must never be stale - unless the layer-graph forward that is
  // running has just packed this tensor in its batch; the dX pass of the same step reuses what its forward packed
  const bool fresh = g_batch_id != 0 && e->batch == g_batch_id && e->version == tag[1] && e->stream == s;
  if (!fresh && (!transpose || e->version != tag[1] || e->stream != s)) {
    e->has_f = layout_ok(Cin, Cout);
    e->has_b = layout_ok(Cout, Cin);
    int pb = cdiv(2 * (long long)total, 256);
    if (pb > num_sms() * 8) pb = num_sms() * 8;
    if (x3 == 2)
      SCN_LAUNCH(k_pack_weights_bf16, pb, 256, 0, s, W, reinterpret_cast<__nv_bfloat16 *>(e->wf), reinterpret_cast<__nv_bfloat16 *>(e->wb),
                                             K, Cin, Cout, e->has_f, e->has_b);
    else
      SCN_LAUNCH(k_pack_weights_both, pb, 256, 0, s, W, e->wf, e->wb, K, Cin, Cout, e->has_f, e->has_b, x3);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    SCN_CUDA(cudaGetLastError());
    e->version = tag[1];
    e->stream = s;
  }
  e->last_use = ++g_pack_clock;
  if (transpose ? !e->has_b : !e->has_f) return 1;
  *out = transpose ? e->wb : e->wf;
  return 0;
}

}  // namespace tc

// Pack the operand images of n weight tensors (forward + dX layouts) in one launch per 64 tensors and mark them
// fresh for the per-op calls that follow until prepack_weights_end() (layer-graph forward).  Tensors without a tag
// or without a tensor-core layout in either direction are skipped (their ops pack for themselves).
int prepack_weights_batch(int n, const int64_t *const *tags, const float *const *W, const int *K, const int *Cin,
                          const int *Cout, int precision, cudaStream_t s) {
  using namespace tc;
  if (precision != SCN_PRECISION_TF32 && precision != SCN_PRECISION_FP32_3XTF32 && precision != SCN_PRECISION_BF16) return 0;
  const int x3 = gemm_mode(precision);
  std::lock_guard<std::mutex> lk(g_pack_mu);
  g_batch_id = ++g_batch_seq;
  PackBatch b;
  int nj = 0;
  auto flush = [&]() -> int {
    if (nj == 0) return 0;
    SCN_LAUNCH(k_pack_weights_batch, dim3(num_sms(), nj), 256, 0, s, b, x3);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    SCN_CUDA(cudaGetLastError());
    nj = 0;
    return 0;
  };
  for (int i = 0; i < n; ++i) {
    if (!tags[i] || !W[i]) continue;
    const bool hf = layout_ok(Cin[i], Cout[i]), hb = layout_ok(Cout[i], Cin[i]);
    if ((!hf && !hb) || ((uintptr_t)W[i] & 15)) continue;
    if (x3 != 2 && ((Cin[i] | Cout[i]) & 31)) continue;     // the tiled pack kernel works on 32 x 32 tiles
    PackEntry *e = nullptr;
    SCN_TRY(pack_entry(tags[i], K[i], Cin[i], Cout[i], x3, &e));
    if (e->batch == g_batch_id) continue;                   // the same tensor twice in one graph
    e->has_f = hf; e->has_b = hb;
    e->version = tags[i][1];
    e->stream = s;
    e->batch = g_batch_id;
    e->last_use = ++g_pack_clock;
    b.job[nj++] = PackJob{W[i], e->wf, e->wb, K[i], Cin[i], Cout[i], (hf ? 1 : 0) | (hb ? 2 : 0)};
    if (nj == PACK_JOBS) SCN_TRY(flush());
  }
  return flush();
}

void prepack_weights_end() {
  std::lock_guard<std::mutex> lk(tc::g_pack_mu);
  tc::g_batch_id = 0;
}

namespace tc {

int g_gemm_grid_limit = 0;   // test knob (scn_set_gemm_grid_limit): force many work items per CTA

static bool tf32_shape_ok(const float *X, const float *W, const float *bias, float *Y, int Kd, int N) {
  auto al = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
  return Kd >= KC && Kd % KC == 0 && N >= 16 && N % 16 == 0 && N <= 256 && al(X) && al(W) && al(Y) && (!bias || al(bias));
}

}  // namespace tc

// Y[stationary] = bias + sum_k X[partner_k] @ Wg[k], where Wg[k] is W[k] (transpose_w = 0, W is
// [K][Kd][N]) or W[k]^T (transpose_w = 1, W is [K][N][Kd]).
int osgemm_tc(const float *X, const float *W, const float *bias, float *Y, int Kd, int N, long long n_rows,
              const TileView &tv, int K, int Kb, int precision, int transpose_w, cudaStream_t s, double prof_bytes,
              double prof_flops, const int64_t *weight_tag) {
  using namespace tc;
  if (precision != SCN_PRECISION_TF32 && precision != SCN_PRECISION_FP32_3XTF32 && precision != SCN_PRECISION_BF16) return 1;
  if (!tf32_shape_ok(X, W, bias, Y, Kd, N)) return 1;
  // kernel mode: 0 tf32, 1 3xTF32, 2 bf16
  const int x3 = gemm_mode(precision);
  // 3xTF32 / bf16: a work item covers <= 128 output columns (TMEM also holds converted A stages)
  const int ncb = (x3 && N > 128) ? N / 128 : 1;
  const int NW = N / ncb;
  if (NW * ncb != N || NW % 16 || (x3 && NW > 128)) return 1;     // (converter modes: TMEM also holds the converted stages)
  float *wp = nullptr;
  const int cin = transpose_w ? N : Kd, cout = transpose_w ? Kd : N;
  if (weight_tag) {
    const int r = cached_pack(weight_tag, W, K, cin, cout, transpose_w, x3, s, &wp);
    if (r) return r > 0 ? -1 : r;
  } else {
    const long long total = (long long)K * Kd * N;
    if (workspace_t(&wp, WS_PACKED_W, (size_t)total * (x3 == 1 ? 2 : 1), s)) return -1;
    int pb = cdiv(2 * total, 256);
    if (pb > num_sms() * 8) pb = num_sms() * 8;
    if (x3 == 2)
      SCN_LAUNCH(k_pack_weights_bf16, pb, 256, 0, s, W, reinterpret_cast<__nv_bfloat16 *>(wp), reinterpret_cast<__nv_bfloat16 *>(wp), K,
                                             cin, cout, !transpose_w, transpose_w);
    else
      SCN_LAUNCH(k_pack_weights_both, pb, 256, 0, s, W, wp, wp, K, cin, cout, !transpose_w, transpose_w, x3);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  // ring depths: 2-3 weight-slice stages, then as many gathered-row stages as fit beside the metadata
  // slots (the gathers are latency-bound: depth = bytes in flight)
  const int b_stage = Smem::b_stage_bytes(NW, x3);
  const int meta_slot = Kb * TILE_M * 4 + TILE_M * 4 + 64;
  // 3xTF32 with the low-order halves in TMEM: three weight stages (the weight ring is latency-bound: a TMA
  // bulk copy takes ~1.2 us, so two 32 KB stages paced the step at 0.7 us)
  const int nsb = x3 == 2 ? 4 : ((x3 && LO_TMEM) ? 3 : (b_stage > 16384 ? 2 : 3));
  const int budget = 226 * 1024;
  auto fit = [&](int ms) {
    const int f = (budget - ms * meta_slot - 4 * 4096 - nsb * b_stage - ((x3 == 1 && !LO_TMEM) ? NLO * A_STAGE : 0)) / A_STAGE;
    return f > NSA_MAX ? NSA_MAX : f;
  };
  // Both rings are latency-bound (profiles/experiments/README.md: period = (refill latency + MMA time) / depth),
  // so every gathered-row stage counts: layers whose work items are long (>= 16 steps even with few active
  // offsets) give up the third metadata slot when that buys another stage.
  int ms = MS;
  const int kchunks_ = Kd / KC;
  if (fit(2) > fit(MS) && Kb * kchunks_ >= 32 && tv.n_tiles * ncb * 2 > num_sms()) ms = 2;   // (never with split work items)
  int nsa_fit = fit(ms);
  static const int env_nsa = getenv("SCN_B200_GEMM_NSA") ? atoi(getenv("SCN_B200_GEMM_NSA")) : 0;   // experiments
  if (env_nsa >= 3 && nsa_fit > env_nsa) nsa_fit = env_nsa;
  if (nsa_fit < 3 || nsb > nsa_fit) return 1;   // (the weight ring is never deeper than the gathered-row ring)
  const int depth = nsa_fit - 2;
  const Smem L(NW, Kb, depth + 2, nsb, x3, ms);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ae = cudaSuccess;
    auto setattr = [&](const void *f) {
      if (ae == cudaSuccess) ae = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    };
#define SCN_SETATTR_MODE(M)                                                                                   \
    setattr((const void *)k_osgemm_tf32<1, M>); setattr((const void *)k_osgemm_tf32<2, M>);                     \
    setattr((const void *)k_osgemm_tf32<3, M>); setattr((const void *)k_osgemm_tf32<4, M>);                     \
    setattr((const void *)k_osgemm_tf32<5, M>); setattr((const void *)k_osgemm_tf32<6, M>);
    SCN_SETATTR_MODE(0)
    SCN_SETATTR_MODE(1)
    SCN_SETATTR_MODE(2)
#undef SCN_SETATTR_MODE
    if (ae != cudaSuccess) {
      set_error("cudaFuncSetAttribute(k_osgemm_tf32) failed: %s", cudaGetErrorString(ae));
      return -1;
    }
    attr_set = true;
  }
  uint32_t cols = 32;
  while ((int)cols < NW) cols <<= 1;
  // small grids leave most SMs idle while one CTA walks K * Kd/32 latency-bound steps: split the
  // steps over up to 32 work items per tile (each keeps >= 4 steps even for one active offset)
  int splits = 1;
  const int kchunks = (Kd + KC - 1) / KC;
  if (tv.n_tiles * ncb * 2 <= num_sms() && !tv.identity) {
    splits = num_sms() / (tv.n_tiles * ncb);
    const int max_split = (Kb * kchunks + 3) / 4;
    if (splits > max_split) splits = max_split;
    if (splits > 32) splits = 32;
    if (splits < 1) splits = 1;
  }
  float *ypart = nullptr;
  const long long n_slots = (long long)tv.n_tiles * TILE_M;
  if (splits > 1 && workspace_t(&ypart, WS_SPLITK, (size_t)splits * n_slots * N, s)) return -1;
  const int n_items = tv.n_tiles * splits * ncb;
  int grid = n_items < num_sms() ? n_items : num_sms();            // persistent: one CTA per SM
  static const int env_limit = getenv("SCN_B200_GEMM_GRID") ? atoi(getenv("SCN_B200_GEMM_GRID")) : 0;   // experiments
  if (env_limit > 0 && grid > env_limit) grid = env_limit;
  if (g_gemm_grid_limit > 0 && grid > g_gemm_grid_limit) grid = g_gemm_grid_limit;
  int *sched = nullptr;
  if (sched_counters(&sched, s)) return -1;
  static const int static_first = getenv("SCN_B200_GEMM_STATIC_FIRST") ? atoi(getenv("SCN_B200_GEMM_STATIC_FIRST")) : 1;
  prof_begin(PROF_GEMM, s);
#define SCN_OSGEMM_LAUNCH(D, T3)                                                                              \
  SCN_LAUNCH_GEMM((k_osgemm_tf32<D, T3>), grid, gemm_threads(T3), L.total, s, X, wp, bias, Y, Kd, NW, N, Kb, n_rows, tv, cols, ypart, \
                                                                n_items, splits, nsb, sched, ms, static_first)
#define SCN_OSGEMM_DEPTH(T3)                                                                                  \
  switch (depth) {                                                                                            \
    case 6: SCN_OSGEMM_LAUNCH(6, T3); break;                                                                  \
    case 5: SCN_OSGEMM_LAUNCH(5, T3); break;                                                                  \
    case 4: SCN_OSGEMM_LAUNCH(4, T3); break;                                                                  \
    case 3: SCN_OSGEMM_LAUNCH(3, T3); break;                                                                  \
    case 2: SCN_OSGEMM_LAUNCH(2, T3); break;                                                                  \
    default: SCN_OSGEMM_LAUNCH(1, T3); break;                                                                 \
  }
  if (x3 == 2) { SCN_OSGEMM_DEPTH(2) } else if (x3 == 1) { SCN_OSGEMM_DEPTH(1) } else { SCN_OSGEMM_DEPTH(0) }
#undef SCN_OSGEMM_DEPTH
#undef SCN_OSGEMM_LAUNCH
  prof_end(PROF_GEMM, s, prof_bytes, prof_flops);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (splits > 1 && e == cudaSuccess) {
    const long long work = n_slots * (N >> 2);
    SCN_LAUNCH(k_splitk_reduce, cdiv(work, 256), 256, 0, s, ypart, bias, Y, N, splits, n_slots, n_rows, tv);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    set_error("k_osgemm_tf32 launch failed: %s", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// weight gradient on the tensor cores: partial[w] = X[in rows of work item w]^T @ dY[out rows]
// M = Cin (zero-padded to 128; two accumulators when Cin = 256), N = Cout, reduction over the pairs.
// Both operands are gathered rows, i.e. MN-major.  For 32-bit operands tcgen05 accepts MN-major only
// in the SWIZZLE_128B_BASE32B layout (measured with tools/umma_probe.py: every other layout type
// yields zeros): atoms of 4 pairs x 128 B (32 channels), the 32-byte chunks of a row XOR-ed with the
// pair index inside the atom (cute Layout_MN_SW128_32B_Atom, Swizzle<2,5,2>).  Atoms along the
// channels are 512 B apart (LBO), atoms along the pairs SBO = 512 * channel-atoms apart.
// ---------------------------------------------------------------------------------------------
namespace tc {

constexpr int DW_NPS = 8;          // pair-list slots
constexpr int DW_PW = 8;                       // producer (gather) warps; warps 0-3 also run the epilogue
constexpr int DW_MMA_W = DW_PW, DW_LOAD_W = DW_PW + 1, DW_CONV_W = DW_PW + 2;
constexpr int NT_DW = (DW_PW + 2) * 32;        // + MMA warp + pair-list loader warp
constexpr int DW_CW = 8;                       // converter warps (3xTF32 mode)
constexpr int NT_DW3 = (DW_PW + 2 + DW_CW) * 32;

__device__ __forceinline__ uint64_t make_desc_b32(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return make_desc(saddr, lbo, sbo) | (1ull << 61);   // layout_type 1 = SWIZZLE_128B_BASE32B
}

struct DwSmem {
  // 3xTF32: stages [ns, ns + NLO) of both operand rings hold the low-order halves
  int a, b, pairs, bars, tmem_slot, total, a_stage, b_stage;
  __host__ __device__ DwSmem(int MA, int NA, int KP, int ns, int x3) {
    a_stage = KP * MA * 128;
    b_stage = KP * NA * 128;
    a = 0;
    b = a + (ns + (x3 ? NLO : 0)) * a_stage;
    pairs = b + (ns + (x3 ? NLO : 0)) * b_stage;
    bars = pairs + DW_NPS * KP * 8;
    tmem_slot = bars + (2 * ns + 2 * DW_NPS + 1 + 2 * NLO) * 8;
    total = tmem_slot + 16;
  }
};

// Warp-specialised like the gather-GEMM: warp 5 streams the work item's (in,out) pair list into a
// ring of shared-memory slots, warps 0-3 gather both operands' rows with cp.async (arrivals lag
// NSTAGE/2 steps behind the issue, so nobody waits for fresh data), warp 4 issues the MMAs.
#ifdef SCN_EXPERIMENT_STALLS
__device__ long long g_stall_dw[20][8];
#define SCN_STALL_WRITE_DW                                                               \
  do {                                                                                   \
    stall_[7] = clock64() - stall_t0_;                                                   \
    if (blockIdx.x == 0 && lane == 0)                                                    \
      for (int i_ = 0; i_ < 8; ++i_) g_stall_dw[warp][i_] = stall_[i_];                  \
  } while (0)
#else
#define SCN_STALL_WRITE_DW do { } while (0)
#endif
template <int NSTAGE, bool X3>
__global__ void __launch_bounds__(X3 ? NT_DW3 : NT_DW)
k_dw_tf32(const float *__restrict__ X, const float *__restrict__ dY, const int32_t *__restrict__ pairs,
          const DwWork *__restrict__ work, float *__restrict__ partial, int Cin, int Cout, int xcol, int ycol,
          long long ident_n, int ident_chunk, int KP, uint32_t tmem_cols) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int CA = Cin >> 5, CB = Cout >> 5;        // real 32-channel atoms
  const int MA = Cin > 128 ? CA : 4;              // atoms per k-atom of A (M padded to 128)
  const int halves = Cin > 128 ? 2 : 1;
  const DwSmem L(MA, CB, KP, NSTAGE, X3 ? 1 : 0);
  const uint32_t a_base = smem_u32(smem + L.a), b_base = smem_u32(smem + L.b);
  const uint32_t bar_full = smem_u32(smem + L.bars);
  const uint32_t bar_empty = bar_full + NSTAGE * 8;
  const uint32_t bar_pfull = bar_empty + NSTAGE * 8;
  const uint32_t bar_pempty = bar_pfull + DW_NPS * 8;
  const uint32_t bar_done = bar_pempty + DW_NPS * 8;
  const uint32_t bar_fullL = bar_done + 8;
  const uint32_t bar_emptyL = bar_fullL + NLO * 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.tmem_slot);
  int2 *sPairs = reinterpret_cast<int2 *>(smem + L.pairs);      // [DW_NPS][KP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  SCN_STALL_DECL;
  const int KA = KP >> 2;                          // k-atoms (4 pairs) per step
  const uint32_t sbo_a = MA * 512, sbo_b = CB * 512;

  // zero the atoms that pad M up to 128 (never written by the gathers)
  if (CA < MA) {
    const int per_ka = (MA - CA) * 32;             // 16-byte units per k-atom
    const int total16 = (NSTAGE + (X3 ? NLO : 0)) * KA * per_ka;
    for (int i = tid; i < total16; i += (X3 ? NT_DW3 : NT_DW)) {
      const int ka = i / per_ka, r = i - ka * per_ka;   // ka counts over all stages
      *reinterpret_cast<float4 *>(smem + L.a + (ka / KA) * L.a_stage + (ka % KA) * sbo_a + CA * 512 + r * 16) =
          make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(bar_full + i * 8, DW_PW * 32);
      mbar_init(bar_empty + i * 8, 1);
    }
    for (int i = 0; i < DW_NPS; ++i) {
      mbar_init(bar_pfull + i * 8, 1);
      mbar_init(bar_pempty + i * 8, DW_PW);
    }
    mbar_init(bar_done, 1);
    if (X3)
      for (int i = 0; i < NLO; ++i) {
        mbar_init(bar_fullL + i * 8, DW_CW);     // one arrival per converter warp
        mbar_init(bar_emptyL + i * 8, 1);
      }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  // shared-memory set-up and the TMEM allocation overlap the previous kernel's tail; nothing global before this point
  pdl_wait();
  long long start;
  int len, out_slot = blockIdx.x;
  if (work) {
    const DwWork w = work[blockIdx.x];
    start = w.start;
    len = w.len;
    out_slot = w.slot;                               // launch order != partial order (ensure_dw_work)
  } else {
    start = (long long)blockIdx.x * ident_chunk;
    len = (int)min((long long)ident_chunk, ident_n - start);
  }
  const int steps = (len + KP - 1) / KP;

  if (warp == DW_LOAD_W) {
    // ===== pair-list loader: lane l carries pairs l and l+32 of a step, 4 steps ahead in registers =====
    auto load_pair = [&](int st, int i) -> int2 {
      const int p = st * KP + i;
      int2 v = make_int2(-1, -1);
      if (i < KP && st < steps && p < len) {
        if (pairs) v = __ldg(reinterpret_cast<const int2 *>(pairs) + start + p);
        else v = make_int2((int)(start + p), (int)(start + p));
      }
      return v;
    };
    int2 ra[4], rb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { ra[u] = load_pair(u, lane); rb[u] = load_pair(u, lane + 32); }
    for (int st0 = 0; st0 < steps; st0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int st = st0 + u;
        if (st < steps) {
          const int slot = st % DW_NPS, use = st / DW_NPS;
          if (use > 0) MBW(bar_pempty + slot * 8, (use - 1) & 1, 0);
          if (lane == 0) SCN_TRACE_DW(0, 1);
          if (lane < KP) sPairs[slot * KP + lane] = ra[u];
          if (KP > 32) sPairs[slot * KP + lane + 32] = rb[u];
          ra[u] = load_pair(st + 4, lane);
          rb[u] = load_pair(st + 4, lane + 32);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_pfull + slot * 8);
        }
      }
    }
  } else if (warp < DW_PW) {
    // ===== operand producers =====
    const int j8 = tid & 7, p4 = (tid >> 3) & 3;
    // byte offset of this thread's 16-byte piece inside an atom: row p4, 32B chunk (j8/2)^p4, half j8&1
    const uint32_t piece = p4 * 128 + ((((uint32_t)j8 >> 1) ^ (uint32_t)p4) << 5) + (j8 & 1) * 16;
    const int ca_sh = __ffs(CA) - 1, cb_sh = __ffs(CB) - 1;
    const float *Xl = X + j8 * 4, *Yl = dY + j8 * 4;
    for (int st = 0; st < steps; ++st) {
      const int stage = st % NSTAGE, use = st / NSTAGE;
      const int slot = st % DW_NPS;
      MBW(bar_pfull + slot * 8, (st / DW_NPS) & 1, 0);
      if (tid == 0) SCN_TRACE_DW(1, 1);
      if (use > 0) MBW(bar_empty + stage * 8, (use - 1) & 1, 1);
      if (tid == 0) SCN_TRACE_DW(1, 2);
      const int2 *sp = sPairs + slot * KP;
      const uint32_t sa = a_base + stage * L.a_stage + piece, sb = b_base + stage * L.b_stage + piece;
      // CA, CB are powers of two (host check): shifts instead of divisions, and a 4-way unroll so that the
      // pair load -> address -> cp.async chains of consecutive pieces overlap (one chain at a time made the
      // ISSUE of a step's 16 copies take 1.2 us - the kernel's bottleneck, tools/dw_trace.py)
#pragma unroll 4
      for (int c = warp; c < (KA << ca_sh); c += DW_PW) {
        const int ka = c >> ca_sh, mi = c & (CA - 1);
        const int2 pr = sp[ka * 4 + p4];
        const int xi = xcol ? pr.y : pr.x;
        cp_async_16(sa + ka * sbo_a + mi * 512, Xl + (size_t)((unsigned)max(xi, 0) * (unsigned)Cin + mi * 32), xi < 0 ? 0 : 16);
      }
#pragma unroll 4
      for (int c = warp; c < (KA << cb_sh); c += DW_PW) {
        const int ka = c >> cb_sh, ni = c & (CB - 1);
        const int2 pr = sp[ka * 4 + p4];
        const int yi = ycol ? pr.y : pr.x;
        cp_async_16(sb + ka * sbo_b + ni * 512, Yl + (size_t)((unsigned)max(yi, 0) * (unsigned)Cout + ni * 32), yi < 0 ? 0 : 16);
      }
      cp_async_mbar_arrive(bar_full + stage * 8);   // completes when this thread's copies have landed
      mbar_arrive(bar_full + stage * 8);
      if (tid == 0) SCN_TRACE_DW(1, 3);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pempty + slot * 8);
    }
  } else if (warp == DW_MMA_W) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues (see the gather-GEMM's MMA role) =====
    {
      const uint32_t idesc = make_idesc(128, Cout, 1, 1);
      int stage = 0, stl = 0;
      uint32_t ph = 0, phl = 0;
      for (int st = 0; st < steps; ++st) {
        // this warp paces the kernel: probe with test_wait first; in 3xTF32 mode the converters' barrier
        // implies the landed stage
        if (X3) {
          if (!mbar_test(bar_fullL + stl * 8, phl)) MBW(bar_fullL + stl * 8, phl, 4);
        } else {
          if (!mbar_test(bar_full + stage * 8, ph)) MBW(bar_full + stage * 8, ph, 2);
        }
        if (lane == 0) SCN_TRACE_DW(2, 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = a_base + stage * L.a_stage, sb = b_base + stage * L.b_stage;
          const uint32_t la = a_base + (NSTAGE + stl) * L.a_stage, lb = b_base + (NSTAGE + stl) * L.b_stage;
          // descriptors differ only in the start address (bits 0-13, 16-byte units; the stages lie below
          // 256 KB, so adding to the low word never carries into the next field)
          uint64_t bd = make_desc_b32(sb, 512, sbo_b), bl = make_desc_b32(lb, 512, sbo_b);
          uint64_t ad0 = make_desc_b32(sa, 512, sbo_a), al0 = make_desc_b32(la, 512, sbo_a);
          const uint64_t da = (2 * sbo_a) >> 4, db = (2 * sbo_b) >> 4, dh = (4 * 512) >> 4;
#pragma unroll 2
          for (int kb = 0; kb < (KP >> 3); ++kb) {      // one MMA consumes 8 pairs = 2 k-atoms
            for (int h = 0; h < halves; ++h) {
              const uint64_t ad = ad0 + h * dh;
              mma_tf32(tmem_d + (uint32_t)(h * Cout), ad, bd, idesc, (st > 0 || kb > 0) ? 1u : 0u);
              if (X3) {   // + lo(x) * dy + x * lo(dy)
                mma_tf32(tmem_d + (uint32_t)(h * Cout), al0 + h * dh, bd, idesc, 1u);
                mma_tf32(tmem_d + (uint32_t)(h * Cout), ad, bl, idesc, 1u);
              }
            }
            ad0 += da; al0 += da; bd += db; bl += db;
          }
          tc_commit(bar_empty + stage * 8);            // one commit per step (the converters wait on it too)
          if (st + 1 == steps) tc_commit(bar_done);
        }
        __syncwarp();
        if (lane == 0) SCN_TRACE_DW(2, 2);
        if (++stage == NSTAGE) { stage = 0; ph ^= 1; }
        if (++stl == NLO) { stl = 0; phl ^= 1; }
      }
    }
  }
  if (X3 && warp >= DW_CONV_W) {
    // ===== converters (3xTF32): low-order halves of both operand stages, same positions =====
    const int ct = tid - DW_CONV_W * 32;
    const int a16 = L.a_stage >> 4, b16 = L.b_stage >> 4;
    for (int st = 0; st < steps; ++st) {
      const int stage = st % NSTAGE, stl = st % NLO;
      MBW(bar_full + stage * 8, (st / NSTAGE) & 1, 2);
      if (st >= NLO) MBW(bar_empty + ((st - NLO) % NSTAGE) * 8, ((st - NLO) / NSTAGE) & 1, 1);
      auto lo4 = [](const float4 v) {
        float4 o;
        o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        return o;
      };
      const float4 *srca = reinterpret_cast<const float4 *>(smem + L.a + stage * L.a_stage);
      float4 *dsta = reinterpret_cast<float4 *>(smem + L.a + (NSTAGE + stl) * L.a_stage);
      if (CA == MA) {
        for (int i = ct; i < a16; i += DW_CW * 32) dsta[i] = lo4(srca[i]);
      } else {
        // M is padded to 128 (Cin = 32 / 64): only the CA real 32-channel atoms of each k-atom carry data; the
        // padding atoms of the low-order stages were zeroed at kernel start and stay zero (converting them was
        // 50 - 75 % of this role's shared-memory traffic)
        const int sh = __ffs(CA) - 1 + 5;                 // CA is a power of two (host check)
        const int real16 = KA << sh;                      // 16-byte units: KA k-atoms x CA atoms x 512 B
        for (int i = ct; i < real16; i += DW_CW * 32) {
          const int ka = i >> sh, r = i & ((1 << sh) - 1);
          const int o = ka * (MA * 32) + r;
          dsta[o] = lo4(srca[o]);
        }
      }
      const float4 *srcb = reinterpret_cast<const float4 *>(smem + L.b + stage * L.b_stage);
      float4 *dstb = reinterpret_cast<float4 *>(smem + L.b + (NSTAGE + stl) * L.b_stage);
      for (int i = ct; i < b16; i += DW_CW * 32) dstb[i] = lo4(srcb[i]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_fullL + stl * 8);
    }
  }
  // ===== epilogue (warps 0-3): accumulator row = input channel, columns = output channels =====
  if (warp < 4) {
    if (steps > 0) {
      SCN_STALL_WRITE_DW;
      MBW(bar_done, 0, 5);
      tc_fence_after();
    }
    float *out = partial + (long long)out_slot * Cin * Cout;
    for (int h = 0; h < halves; ++h) {
      const int ci = h * 128 + warp * 32 + lane;
      for (int c0 = 0; c0 < Cout; c0 += 32) {
        uint32_t v[32];
        if (steps > 0) {
          tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * Cout + c0), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (ci < Cin) {
          float *o = out + (long long)ci * Cout + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4 *>(o + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                             __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
      }
    }
  }
  if (warp >= 4) SCN_STALL_WRITE_DW;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

}  // namespace tc

// partial[w] (w < n_work) = X[rows]^T @ dY[rows] over the pairs of work item w.  >0: shape not handled.
}  // namespace scn
#if defined(SCN_EXPERIMENT_TRACE) || defined(SCN_EXPERIMENT_TRACE_DW)
extern "C" int scn_debug_trace_read(unsigned long long *out, int *counts, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, scn::tc::g_trace, sizeof(unsigned long long) * 5 * 8192);
  cudaMemcpyFromSymbol(counts, scn::tc::g_trace_n, sizeof(int) * 5);
  if (reset) {
    int z[5] = {0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(scn::tc::g_trace_n, z, sizeof(z));
  }
  return 0;
}
#endif
#ifdef SCN_EXPERIMENT_STALLS
extern "C" int scn_debug_stalls_read(long long *out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, scn::tc::g_stall, sizeof(long long) * 20 * 8);
  return 0;
}
extern "C" int scn_debug_stalls_dw_read(long long *out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, scn::tc::g_stall_dw, sizeof(long long) * 20 * 8);
  return 0;
}
extern "C" int scn_debug_cta_read(long long *out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, scn::tc::g_cta, sizeof(long long) * 160 * 4);
  return 0;
}
#endif
extern "C" int scn_set_gemm_grid_limit(int max_ctas) {
  scn::tc::g_gemm_grid_limit = max_ctas;
  return 0;
}
namespace scn {

int dw_partial_tc(const float *X, const float *dY, const int32_t *pairs, const DwWork *work, float *partial,
                  int Cin, int Cout, int xcol, int ycol, int n_work, long long ident_n, int ident_chunk,
                  int precision, cudaStream_t s) {
  using namespace tc;
  auto al = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
  // bf16 mode: the weight gradient runs on the single-pass tf32 kernel (both operands are gathered fp32 rows)
  if (precision != SCN_PRECISION_TF32 && precision != SCN_PRECISION_FP32_3XTF32 && precision != SCN_PRECISION_BF16) return 1;
  const bool x3 = precision == SCN_PRECISION_FP32_3XTF32;
  if (Cin < 32 || Cin % 32 || (Cin > 128 && Cin != 256) || Cout < 32 || Cout % 32 || Cout > 256) return 1;
  if (((Cin >> 5) & ((Cin >> 5) - 1)) || ((Cout >> 5) & ((Cout >> 5) - 1))) return 1;   // 32-channel atoms: power of two
  if (!al(X) || !al(dY) || !al(partial)) return 1;
  const int MA = Cin > 128 ? Cin >> 5 : 4, NA = Cout >> 5;
  int KP = (MA + NA) <= 5 ? 64 : 32;
  if (x3 && KP * (MA + NA) * 128 > 40 * 1024) KP = 16;     // 3xTF32 keeps raw + low-order stages: smaller steps
  const int stage_bytes = KP * (MA + NA) * 128;
  int ns = (205 * 1024 - (x3 ? NLO * stage_bytes : 0)) / stage_bytes;
  ns = ns >= 6 ? 6 : (ns >= 4 ? 4 : 3);
  const DwSmem L(MA, NA, KP, ns, x3 ? 1 : 0);
  if (L.total > 227 * 1024) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ae = cudaSuccess;
    auto setattr = [&](const void *f) {
      if (ae == cudaSuccess) ae = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    };
    setattr((const void *)k_dw_tf32<3, false>); setattr((const void *)k_dw_tf32<4, false>);
    setattr((const void *)k_dw_tf32<6, false>); setattr((const void *)k_dw_tf32<3, true>);
    setattr((const void *)k_dw_tf32<4, true>);  setattr((const void *)k_dw_tf32<6, true>);
    if (ae != cudaSuccess) {
      set_error("cudaFuncSetAttribute(k_dw_tf32) failed: %s", cudaGetErrorString(ae));
      return -1;
    }
    attr_set = true;
  }
  uint32_t cols = 32;
  while ((int)cols < Cout * (Cin > 128 ? 2 : 1)) cols <<= 1;
#define SCN_DW_LAUNCH(NS, T3)                                                                                       \
  SCN_LAUNCH_GEMM((k_dw_tf32<NS, T3>), n_work, T3 ? NT_DW3 : NT_DW, L.total, s, X, dY, pairs, work, partial, Cin, Cout, xcol, ycol, \
                                                                 ident_n, ident_chunk, KP, cols)
  if (x3) {
    if (ns == 3) SCN_DW_LAUNCH(3, true);
    else if (ns == 4) SCN_DW_LAUNCH(4, true);
    else SCN_DW_LAUNCH(6, true);
  } else {
    if (ns == 3) SCN_DW_LAUNCH(3, false);
    else if (ns == 4) SCN_DW_LAUNCH(4, false);
    else SCN_DW_LAUNCH(6, false);
  }
#undef SCN_DW_LAUNCH
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("k_dw_tf32 launch failed: %s", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

}  // namespace scn
