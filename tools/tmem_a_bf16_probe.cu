// Developer probe (GPU box): tcgen05.mma.kind::f16 (bf16 operands) with the A operand in TENSOR MEMORY and a
// no-swizzle K-major bf16 B operand in shared memory (derived from tmem_a_probe.cu).
// Hypothesis: row i of A (M = 128) lives in TMEM lane i, its K = 8 tf32 values of one instruction in 8
// consecutive 32-bit columns; a [128 x 32] slice written with tcgen05.st 32x32b.x32 (thread = row) is
// consumed by 4 instructions at column offsets 0, 8, 16, 24.  B: K-major SWIZZLE_128B in shared memory.
// Prints max |D - A B^T|.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(16 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
constexpr int N = 64, KC = 32;

__global__ void k_probe(const float *A, const float *B, float *D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // B: row n at (n/8)*1024 + (n%8)*128, 16-byte chunk j at j ^ (n%8)
  for (int i = tid; i < N * KC; i += 128) {
    const int n = i / KC, k = i % KC;
    const int off = (k >> 3) * N * 16 + n * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16 *>(smem + off) = __float2bfloat16(B[n * KC + k]);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tslot;
  const uint32_t t_acc = tb, t_a = tb + 64;          // columns 0-63 accumulator, 64-95 A slice
  // thread = row: 32 values -> TMEM lane tid, columns 64..95
  uint32_t v[32];
  for (int k = 0; k < 16; ++k) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(A[tid * KC + 2 * k], A[tid * KC + 2 * k + 1]);   // .x (low half) = even k
    v[k] = *reinterpret_cast<const uint32_t *>(&p);
  }
  for (int k = 16; k < 32; ++k) v[k] = 0;
  const uint32_t ta = t_a + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(ta),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int kk = 0; kk < 2; ++kk) {
      // no swizzle, K-major: LBO = distance between core matrices along K = N*16, SBO = 8 rows * 16 B = 128
      const uint32_t sb = smem_u32(smem) + kk * 2 * N * 16;
      const uint64_t bd = (uint64_t)((sb & 0x3FFFFu) >> 4) | ((uint64_t)((N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      const uint32_t acc = kk > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(t_acc),
          "r"(t_a + (uint32_t)kk * 8), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t w[8];
    const uint32_t tad = t_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "r"(tad));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int i = 0; i < 8; ++i) D[tid * N + c0 + i] = __uint_as_float(w[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" ::"r"(tb) : "memory");
}

int main() {
  const int M = 128;
  float *hA = (float *)malloc(M * KC * 4), *hB = (float *)malloc(N * KC * 4), *hD = (float *)malloc(M * N * 4);
  srand(1);
  for (int i = 0; i < M * KC; ++i) hA[i] = (float)((rand() % 255) - 127) / 16.f;     // tf32-exact
  for (int i = 0; i < N * KC; ++i) hB[i] = (float)((rand() % 255) - 127) / 32.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, M * KC * 4); cudaMalloc(&dB, N * KC * 4); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * KC * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * KC * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, M * N * 4);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k_probe<<<1, 128, 16 * 1024>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
  double worst = 0, ref_max = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < KC; ++k) r += (double)hA[m * KC + k] * hB[n * KC + k];
      worst = fmax(worst, fabs(r - hD[m * N + n]));
      ref_max = fmax(ref_max, fabs(r));
    }
  printf("A-in-TMEM bf16 MMA (B no-swizzle K-major): max |D - A B^T| = %g (max |ref| %g) -> %s\n", worst, ref_max, worst < 1e-3 ? "LAYOUT CONFIRMED" : "MISMATCH");
  return 0;
}
