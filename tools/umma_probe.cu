// Developer probe (not product code): run ONE tcgen05.mma.kind::tf32 (or several along K) on
// host-prepared shared-memory images to pin down operand layouts / descriptor fields on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libumma_probe.so umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const uint8_t *a_img, const uint8_t *b_img, int a_bytes, int b_bytes, uint64_t a_desc_hi,
                        uint64_t b_desc_hi, uint32_t idesc, int n_mma, int a_step, int b_step, int N, float *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t *sa = smem, *sb = smem + ((a_bytes + 1023) / 1024) * 1024;
  for (int i = threadIdx.x; i < a_bytes; i += blockDim.x) sa[i] = a_img[i];
  for (int i = threadIdx.x; i < b_bytes; i += blockDim.x) sb[i] = b_img[i];
  const uint32_t barp = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(barp) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_mma; ++i) {
      const uint64_t ad = a_desc_hi | (uint64_t)(((smem_u32(sa) + i * a_step) & 0x3FFFF) >> 4);
      const uint64_t bd = b_desc_hi | (uint64_t)(((smem_u32(sb) + i * b_step) & 0x3FFFF) >> 4);
      const uint32_t acc = i > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(barp) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(barp) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 128 threads: warp w reads lanes 32w..
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256) : "memory");
}

extern "C" int umma_probe(const uint8_t *a_img, const uint8_t *b_img, int a_bytes, int b_bytes, uint64_t a_desc_hi,
                          uint64_t b_desc_hi, uint32_t idesc, int n_mma, int a_step, int b_step, int N, float *out) {
  const int smem = ((a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_probe<<<1, 128, smem>>>(a_img, b_img, a_bytes, b_bytes, a_desc_hi, b_desc_hi, idesc, n_mma, a_step, b_step, N, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
