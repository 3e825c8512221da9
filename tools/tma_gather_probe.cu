// Developer probe (not product code): semantics of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a -
// tensor-map box shape, where the 4 rows land in shared memory under SWIZZLE_128B, and what happens for
// out-of-range row indices (the gather-GEMM wants zero-fill for missing partners).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_gather_probe tma_gather_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k_probe(const __grid_constant__ CUtensorMap tm, int col, int r0, int r1, int r2, int r3, uint32_t bytes,
                        float *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t barp = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = -7.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(barp) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(barp), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(dst),
        "l"(&tm), "r"(barp), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(barp) : "memory");
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<float *>(smem)[i];
}

int main() {
  const int R = 1000, C = 128;
  std::vector<float> h((size_t)R * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = r * 1000.f + c;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, 4096);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &q);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  for (int boxrows = 1; boxrows <= 4; boxrows *= 4) {
    for (int sw = 0; sw < 2; ++sw) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}, strides[1] = {(cuuint64_t)C * 4};
      cuuint32_t box[2] = {32, (cuuint32_t)boxrows}, es[2] = {1, 1};
      CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("== box rows %d swizzle %d encode rc=%d\n", boxrows, sw, (int)rc);
      if (rc) continue;
      cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
      k_probe<<<1, 128, 8192>>>(tm, 32, 5, -1, 999, 1000, 512, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
      float o[1024];
      cudaMemcpy(o, out, 4096, cudaMemcpyDeviceToHost);
      for (int row = 0; row < 4; ++row) {
        printf("smem row %d (128B):", row);
        for (int c = 0; c < 32; c += 4) printf(" %.0f", o[row * 32 + c]);
        printf("\n");
      }
      printf("beyond 512B: %.0f %.0f\n", o[128], o[160]);
    }
  }
  return 0;
}
