// Developer probe (GPU box): throughput of 16-byte cp.async (LDGSTS) row gathers per SM as a function of how the
// 32 lanes of one instruction are laid over rows - 8 lanes per row (4 rows x 128 B, the gather-GEMM / weight-gradient
// producers), 16 per row (2 x 256 B), 32 per row (1 x 512 B) - and of plain LDG.128 + STS.  Rows are picked at random
// from tables of 64 MB (L2-resident), 128 MB, 256 MB and 1 GB (DRAM) filled with incompressible data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ldgsts_probe tools/ldgsts_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void k_fill(uint32_t *p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = hash((uint32_t)i) | 0x3f000000u;
}

// LPR = lanes per row (8, 16, 32); every warp instruction copies 32/LPR rows x LPR*16 bytes
template <int LPR, bool LDG>
__global__ void __launch_bounds__(256) k_gather(const float4 *__restrict__ tab, uint32_t n_rows, int iters, long long *cycles,
                                                float *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];     // 128 KB staging: 8 slots of 16 KB
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPR, piece = lane % LPR;       // row within the instruction, 16-byte piece within the row
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // one "stage" = 16 KB = 32 warp instructions over the CTA: 4 per warp
    const uint32_t slot = base + (it & 7) * 16384;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t op = (uint32_t)((it * 4 + u) * 8 + warp);
      const uint32_t row = hash(op * (32 / LPR) + sub + blockIdx.x * 0x9e3779b9u) % n_rows;
      const float4 *src = tab + (size_t)row * LPR + piece;         // rows of LPR x 16 bytes: table size = working set
      const uint32_t dst = slot + (u * 8 + warp) * 512 + lane * 16;
      if (LDG) {
        const float4 v = __ldg(src);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      } else {
        cp16(dst, src);
      }
    }
    if (!LDG) {
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      asm volatile("cp.async.wait_group 6;\n" ::: "memory");       // up to 7 stages (112 KB) in flight
    }
  }
  if (!LDG) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  acc += reinterpret_cast<float *>(smem)[tid];
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

template <int LPR, bool LDG>
static void run(const char *name, const float4 *tab, size_t table_bytes, long long *dcyc, float *sink) {
  const uint32_t n_rows = (uint32_t)(table_bytes / (LPR * 16));
  const int iters = 4000, grid = 148;
  cudaFuncSetAttribute(k_gather<LPR, LDG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  k_gather<LPR, LDG><<<grid, 256, 131072>>>(tab, n_rows, 200, dcyc, sink);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k_gather<LPR, LDG><<<grid, 256, 131072>>>(tab, n_rows, iters, dcyc, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long h[148];
  cudaMemcpy(h, dcyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double)h[i];
  mean /= grid;
  const double bytes = (double)iters * 16384.0;
  printf("%-44s table %4zu MB: %6.1f B/clk/SM  (%.0f cycles per 16 KB stage), chip %.2f TB/s, err %s\n", name, table_bytes >> 20, bytes / mean,
         mean / iters, bytes * grid / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const size_t big = (size_t)1 << 30;
  float4 *tab;
  long long *dcyc;
  float *sink;
  cudaMalloc(&tab, big);
  k_fill<<<148 * 8, 256>>>(reinterpret_cast<uint32_t *>(tab), big / 4);      // incompressible contents
  cudaMalloc(&dcyc, 148 * 8);
  cudaMalloc(&sink, 4);
  for (size_t n_rows : {(size_t)64 << 20 /* L2-resident */, (size_t)128 << 20, (size_t)256 << 20, (size_t)1 << 30 /* DRAM */}) {
    run<8, false>("cp.async 8 lanes/row (4 rows x 128 B per op)", tab, n_rows, dcyc, sink);
    run<16, false>("cp.async 16 lanes/row (2 rows x 256 B per op)", tab, n_rows, dcyc, sink);
    run<32, false>("cp.async 32 lanes/row (1 row x 512 B per op)", tab, n_rows, dcyc, sink);
    run<8, true>("LDG.128 + STS 8 lanes/row", tab, n_rows, dcyc, sink);
    run<32, true>("LDG.128 + STS 32 lanes/row", tab, n_rows, dcyc, sink);
  }
  return 0;
}
