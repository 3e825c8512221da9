// pdl_probe.cu - developer probe: what does programmatic dependent launch (griddepcontrol) buy on a chain of
// small dependent kernels?  Chains of N kernels (each: wait for the predecessor, then a short dependent
// read-modify-write of one buffer) are timed with CUDA events, launched (a) plainly, (b) with
// cudaLaunchAttributeProgrammaticStreamSerialization, on the legacy default stream and on a non-blocking stream.
// The result is checked (every kernel adds 1 to every element), so a broken dependency shows up as a wrong sum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pdl_probe tools/pdl_probe.cu && tools/pdl_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void k_step(float *buf, int n, int spin) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = buf[i];
    for (int s = 0; s < spin; ++s) v = v * 1.0000001f + 1e-9f;
    buf[(i + 977) % n] = v + 1.f;      // write a DIFFERENT element: a kernel that starts early reads stale data
  }
}

static double run(cudaStream_t s, int pdl, int n_kernels, int blocks, int spin, float *buf, int n) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaMemsetAsync(buf, 0, n * sizeof(float), s);
  cudaStreamSynchronize(s);
  cudaEventRecord(a, s);
  for (int i = 0; i < n_kernels; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_step, buf, n, spin);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  cudaEventRecord(b, s);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  float *h = (float *)malloc(n * sizeof(float));
  cudaMemcpy(h, buf, n * sizeof(float), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (int i = 0; i < n; ++i) sum += h[i];
  free(h);
  const double expect = (double)n * n_kernels;
  printf("  %s  %5d kernels x %4d blocks spin %4d: %8.3f ms = %6.2f us/kernel   sum/expected %.6f\n",
         pdl ? "PDL  " : "plain", n_kernels, blocks, spin, ms, 1e3 * ms / n_kernels, sum / expect);
  return ms;
}

int main() {
  cudaStream_t nb;
  cudaStreamCreateWithFlags(&nb, cudaStreamNonBlocking);
  for (int which = 0; which < 2; ++which) {
    cudaStream_t s = which ? nb : 0;
    printf("%s stream\n", which ? "non-blocking" : "legacy default");
    const int cfgs[4][2] = {{1, 0}, {148, 0}, {148 * 8, 200}, {148 * 8, 2000}};
    for (auto &c : cfgs) {
      const int blocks = c[0], n = blocks * 256;
      float *buf;
      cudaMalloc(&buf, n * sizeof(float));
      run(s, 0, 200, blocks, c[1], buf, n);      // warm-up
      run(s, 0, 2000, blocks, c[1], buf, n);
      run(s, 1, 2000, blocks, c[1], buf, n);
      cudaFree(buf);
    }
  }
  return 0;
}
