// Micro-benchmark: per-launch cost of a chain of dependent tiny kernels in a CUDA graph,
// with and without programmatic dependent launch (PDL).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void step_plain(float* x, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x[i] * 1.0001f + 1.f;
}
__global__ void step_pdl(float* x, int n) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x[i] * 1.0001f + 1.f;
}
template <typename K>
float run(K kern, bool pdl, float* x, int n, int chain, int grid) {
  cudaStream_t st; cudaStreamCreate(&st);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < chain; ++i) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, x, n);
  }
  cudaStreamEndCapture(st, &g);
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return -1; }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(a, st);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(b, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms * 1000.f / (10 * chain);
}
int main() {
  int n = 16 * 512; float* x; cudaMalloc(&x, (1 << 22) * sizeof(float)); cudaMemset(x, 0, (1 << 22) * 4);
  for (int grid : {32, 148, 1024}) {
    int nn = grid * 256;
    printf("grid %4d: plain %.2f us/launch | pdl %.2f us/launch | plain kernel w/ pdl attr %.2f\n", grid,
           run(step_plain, false, x, nn, 200, grid), run(step_pdl, true, x, nn, 200, grid),
           run(step_plain, true, x, nn, 200, grid));
  }
  float h; cudaMemcpy(&h, x, 4, cudaMemcpyDeviceToHost); printf("check %.3f\n", h);
  return 0;
}
