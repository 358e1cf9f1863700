// micro-benchmark: does the packed fp32 FMA (fma.rn.f32x2, SASS FFMA2) of sm_100a raise the
// fp32 rate per issue slot?  Used to decide the two-targets-per-lane layout of the tree walk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
  // 8 independent accumulator pairs per thread
  float x0[8], x1[8];
  for (int i = 0; i < 8; i++) { x0[i] = threadIdx.x * 1e-3f + i; x1[i] = x0[i] + 0.5f; }
  unsigned long long aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  int ic = threadIdx.x;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {                 // 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 8; i++) { x0[i] = fmaf(x0[i], a, b); x1[i] = fmaf(x1[i], a, b); }
    } else if (MODE == 1) {          // 8 FFMA2 (same flops)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        unsigned long long v;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x0[i]), "f"(x1[i]));
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0[i]), "=f"(x1[i]) : "l"(v));
      }
    } else if (MODE == 2) {          // 8 FFMA2 + 8 integer ops (do the freed issue slots help?)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        unsigned long long v;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x0[i]), "f"(x1[i]));
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0[i]), "=f"(x1[i]) : "l"(v));
        ic = (ic ^ (ic >> 3)) + i;
      }
    } else {                         // 16 scalar FFMA + 8 integer ops
#pragma unroll
      for (int i = 0; i < 8; i++) { x0[i] = fmaf(x0[i], a, b); x1[i] = fmaf(x1[i], a, b); ic = (ic ^ (ic >> 3)) + i; }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += x0[i] + x1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + ic;
}

template <int MODE>
void run(const char *name, float *d) {
  const int iters = 20000, blocks = 148 * 8;
  k<MODE><<<blocks, 256>>>(d, 100, 0.999f, 0.001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(d, iters, 0.999f, 0.001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)blocks * 256 * iters * 16;
  printf("%-28s %8.3f ms  %7.2f TFLOP/s fp32 (2 flop/fma)  %6.2f G warp-fma-lanes/s\n", name, ms, 2 * fma / ms * 1e-9, fma / 32 / ms * 1e-6);
}

int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<0>("scalar FFMA x16", d);
  run<1>("FFMA2 x8", d);
  run<2>("FFMA2 x8 + 8 int", d);
  run<3>("scalar FFMA x16 + 8 int", d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
