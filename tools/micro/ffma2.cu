// Microbenchmark: issue throughput of FFMA vs FFMA2 (fma.rn.f32x2) vs FMNMX on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ffma2 tools/micro/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float& dx, float& dy, float ax, float ay, float bx, float by) {
  asm volatile("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rd, {%0, %1};"
      " fma.rn.f32x2 rd, ra, rb, rd; mov.b64 {%0, %1}, rd; }"
      : "+f"(dx), "+f"(dy) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) { r[i] = fmaf(r[i], a, b); r[i + 1] = fmaf(r[i + 1], a, b); }
      else if (MODE == 1) { ffma2(r[i], r[i + 1], a, a, b, b); }
      else if (MODE == 2) { r[i] = fminf(r[i], a) ; r[i + 1] = fmaxf(r[i + 1], b); }
      else { r[i] = fmaf(r[i], a, b); r[i + 1] = fminf(r[i + 1], a); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int elems_per_inst) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, 16, 0.999f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iters, 0.999f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // element-ops per thread per iter = 16
  double ops = 148.0 * 8 * 256 * iters * 16.0;
  double warp_inst = ops / 32 / elems_per_inst;
  printf("%-10s %.3f ms  %.2f Gelem-ops/s  %.3f warp-inst/clk/SM (at 1.965 GHz)\n", name, ms, ops / ms * 1e-6,
         warp_inst / (ms * 1e-3) / 1.965e9 / 148);
  cudaFree(out);
}
int main() {
  run<0>("FFMA", 1); run<1>("FFMA2", 2); run<2>("FMNMX", 1); run<3>("FFMA+FMNMX", 1);
  return 0;
}
