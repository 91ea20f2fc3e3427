// TMA probe: which forms of tensor-map prefetch / load work on this driver + sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../astronomix_b200/csrc tma_probe.cu -o tma_probe
// run:   ./tma_probe <test> ; tests are separate processes because a TMA fault is sticky
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "amx_tma.h"
using namespace amx;

struct Maps { CUtensorMap m; int on; };

__global__ void k_prefetch(const __grid_constant__ Maps maps, int z, int y, int x, int mode) {
  if (mode == 0) { if (threadIdx.x == 0) tma_prefetch_4d(&maps.m, z, y, x, 0); }
  else if (mode == 1) {
    if (threadIdx.x < 32) {
      unsigned pred;
      asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.b32 %0, 1, 0, P; }" : "=r"(pred));
      if (pred) tma_prefetch_4d(&maps.m, z, y, x, 0);
    }
  } else if (mode == 2) { if (threadIdx.x == 480) tma_prefetch_4d(&maps.m, z, y, x, 0); }
}

__global__ void k_load(const __grid_constant__ Maps maps, int z, int y, int x, float* out, int nfl) {
  extern __shared__ __align__(128) float sm[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(nfl * 4) : "memory");
    tma_load_4d(sm, &maps.m, z, y, x, 0, &bar);
  }
  asm volatile("{ .reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD: }" ::"r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
  for (int i = threadIdx.x; i < nfl; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char** argv) {
  const int test = argc > 1 ? atoi(argv[1]) : 0;
  const int R = 3, NX = 16, NY = 32, NZ = 64;
  std::vector<float> h((size_t)R * NX * NY * NZ);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  View v; v.p = d; v.s[0] = NY * NZ; v.s[1] = NZ; v.s[2] = 1; v.vs = (long long)NX * NY * NZ;
  Geom g; g.nd = 3; g.n[0] = NX; g.n[1] = NY; g.n[2] = NZ;
  unsigned bz = 36, by = 6, br = 3;
  if (test == 5) { bz = 68; by = 15; }      // box larger than the array in z (64)
  Maps maps; maps.on = 1;
  bool ok = tma_map_rows_3d(&maps.m, v, g, R, 0, bz, by, br);
  printf("test %d encode %d\n", test, (int)ok);
  if (!ok) return 1;
  cudaError_t e = cudaSuccess;
  if (test == 0) k_prefetch<<<1, 512>>>(maps, 4, 4, 2, 0);          // in bounds, thread 0
  if (test == 1) k_prefetch<<<1, 512>>>(maps, 4, 4, 2, 1);          // in bounds, elect
  if (test == 2) k_prefetch<<<1, 512>>>(maps, 4, 4, 2, 2);          // in bounds, thread 480
  if (test == 3) k_prefetch<<<1, 512>>>(maps, -2, -2, 2, 0);        // negative coordinates
  if (test == 4) k_prefetch<<<1, 512>>>(maps, 40, 30, 2, 0);        // box sticks out on the high side
  if (test == 5) k_prefetch<<<1, 512>>>(maps, 0, 0, 2, 0);          // oversized box
  if (test == 10) k_prefetch<<<1, 512>>>(maps, 2, 4, 2, 0);         // z not 16-byte aligned, in bounds
  if (test == 11) k_prefetch<<<1, 512>>>(maps, -4, 4, 2, 0);        // z negative, aligned
  if (test == 12) k_prefetch<<<1, 512>>>(maps, 4, -2, 2, 0);        // y negative
  if (test == 13) k_prefetch<<<1, 512>>>(maps, 62, 4, 2, 0);        // z unaligned, sticks out
  if (test == 14) k_prefetch<<<1, 512>>>(maps, 4, 4, -1, 0);        // x negative
  if (test >= 6 && test < 10 || test >= 20) {
    const int nfl = bz * by * br;
    float* o; cudaMalloc(&o, nfl * 4);
    int z = test == 6 ? 4 : -2, y = test == 6 ? 4 : -2;        // 6: in bounds, 7: negative (zero fill)
    if (test == 20) { z = 2; y = 4; }     // unaligned z, in bounds
    if (test == 21) { z = -4; y = 4; }    // negative aligned z
    if (test == 22) { z = 4; y = -2; }    // negative y
    if (test == 23) { z = 62; y = 4; }    // unaligned, sticks out
    k_load<<<1, 128, nfl * 4>>>(maps, z, y, 2, o, nfl);
    e = cudaDeviceSynchronize();
    std::vector<float> r(nfl);
    cudaMemcpy(r.data(), o, nfl * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < (int)br; ++rr) for (int j = 0; j < (int)by; ++j) for (int i = 0; i < (int)bz; ++i) {
      const int gz = z + i, gy = y + j;
      float want = 0.f;
      if (gz >= 0 && gz < NZ && gy >= 0 && gy < NY) want = h[((size_t)rr * NX + 2) * NY * NZ + (size_t)gy * NZ + gz];
      if (r[(rr * by + j) * bz + i] != want) ++bad;
    }
    printf("load mismatches %d of %d\n", bad, nfl);
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  printf("test %d result: %s\n", test, cudaGetErrorString(e));
  return 0;
}
