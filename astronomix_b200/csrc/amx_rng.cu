// amx_rng.cu -- jax's default PRNG on the device, for the turbulent driving: the reference
// draws raw_noise = normal(sk1, (3, n, n, n)) + 1j * normal(sk2, (3, n, n, n)) every step from keys split off
// jax.random.key(42) (_turbulent_forcing.py:48-51, time_integration.py:787-789).  Threefry-2x32-20 block
// function, jax's two counter layouts (threefry_partitionable on / off), bits -> uniform in
// [nextafter(-1, 0), 1) -> sqrt(2) * erf_inv (fp32: XLA's ErfInv32 = Giles' polynomial; fp64: erfinv()).
// oracle/threefry.py is the NumPy restatement the tests compare with (bit-exact for the bits and the uniforms).
#include "../../include/amx.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int blk = 0; blk < 5; ++blk) {
    const int r0 = (blk & 1) ? 17 : 13, r1 = (blk & 1) ? 29 : 15, r2 = (blk & 1) ? 16 : 26, r3 = (blk & 1) ? 24 : 6;
    x0 += x1; x1 = rotl(x1, r0); x1 ^= x0;
    x0 += x1; x1 = rotl(x1, r1); x1 ^= x0;
    x0 += x1; x1 = rotl(x1, r2); x1 ^= x0;
    x0 += x1; x1 = rotl(x1, r3); x1 ^= x0;
    x0 += ks[(blk + 1) % 3];
    x1 += ks[(blk + 2) % 3] + (uint32_t)(blk + 1);
  }
}

__device__ __forceinline__ float erfinv32(float x) {  // xla/client/lib/math.cc ErfInv32
  float w = -log1pf(-(x * x));
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p * w;
    p = -3.5233877e-06f + p * w;
    p = -4.39150654e-06f + p * w;
    p = 0.00021858087f + p * w;
    p = -0.00125372503f + p * w;
    p = -0.00417768164f + p * w;
    p = 0.246640727f + p * w;
    p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = 0.000100950558f + p * w;
    p = 0.00134934322f + p * w;
    p = -0.00367342844f + p * w;
    p = 0.00573950773f + p * w;
    p = -0.0076224613f + p * w;
    p = 0.00943887047f + p * w;
    p = 1.00167406f + p * w;
    p = 2.83297682f + p * w;
  }
  return fabsf(x) == 1.0f ? copysignf(INFINITY, x) : p * x;
}

__device__ __forceinline__ float normal_from_bits(uint32_t bits) {
  const float u01 = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(u01, __fsub_rn(1.0f, lo)), lo));
  return __fmul_rn(1.41421354f, erfinv32(u));
}
__device__ __forceinline__ double normal_from_bits(uint64_t bits) {
  const double u01 = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ull)) - 1.0;
  const double lo = -0.99999999999999989;  // nextafter(-1, 0)
  const double u = fmax(lo, __dadd_rn(__dmul_rn(u01, __dsub_rn(1.0, lo)), lo));
  return 1.4142135623730951 * erfinv(u);
}

// partitionable layout: sample i <- block(counter = (i >> 32, i)); fp32 bits = out0 ^ out1, fp64 bits = out0:out1
template <typename T>
__global__ void normal_partitionable(uint32_t k0, uint32_t k1, long long n, T* out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t x0 = (uint32_t)((unsigned long long)i >> 32), x1 = (uint32_t)i;
    threefry2x32(k0, k1, x0, x1);
    if constexpr (sizeof(T) == 4) out[i] = normal_from_bits(x0 ^ x1);
    else out[i] = normal_from_bits(((uint64_t)x0 << 32) | x1);
  }
}

// original layout: the 32-bit words w[j], j < W (W = n words for fp32, 2n for fp64, padded to even), come from
// blocks b < W/2 with counter (b, b + W/2): w[b] = out0, w[b + W/2] = out1; fp64 sample i = w[i] : w[i + n]
__global__ void normal_original_f32(uint32_t k0, uint32_t k1, long long n, float* out) {
  const long long half = (n + 1) / 2;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < half; b += (long long)gridDim.x * blockDim.x) {
    uint32_t x0 = (uint32_t)b, x1 = (uint32_t)(b + half);
    if ((n & 1) && b + half == n) x1 = 0u;  // the padding word of an odd count is a zero counter
    threefry2x32(k0, k1, x0, x1);
    out[b] = normal_from_bits(x0);
    if (b + half < n) out[b + half] = normal_from_bits(x1);
  }
}
__global__ void normal_original_f64(uint32_t k0, uint32_t k1, long long n, double* out) {
  // W = 2n words (even): block b < n gives w[b] (high word of sample b) and w[b + n] (low word of sample b)
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += (long long)gridDim.x * blockDim.x) {
    uint32_t x0 = (uint32_t)b, x1 = (uint32_t)(b + n);
    threefry2x32(k0, k1, x0, x1);
    out[b] = normal_from_bits(((uint64_t)x0 << 32) | x1);
  }
}

// samples [offset, offset + count) of a draw of `total` samples, either layout (one block per sample): what a
// slab rank needs of a global (3, n, n, n) draw
template <typename T>
__global__ void normal_range(uint32_t k0, uint32_t k1, long long total, long long offset, long long count, int partitionable,
                             T* out) {
  constexpr bool f64 = sizeof(T) == 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const long long j = offset + i;
    if (partitionable) {
      uint32_t x0 = (uint32_t)((unsigned long long)j >> 32), x1 = (uint32_t)j;
      threefry2x32(k0, k1, x0, x1);
      if constexpr (f64) out[i] = normal_from_bits(((uint64_t)x0 << 32) | x1);
      else out[i] = normal_from_bits(x0 ^ x1);
    } else if constexpr (f64) {
      uint32_t x0 = (uint32_t)j, x1 = (uint32_t)(j + total);
      threefry2x32(k0, k1, x0, x1);
      out[i] = normal_from_bits(((uint64_t)x0 << 32) | x1);
    } else {
      const long long half = (total + 1) / 2;
      const bool second = j >= half;
      const long long b = second ? j - half : j;
      uint32_t x0 = (uint32_t)b, x1 = (uint32_t)(b + half);
      if ((total & 1) && b + half == total) x1 = 0u;
      threefry2x32(k0, k1, x0, x1);
      out[i] = normal_from_bits(second ? x1 : x0);
    }
  }
}

// ---- turbulent driving, fused (_turbulent_forcing.py:15-65 up to the inverse FFT) -------------------------
// One thread per wave vector of the rank's k_x slab [x0, x0 + nl) x n x n: the six normal samples of the
// step (real parts from key A, imaginary parts from key B, component c at linear index ((c n + ix) n + iy) n + iz
// of a (3, n, n, n) draw), spectrum sqrt(k^6 exp(-8 k / k_peak)), DC mode zero, compressive part projected
// out; writes the three complex coefficients (interleaved re, im) that the inverse FFT reads.
template <typename T>
__device__ __forceinline__ T sample_at(uint32_t k0, uint32_t k1, long long total, long long j, int partitionable) {
  constexpr bool f64 = sizeof(T) == 8;
  uint32_t x0, x1;
  bool second = false;
  if (partitionable) {
    x0 = (uint32_t)((unsigned long long)j >> 32); x1 = (uint32_t)j;
  } else if (f64) {
    x0 = (uint32_t)j; x1 = (uint32_t)(j + total);
  } else {
    const long long half = (total + 1) / 2;
    second = j >= half;
    const long long b = second ? j - half : j;
    x0 = (uint32_t)b; x1 = (uint32_t)(b + half);
    if ((total & 1) && b + half == total) x1 = 0u;
  }
  threefry2x32(k0, k1, x0, x1);
  if constexpr (f64) return normal_from_bits(((uint64_t)x0 << 32) | x1);
  else return normal_from_bits(partitionable ? (x0 ^ x1) : (second ? x1 : x0));
}

template <typename T>
__global__ void forcing_spectrum_kernel(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, int n, int x0, int nl,
                                        double box, int partitionable, T* out) {
  const long long cells = (long long)nl * n * n, total = 3ll * n * n * n;
  const T two_pi = (T)6.283185307179586476925286766559;
  const T d = (T)(box / n);
  const T kpk = (T)(4.0 * 3.14159265358979323846 / box);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
    const int iz = (int)(i % n), iy = (int)((i / n) % n), ixl = (int)(i / ((long long)n * n));
    const int ix = x0 + ixl;
    // numpy.fft.fftfreq(n, d): f = (i < (n + 1) / 2 ? i : i - n) / (n d); k = 2 pi f
    auto kof = [&](int q) { return two_pi * ((T)(q < (n + 1) / 2 ? q : q - n) / ((T)n * d)); };
    const T kx = kof(ix), ky = kof(iy), kz = kof(iz);
    const T k2 = kx * kx + ky * ky + kz * kz;
    const T kk = sqrt(k2);
    const T k3 = kk * kk * kk;
    const T sp = sqrt(k3 * k3 * exp((T)-8.0 * kk / kpk));
    const bool dc = (ix | iy | iz) == 0;
    T re[3], im[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long j = ((long long)c * n + ix) * n * n + (long long)iy * n + iz;
      re[c] = dc ? (T)0 : sp * sample_at<T>(a0, a1, total, j, partitionable);
      im[c] = dc ? (T)0 : sp * sample_at<T>(b0, b1, total, j, partitionable);
    }
    const T k2s = k2 == (T)0 ? (T)1 : k2;
    const T dre = dc ? (T)0 : (kx * re[0] + ky * re[1] + kz * re[2]) / k2s;
    const T dim = dc ? (T)0 : (kx * im[0] + ky * im[1] + kz * im[2]) / k2s;
    const T kv[3] = {kx, ky, kz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T* o = out + 2 * ((long long)c * cells + i);
      o[0] = re[c] - kv[c] * dre;
      o[1] = im[c] - kv[c] * dim;
    }
  }
}

// The two sums of the amplitude equation (_turbulent_forcing.py:103-104) over the rank's cells, in double,
// deterministic: per-block partial sums, then block 0 of a second launch adds them in index order.
template <typename T>
__global__ void __launch_bounds__(256) forcing_sums_kernel(const T* rho, const T* vx, const T* vy, const T* vz, const T* w,
                                                         long long cells, double* partial) {
  double sa = 0.0, sb = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
    const double r = rho[i], w0 = w[i], w1 = w[cells + i], w2 = w[2 * cells + i];
    sa += r * (w0 * w0 + w1 * w1 + w2 * w2);
    sb += r * ((double)vx[i] * w0 + (double)vy[i] * w1 + (double)vz[i] * w2);
  }
  __shared__ double red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sa; red[1][threadIdx.x >> 5] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 8; ++k) { a += red[0][k]; b += red[1][k]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
  }
}
__global__ void forcing_sums_finish(const double* partial, int nblocks, double* out2) {
  double a = 0.0, b = 0.0;
  for (int k = 0; k < nblocks; ++k) { a += partial[2 * k]; b += partial[2 * k + 1]; }
  out2[0] = 0.5 * a;
  out2[1] = b;
}

}  // namespace

extern "C" {

int amx_threefry_normal(uint32_t key0, uint32_t key1, int64_t count, void* out, int dtype, int partitionable,
                        void* stream) {
  if (count < 0 || (count > 0 && !out) || (dtype != AMX_F32 && dtype != AMX_F64)) return AMX_EINVAL;
  if (!partitionable && count * (dtype == AMX_F64 ? 2 : 1) > 0xFFFFFFFFll) return AMX_EUNSUPPORTED;  // 32-bit iota
  if (count == 0) return AMX_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  const long long work = partitionable || dtype == AMX_F64 ? count : (count + 1) / 2;
  const unsigned blocks = (unsigned)((work + threads - 1) / threads < 148ll * 16 ? (work + threads - 1) / threads : 148ll * 16);
  if (partitionable) {
    if (dtype == AMX_F32) normal_partitionable<float><<<blocks, threads, 0, st>>>(key0, key1, count, static_cast<float*>(out));
    else normal_partitionable<double><<<blocks, threads, 0, st>>>(key0, key1, count, static_cast<double*>(out));
  } else {
    if (dtype == AMX_F32) normal_original_f32<<<blocks, threads, 0, st>>>(key0, key1, count, static_cast<float*>(out));
    else normal_original_f64<<<blocks, threads, 0, st>>>(key0, key1, count, static_cast<double*>(out));
  }
  return cudaGetLastError() == cudaSuccess ? AMX_OK : AMX_ECUDA;
}

int amx_threefry_normal_range(uint32_t key0, uint32_t key1, int64_t total, int64_t offset, int64_t count, void* out,
                              int dtype, int partitionable, void* stream) {
  if (total < 0 || offset < 0 || count < 0 || offset + count > total || (count > 0 && !out) ||
      (dtype != AMX_F32 && dtype != AMX_F64))
    return AMX_EINVAL;
  if (!partitionable && total * (dtype == AMX_F64 ? 2 : 1) > 0xFFFFFFFFll) return AMX_EUNSUPPORTED;
  if (count == 0) return AMX_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  const unsigned blocks = (unsigned)((count + threads - 1) / threads < 148ll * 16 ? (count + threads - 1) / threads : 148ll * 16);
  if (dtype == AMX_F32)
    normal_range<float><<<blocks, threads, 0, st>>>(key0, key1, total, offset, count, partitionable, static_cast<float*>(out));
  else
    normal_range<double><<<blocks, threads, 0, st>>>(key0, key1, total, offset, count, partitionable, static_cast<double*>(out));
  return cudaGetLastError() == cudaSuccess ? AMX_OK : AMX_ECUDA;
}

int amx_forcing_spectrum(uint32_t re_key0, uint32_t re_key1, uint32_t im_key0, uint32_t im_key1, int n, int x0, int nl,
                         double box_size, void* out_complex, int dtype, int partitionable, void* stream) {
  if (n < 2 || x0 < 0 || nl < 1 || x0 + nl > n || !out_complex || (dtype != AMX_F32 && dtype != AMX_F64) || !(box_size > 0))
    return AMX_EINVAL;
  if (!partitionable && 3ll * n * n * n * (dtype == AMX_F64 ? 2 : 1) > 0xFFFFFFFFll) return AMX_EUNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long cells = (long long)nl * n * n;
  const unsigned blocks = (unsigned)((cells + 255) / 256 < 148ll * 32 ? (cells + 255) / 256 : 148ll * 32);
  if (dtype == AMX_F32)
    forcing_spectrum_kernel<float><<<blocks, 256, 0, st>>>(re_key0, re_key1, im_key0, im_key1, n, x0, nl, box_size,
                                                           partitionable, static_cast<float*>(out_complex));
  else
    forcing_spectrum_kernel<double><<<blocks, 256, 0, st>>>(re_key0, re_key1, im_key0, im_key1, n, x0, nl, box_size,
                                                            partitionable, static_cast<double*>(out_complex));
  return cudaGetLastError() == cudaSuccess ? AMX_OK : AMX_ECUDA;
}

int amx_forcing_sums(const void* rho, const void* vx, const void* vy, const void* vz, const void* w3, int64_t cells,
                     void* scratch, void* sums2_out, int dtype, void* stream) {
  if (!rho || !vx || !vy || !vz || !w3 || !scratch || !sums2_out || cells < 1 || (dtype != AMX_F32 && dtype != AMX_F64))
    return AMX_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (int)((cells + 255) / 256 < 1024 ? (cells + 255) / 256 : 1024);
  double* partial = static_cast<double*>(scratch);
  if (dtype == AMX_F32)
    forcing_sums_kernel<float><<<blocks, 256, 0, st>>>((const float*)rho, (const float*)vx, (const float*)vy, (const float*)vz,
                                                       (const float*)w3, cells, partial);
  else
    forcing_sums_kernel<double><<<blocks, 256, 0, st>>>((const double*)rho, (const double*)vx, (const double*)vy,
                                                        (const double*)vz, (const double*)w3, cells, partial);
  forcing_sums_finish<<<1, 1, 0, st>>>(partial, blocks, static_cast<double*>(sums2_out));
  return cudaGetLastError() == cudaSuccess ? AMX_OK : AMX_ECUDA;
}

}  // extern "C"
