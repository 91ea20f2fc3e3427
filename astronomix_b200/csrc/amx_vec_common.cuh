// amx_vec_common.cuh -- helpers shared by the z-vectorised kernel files (amx_vec.cu, amx_pair.cu):
// streamed 64-bit global / shared accessors, boundary index mapping, launch-shape helpers.
#pragma once
#include "amx_internal.h"
#include "amx_math.cuh"

#include <cstdint>
#include <cstdlib>
#include <type_traits>

#if AMX_STRICT
#define AMX_NS amx_strict
#else
#define AMX_NS amx_fast
#endif

namespace AMX_NS {
using namespace amx;
namespace vec {


AMX_DI F2 max2(F2 a, F2 b) { return mk2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
AMX_DI F2 abs2(F2 a) { return mk2(fabsf(a.x), fabsf(a.y)); }

// streamed global loads: every input line is used once per block, so it must not take an L1
// line (the L1 left beside ~180 KB of shared memory is smaller than the loads in flight)
template <int LD>
AMX_DI F2 ldg2(const float* p) {
  F2 r;
  if constexpr (LD == 0) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    r = mk2(v.x, v.y);
  } else if constexpr (LD == 1) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  }
  return r;
}
template <int LD>
AMX_DI float ldg1(const float* p) {
  float r;
  if constexpr (LD == 0) r = __ldg(p);
  else asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
AMX_DI void stg2(float* p, F2 v) { *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y); }
AMX_DI F2 lds2(const float* p) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  return mk2(v.x, v.y);
}
AMX_DI void sts2(float* p, F2 v) { *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y); }

AMX_DI int map_index(int i, int n, int bcl, int bcr, bool& flip) {
  flip = false;
  if (i < 0) {
    if (bcl == BC_PERIODIC) return i + n;
    if (bcl == BC_OPEN) return 0;
    if (bcl == BC_REFLECTIVE) { flip = true; return -1 - i; }
    return i;
  }
  if (i >= n) {
    if (bcr == BC_PERIODIC) return i - n;
    if (bcr == BC_OPEN) return n - 1;
    if (bcr == BC_REFLECTIVE) { flip = true; return 2 * n - 1 - i; }
    return i;
  }
  return i;
}

// ---- shared-memory mbarriers (split arrive / wait, TMA completion)
AMX_DI unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
AMX_DI void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
AMX_DI void mbar_arrive(unsigned long long* b) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(smem_u32(b)) : "memory");
}
AMX_DI void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// AMX_MBAR_HINT_NS > 0: suspend-time hint of try_wait (a waiting warp sleeps that long before it polls again
// instead of the default, so that waiters -- the halo warp most of all -- take fewer issue slots from the warps
// that work)
#ifndef AMX_MBAR_HINT_NS
#define AMX_MBAR_HINT_NS 0
#endif
AMX_DI void mbar_wait(unsigned long long* b, unsigned parity) {
#if AMX_MBAR_HINT_NS > 0
  asm volatile(
      "{ .reg .pred p;\n"
      "W_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=: }"
      ::"r"(smem_u32(b)), "r"(parity), "r"((unsigned)AMX_MBAR_HINT_NS) : "memory");
#else
  asm volatile(
      "{ .reg .pred p;\n"
      "W_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=: }"
      ::"r"(smem_u32(b)), "r"(parity) : "memory");
#endif
}
// one arrival per warp: the lanes' shared-memory writes are ordered before lane 0's (releasing) arrive
AMX_DI void mbar_arrive_warp(unsigned long long* b) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(b);
}

AMX_DI unsigned long long bits_of(float v) { return (unsigned long long)__float_as_uint(v); }

// ---- limited slopes and face states of the plane-streaming hydro kernels (amx_vec.cu, amx_tma_stage.cu)
template <int LIM>
AMX_DI F2 half_slope(F2 dm, F2 dp, int lim, const Consts<F2>& c) {
#if AMX_STRICT
  const F2 a = dm / c.dx, b = dp / c.dx;
  const F2 g = (LIM == LIM_MINMOD) ? minmod(a, b) : limited_slope(a, b, lim, c);
  return (g * c.dx) / F2(2);
#else
  if constexpr (LIM == LIM_MINMOD) {
    return F2(0.5f) * minmod(dm, dp);
  } else {
    const F2 g = limited_slope(dm * c.inv_dx, dp * c.inv_dx, lim, c);
    return (g * c.dx) * F2(0.5f);
  }
#endif
}

// The slope quantity a thread keeps or shares per variable, and the two face states it gives:
// FAST + minmod: the unscaled minmod(dm, dp), faces s +- 0.5 * q as one FFMA2 each; otherwise the
// half slope (g * dx) / 2 in the reference's operation order, faces s +- q.
template <int LIM>
AMX_DI F2 slope_q(F2 dm, F2 dp, int lim, const Consts<F2>& c) {
#if !AMX_STRICT
  if constexpr (LIM == LIM_MINMOD) return minmod(dm, dp);
#endif
  return half_slope<LIM>(dm, dp, lim, c);
}
template <int LIM>
AMX_DI F2 face_hi(F2 s, F2 q) {
#if !AMX_STRICT
  if constexpr (LIM == LIM_MINMOD) return fma2(F2(0.5f), q, s);
#endif
  return s + q;
}
template <int LIM>
AMX_DI F2 face_lo(F2 s, F2 q) {
#if !AMX_STRICT
  if constexpr (LIM == LIM_MINMOD) return fma2(F2(-0.5f), q, s);
#endif
  return s - q;
}

AMX_DI Consts<F2> make_consts2(const NumPar& np) {
  Consts<F2> c;
  const float gamma = (float)np.gamma;
  const float gm1 = gamma - 1.0f;
  c.gamma = F2(gamma);
  c.gm1 = F2(gm1);
  c.inv_gm1 = F2(1.0f / gm1);
  c.dx = F2((float)np.dx);
  c.inv_dx = F2((float)(1.0 / np.dx));
  c.half_dx = F2((float)(np.dx / 2));
  c.va_eps = F2((float)(3 * np.dx));
  return c;
}

// A column is the element offset of its plane-0 cell, shared by every array of the launch
// (vec_mag_supported checks that all views have the same x / y strides).
using Col = long long;
template <int N> using IC = std::integral_constant<int, N>;

// number of x chunks minimising (waves of blocks) x (planes per chunk + prologue planes)
inline int pick_chunks(int nx, long long tiles, int blocks_per_sm, int prologue) {
  const long long slots = 148LL * blocks_per_sm;
  int best = 1;
  long long best_cost = -1;
  const int maxch = nx / 8 > 0 ? nx / 8 : 1;
  for (int nch = 1; nch <= maxch; ++nch) {
    const long long blocks = tiles * nch;
    const long long waves = (blocks + slots - 1) / slots;
    const long long cost = waves * ((nx + nch - 1) / nch + prologue);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = nch; }
  }
  return best;
}

// Equal shares of a launch's plane-tiles for a grid that is resident all at once: `tiles` columns of nx
// planes, 148 * blocks_per_sm blocks (fewer when a share would drop below 16 planes, where the 4-plane
// prologue of a segment dominates).  The x axis is cut into chunks of about one share, the work list runs
// (chunk, tile, plane): a block gets `work` consecutive entries, i.e. two or three segments of neighbouring
// tiles, and neighbouring tiles run only `work - xchunk` planes apart (their halo rows meet in L2).
inline int pick_shares(int nx, long long tiles, int blocks_per_sm, int& xchunk, int& work) {
  const long long slots = 148LL * blocks_per_sm;
  const long long total = tiles * nx;
  long long nb = total / 16 < slots ? (total / 16 > 0 ? total / 16 : 1) : slots;
  long long w = (total + nb - 1) / nb;
  int nch = (int)((nx + w / 2) / w);
  if (nch < 1) nch = 1;
  xchunk = (nx + nch - 1) / nch;
  nch = (nx + xchunk - 1) / xchunk;
  const long long padded = (long long)nch * tiles * xchunk;  // a ragged last chunk leaves dead entries
  w = (padded + nb - 1) / nb;
  // AMX_SHARE_SKEW=k (tests): shares k entries longer, so that they stop lining up with the tiles and
  // every block runs several segments, short ones included
  static const int skew = [] { const char* e = getenv("AMX_SHARE_SKEW"); return e ? atoi(e) : 0; }();
  if (skew > 0) w += skew;
  work = (int)w;
  return (int)((padded + w - 1) / w);
}

inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
inline bool view_ok(const View& v, int nd) {
  const int la = nd - 1;
  if (v.s[la] != 1 || (v.vs & 1) || !aligned8(v.p)) return false;
  for (int ax = 0; ax < la; ++ax)
    if (v.s[ax] & 1) return false;
  return true;
}

}  // namespace vec
}  // namespace AMX_NS
