// amx_vec.cu -- z-vectorised plane-streaming kernels (fp32): every thread owns a PAIR of cells
// adjacent along the contiguous axis, global traffic is 64-bit, the arithmetic of the two cells
// is issued as packed f32x2 instructions (FADD2 / FMUL2 / FFMA2 on sm_100a), the x direction of
// the stencil lives in registers and only the (y, z) neighbours go through shared memory.
//
// mag_iter_vec<ND>: one evaluation R_{k+1} = R_0 - dt * Psi(arg(R_k, R_0)) of the Pang-Wu fixed
// point (_magnetic_field_update.py:82-165; curls: _vector_maths.py:55-100).
//   * block = 8 compute warps + 1 halo warp.  3-D: the compute threads cover a 16 (y) x 32 (z)
//     column tile as 16 x 16 pairs; the halo warp loads the two y-halo rows (one pair per lane)
//     and the 2 x 16 z-halo cells (one scalar per lane).  2-D: 1 x 512 tile, two halo scalars.
//   * the block marches along x over a chunk of planes.  Per plane a thread loads the raw
//     B_0, v_0, rho, B_k, v_k of its pair (13 x LDG.64), forms the fixed-point argument and
//     E = B x v once, publishes (B, E) to a 3-slot shared-memory ring for its y / z neighbours and
//     keeps its own column of (B, E) in registers for the x differences: one barrier per plane.
//   * 41 load/store instructions per pair per plane (the scalar tiled kernel: ~100 per cell).
// Compiled twice like the other kernel files (fast / strict arithmetic); STRICT keeps the
// reference's operation order (no contraction), so it is bitwise equal to the generic kernel.
#include "amx_vec_common.cuh"

#include <cstdlib>

namespace AMX_NS {
using namespace amx;

namespace vec {

// 7 compute warps + 1 halo warp (register allocation is per 4 warps: 8 warps x 128 registers
// x 2 blocks fills the register file exactly)
// CW compute warps + 1 halo warp
template <int ND, int CW> struct VT;
template <int CW> struct VT<3, CW> { static constexpr int TY = 2 * CW, TQ = 16; };
template <int CW> struct VT<2, CW> { static constexpr int TY = 1, TQ = 32 * CW; };

template <int ND, int CW>
constexpr size_t mag_vec_smem_bytes() {
  constexpr int TZ = 2 * VT<ND, CW>::TQ, WY = ND == 3 ? VT<ND, CW>::TY + 2 : 1, RS = TZ + 4;
  return ((size_t)2 * 6 * WY * RS + (size_t)2 * 13 * 2 * 32 * CW) * sizeof(float);
}

// raw fixed-point data of one pair / one halo cell in one plane
struct Raw {
  F2 B0[3], v0[3], rho, Bk[3], vk[3];
};

// MODE 0: first iterate (R_k = R_0, Bk / vk absent); 1: implicit midpoint; 2: implicit Euler
#ifndef AMX_MAG_MINB8
#define AMX_MAG_MINB8 1  // blocks per SM of the 8-compute-warp flavour (AMX_VEC_CW=8)
#endif
template <int ND, int MODE, int LD, int CW>
__global__ void __launch_bounds__(32 * CW + 32, CW == 7 ? 2 : AMX_MAG_MINB8) mag_iter_vec(const MagIterArgs a, const int xchunk) {
  if (a.ctl_ro) {
    if (a.ctl_ro->done) return;
    if (a.fix_base >= 0) {
      if (!a.ctl_ro->mag_fix || a.ctl_ro->mag_fix_base != a.fix_base) return;
    } else if (!a.first && !a.ctl_ro->mag_active) return;
  }
  constexpr int kComputeThreads = 32 * CW, kThreads = kComputeThreads + 32;
  constexpr int TY = VT<ND, CW>::TY, TQ = VT<ND, CW>::TQ, TZ = 2 * TQ;
  constexpr int WY = ND == 3 ? TY + 2 : 1, RS = TZ + 4, YO = ND == 3 ? 1 : 0, LA = ND - 1;
  constexpr int PL = WY * RS;   // floats per field per ring slot
  constexpr int SLOT = 6 * PL;  // fields: B0 B1 B2 E0 E1 E2
  constexpr bool have_k = MODE != 0;
  extern __shared__ __align__(16) float smem[];  // ring [2][6][WY][RS] (own pairs at column 2 + 2*tq), stash

  const Geom& g = a.g;
  const int tid = threadIdx.x;
  const bool compute = tid < kComputeThreads;
  const int z0 = blockIdx.x * TZ;
  const int y0 = ND == 3 ? blockIdx.y * TY : 0;
  const int chunk = ND == 3 ? blockIdx.z : blockIdx.y;
  const int xs = chunk * xchunk;
  const int xe = min(xs + xchunk, g.n[0]);
  if (xs >= xe) return;
  const float dt = *static_cast<const float*>(a.dt);
#if AMX_STRICT
  const float dxf = (float)a.np.dx;
#else
  const float hdx = 0.5f * (float)(1.0 / a.np.dx);
#endif
  // ragged tile edges: active rows / pairs of this tile
  const int ta = ND == 3 ? min(TY, g.n[1] - y0) : 1;
  const int tqa = min(TQ, (g.n[LA] - z0) >> 1);
  // common element strides of a plane (x) and a row (y); uniform per-field base pointers
  const long long sx = a.B0.s[0], sy = ND == 3 ? a.B0.s[1] : 0;
  const float* const fB0 = static_cast<const float*>(a.B0.p);
  const float* const fG = static_cast<const float*>(a.gas0.p);
  const float* const fBk = static_cast<const float*>(a.Bk.p);
  const float* const fvk = static_cast<const float*>(a.vk.p);
  const long long vB0 = a.B0.vs, vG = a.gas0.vs, vBk = a.Bk.vs, vvk = a.vk.vs;

  auto make_col = [&](int gy, int gz, Col& c) { c = gz + (ND == 3 ? (long long)gy * sy : 0); };
  // planes -1 and n[0] are the only ones the x boundary rule can touch
  bool flo, fhi;
  const int mlo = map_index(-1, g.n[0], g.bc[0][0], g.bc[0][1], flo);
  const int mhi = map_index(g.n[0], g.n[0], g.bc[0][0], g.bc[0][1], fhi);
  auto plane_map = [&](int p, bool& fx) {
    fx = p < 0 ? flo : (p >= g.n[0] ? fhi : false);
    return p < 0 ? mlo : (p >= g.n[0] ? mhi : p);
  };

  // ---- loads
  auto load_pair = [&](const Col& c, int mx, Raw& r) {
    const long long o = c + (long long)mx * sx;
    const float* pb = fB0 + o;
    const float* pg = fG + o;
#pragma unroll
    for (int k = 0; k < 3; ++k) r.B0[k] = ldg2<LD>(pb + k * vB0);
#pragma unroll
    for (int k = 0; k < 3; ++k) r.v0[k] = k < ND ? ldg2<LD>(pg + (k + 1) * vG) : bc2(0.f);
    r.rho = ldg2<LD>(pg);
    if constexpr (have_k) {
      const float* pbk = fBk + o;
      const float* pvk = fvk + o;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        r.Bk[k] = ldg2<LD>(pbk + k * vBk);
        r.vk[k] = ldg2<LD>(pvk + k * vvk);
      }
    }
  };
  auto load_scalar = [&](const Col& c, int mx, Raw& r) {
    const long long o = c + (long long)mx * sx;
    const float* pb = fB0 + o;
    const float* pg = fG + o;
#pragma unroll
    for (int k = 0; k < 3; ++k) r.B0[k] = bc2(ldg1<LD>(pb + k * vB0));
#pragma unroll
    for (int k = 0; k < 3; ++k) r.v0[k] = k < ND ? bc2(ldg1<LD>(pg + (k + 1) * vG)) : bc2(0.f);
    r.rho = bc2(1.f);
    if constexpr (have_k) {
      const float* pbk = fBk + o;
      const float* pvk = fvk + o;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        r.Bk[k] = bc2(ldg1<LD>(pbk + k * vBk));
        r.vk[k] = bc2(ldg1<LD>(pvk + k * vvk));
      }
    }
  };
  // fixed-point argument (R_k + R_0)/2 or R_k, and E = B x v (_vector_maths.py:6-12)
  auto argument = [&](const Raw& r, bool fx, bool f2, F2 (&B)[3], F2 (&E)[3]) {
    F2 v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if constexpr (MODE == 1) {
        B[k] = (r.Bk[k] + r.B0[k]) * bc2(0.5f);
        v[k] = (r.vk[k] + r.v0[k]) * bc2(0.5f);
      } else if constexpr (MODE == 2) {
        B[k] = r.Bk[k];
        v[k] = r.vk[k];
      } else {
        B[k] = r.B0[k];
        v[k] = r.v0[k];
      }
    }
    if (fx) { B[1] = -B[1]; v[1] = -v[1]; }
    if (f2) { B[2] = -B[2]; v[2] = -v[2]; }
    E[0] = B[1] * v[2] - B[2] * v[1];
    E[1] = B[2] * v[0] - B[0] * v[2];
    E[2] = B[0] * v[1] - B[1] * v[0];
  };
  auto publish_pair = [&](float* slot, int off, const F2 (&B)[3], const F2 (&E)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      sts2(slot + k * PL + off, B[k]);
      sts2(slot + (3 + k) * PL + off, E[k]);
    }
  };
  auto publish_scalar = [&](float* slot, int off, const F2 (&B)[3], const F2 (&E)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      slot[k * PL + off] = B[k].x;
      slot[(3 + k) * PL + off] = E[k].x;
    }
  };
  // central difference 0.5*(fp - fm)/dx
  auto cd = [&](F2 fp, F2 fm) -> F2 {
#if AMX_STRICT
    const F2 h = bc2(0.5f) * (fp - fm);
    return mk2(h.x / dxf, h.y / dxf);
#else
    return bc2(hdx) * (fp - fm);
#endif
  };

  F2 ch0 = bc2(0.f), ch1 = bc2(0.f);

  if (!compute) {
    // ================= halo warp: one y-halo pair (3-D) and one lane-axis halo cell per lane.
    // Q12: a reflective y side (3-D) flips 3-vector component 2; in 2-D the lane axis IS y, so
    // its halo cells flip component 2; the z axis of a 3-D grid flips nothing.
    const int l = tid - kComputeThreads;
    Col colA, colB;
    bool haveA = false, haveB = false, flipA = false, flipB = false;
    int smA = 0, smB = 0;
    if constexpr (ND == 3) {
      const int j = l < TQ ? 0 : ta + 1;  // window row of the halo pair
      const int my = map_index(y0 - 1 + j, g.n[1], g.bc[1][0], g.bc[1][1], flipA);
      make_col(my, z0 + 2 * (l % TQ), colA);
      haveA = (l % TQ) < tqa;
      smA = j * RS + 2 + 2 * (l % TQ);
      const int r = l % 16, side = l / 16;  // up to 2 x 16 z-halo cells
      bool fz;
      const int mz = map_index(side == 0 ? z0 - 1 : z0 + 2 * tqa, g.n[2], g.bc[2][0], g.bc[2][1], fz);
      make_col(y0 + r, mz, colB);
      haveB = r < ta;
      smB = (r + YO) * RS + (side == 0 ? 1 : 2 + 2 * tqa);
    } else {
      const int mz = map_index(l == 0 ? z0 - 1 : z0 + 2 * tqa, g.n[1], g.bc[1][0], g.bc[1][1], flipB);
      make_col(0, mz, colB);
      haveB = l < 2;
      smB = l == 0 ? 1 : 2 + 2 * tqa;
    }
    Raw rawL, rawS;
    auto issue = [&](int p) {
      bool fx;
      const int mx = plane_map(p, fx);
      if (haveA) load_pair(colA, mx, rawL);
      if (haveB) load_scalar(colB, mx, rawS);
    };
    auto process = [&](int p, int s) {
      bool fx;
      (void)plane_map(p, fx);
      float* slot = smem + s * SLOT;
      F2 B[3], E[3];
      if (haveA) {
        argument(rawL, fx, flipA, B, E);
        publish_pair(slot, smA, B, E);
      }
      if (haveB) {
        argument(rawS, fx, flipB, B, E);
        publish_scalar(slot, smB, B, E);
      }
    };
    // plane p lives in ring slot (p - xs) & 1; step x turns the loads in flight into plane x+1
    issue(xs);
    process(xs, 0);
    issue(xs + 1);
    __syncthreads();
    for (int x = xs; x < xe; ++x) {
      process(x + 1, (x + 1 - xs) & 1);
      if (x + 2 <= xe) issue(x + 2);
      __syncthreads();
    }
  } else {
    // ================= compute threads: pair (ty, tq)
    const int tq = tid % TQ, ty = tid / TQ;
    const bool active = tq < tqa && ty < ta;
    Col colA;
    make_col(y0 + ty, z0 + 2 * tq, colA);
    const int smA = (ty + YO) * RS + 2 + 2 * tq;
    // components 1 and 2 of (B, E) of the own pair in planes p with p mod 3 == slot: registers
    // mirror the ring slots (component 0 is only needed of the centre plane: re-read from the ring)
    F2 RB[3][2], RE[3][2];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int k = 0; k < 2; ++k) RB[s][k] = RE[s][k] = bc2(0.f);
    Raw rawL;  // loads in flight
#pragma unroll
    for (int k = 0; k < 3; ++k) rawL.B0[k] = rawL.v0[k] = rawL.Bk[k] = rawL.vk[k] = bc2(0.f);
    rawL.rho = bc2(1.f);
    float* const fB1 = static_cast<float*>(a.Bk1.p);
    float* const fv1 = static_cast<float*>(a.vk1.p);
    const long long vB1 = a.Bk1.vs, vv1 = a.vk1.vs;
    // thread-private stash of the raw data between process(p) and update(p): [2][13][pairs]
    float* const stash = smem + 2 * SLOT + 2 * tid;
    constexpr int SP = 2 * kComputeThreads;  // floats per stashed field
    constexpr int SS = 13 * SP;              // floats per stash slot

    auto issue = [&](int p) {
      bool fx;
      const int mx = plane_map(p, fx);
      if (active) load_pair(colA, mx, rawL);
    };
    // consume the loads in flight as plane p: argument + E into register set S, publish to ring
    // slot q, park the raw values in stash slot q until update(p)   (q < 0: registers only)
    auto process = [&](int p, auto Sc, int q) {
      constexpr int S = decltype(Sc)::value;
      const bool publish = q >= 0;
      bool fx;
      (void)plane_map(p, fx);
      F2 B[3], E[3];
      argument(rawL, fx, false, B, E);
      RB[S][0] = B[1]; RB[S][1] = B[2];
      RE[S][0] = E[1]; RE[S][1] = E[2];
      if (publish && active) {  // rows / pairs beyond the grid must not touch the halo entries
        publish_pair(smem + q * SLOT, smA, B, E);
        float* st = stash + q * SS;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          sts2(st + k * SP, rawL.B0[k]);
          if (k < ND) sts2(st + (3 + k) * SP, rawL.v0[k]);
          if constexpr (have_k) {
            sts2(st + (7 + k) * SP, rawL.Bk[k]);
            sts2(st + (10 + k) * SP, rawL.vk[k]);
          }
        }
        sts2(st + 6 * SP, rawL.rho);
      }
    };
    // update of plane x (register set SC) from: ring slot q (y / z neighbours), register sets
    // SN = plane x+1 and SF = plane x-1 (x neighbours), stash slot q (R_0, R_k, rho)
    auto update = [&](int x, auto SCc, auto SNc, auto SFc, int q) {
      constexpr int SC = decltype(SCc)::value, SN = decltype(SNc)::value, SF = decltype(SFc)::value;
      const float* w = smem + q * SLOT + smA;
      const float* st = stash + q * SS;
      const F2 BC[3] = {lds2(w), RB[SC][0], RB[SC][1]};
      const F2 EC[3] = {lds2(w + 3 * PL), RE[SC][0], RE[SC][1]};
      // lane-axis neighbours of the pair (a, b): left cell l, right cell r
      auto ddl = [&](int f, F2 own) -> F2 {
        const float l = w[f * PL - 1], r = w[f * PL + 2];
        return cd(mk2(own.y, r), mk2(l, own.x));
      };
      auto ddx = [&](const F2(&R)[3][2], int k) -> F2 { return cd(R[SN][k - 1], R[SF][k - 1]); };
      F2 curlB[3], curlE[3];
      if constexpr (ND == 3) {  // curl3D: lane axis is z, row axis is y
        auto ddy = [&](int f) -> F2 { return cd(lds2(w + f * PL + RS), lds2(w + f * PL - RS)); };
        curlB[0] = ddy(2) - ddl(1, BC[1]);
        curlB[1] = ddl(0, BC[0]) - ddx(RB, 2);
        curlB[2] = ddx(RB, 1) - ddy(0);
        curlE[0] = ddy(5) - ddl(4, EC[1]);
        curlE[1] = ddl(3, EC[0]) - ddx(RE, 2);
        curlE[2] = ddx(RE, 1) - ddy(3);
      } else {  // curl2D: lane axis is y
        curlB[0] = ddl(2, BC[2]);
        curlB[1] = -ddx(RB, 2);
        curlB[2] = ddx(RB, 1) - ddl(0, BC[0]);
        curlE[0] = ddl(5, EC[2]);
        curlE[1] = -ddx(RE, 2);
        curlE[2] = ddx(RE, 1) - ddl(3, EC[0]);
      }
      F2 cr[3];
      cr[0] = BC[1] * curlB[2] - BC[2] * curlB[1];
      cr[1] = BC[2] * curlB[0] - BC[0] * curlB[2];
      cr[2] = BC[0] * curlB[1] - BC[1] * curlB[0];
      const F2 rho = lds2(st + 6 * SP);
#if !AMX_STRICT
      const F2 ir = mk2(rdiv(1.f, rho.x), rdiv(1.f, rho.y));
#endif
      F2 dB = bc2(0.f), dv = bc2(0.f);
      const long long o = colA + (long long)x * sx;
      float* qB = fB1 + o;
      float* qv = fv1 + o;
      const F2 dt2 = bc2(dt);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
#if AMX_STRICT
        const F2 phiB = mk2(cr[k].x / rho.x, cr[k].y / rho.y);
#else
        const F2 phiB = cr[k] * ir;
#endif
        const F2 b0 = lds2(st + k * SP);
        const F2 v0 = k < ND ? lds2(st + (3 + k) * SP) : bc2(0.f);
        const F2 Bn = b0 - dt2 * curlE[k];
        const F2 vn = v0 - dt2 * phiB;
        const F2 bk = have_k ? lds2(st + (7 + k) * SP) : b0;
        const F2 vk = have_k ? lds2(st + (10 + k) * SP) : v0;
        const F2 eB = abs2(bk - Bn), ev = abs2(vk - vn);
        dB = dB + eB * eB;
        dv = dv + ev * ev;
        stg2(qB + k * vB1, Bn);
        stg2(qv + k * vv1, vn);
        if (a.peers.flags) {  // slab rank: edge planes -> the neighbours' ghost planes
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            long long bd;
            int ps;
            if (peer_target(a.peers, x, g.n[0], side, bd, ps)) {
              stg2(reinterpret_cast<float*>(reinterpret_cast<char*>(qB + k * vB1) + bd) + (long long)ps * sx, Bn);
              stg2(reinterpret_cast<float*>(reinterpret_cast<char*>(qv + k * vv1) + bd) + (long long)ps * sx, vn);
            }
          }
        }
      }
      ch0 = max2(ch0, dB);
      ch1 = max2(ch1, dv);
    };

    // ---- prologue: plane xs-1 in registers only, plane xs published; loads of xs+1 in flight
    issue(xs - 1);
    process(xs - 1, IC<0>{}, -1);
    issue(xs);
    process(xs, IC<1>{}, 0);
    issue(xs + 1);
    __syncthreads();

    // one plane: the loads in flight become plane x+1 (register set SN, ring / stash slot q^1),
    // the loads of plane x+2 start, then plane x (set SC, slot q) is updated with set SF = plane
    // x-1.  Ring slot q^1 was last read by update(x-1), before the previous barrier.
    int q = 0;
    auto step = [&](int x, auto SCc, auto SNc, auto SFc) {
      process(x + 1, SNc, q ^ 1);
      if (x + 2 <= xe) issue(x + 2);
      if (active) update(x, SCc, SNc, SFc, q);
      __syncthreads();
      q ^= 1;
    };
    for (int x = xs;;) {
      step(x, IC<1>{}, IC<2>{}, IC<0>{});
      if (++x >= xe) break;
      step(x, IC<2>{}, IC<0>{}, IC<1>{});
      if (++x >= xe) break;
      step(x, IC<0>{}, IC<1>{}, IC<2>{});
      if (++x >= xe) break;
    }
  }

  // ---- change norms: max over cells of the squared 3-norms, sqrt (monotone), block max, slots
  float chs[2] = {rsqrt_(fmaxf(ch0.x, ch0.y)), rsqrt_(fmaxf(ch1.x, ch1.y))};
  __shared__ float red[2][kThreads / 32];
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chs[k] = fmaxf(chs[k], __shfl_xor_sync(0xffffffffu, chs[k], o));
    if (lane == 0) red[k][wid] = chs[k];
  }
  __syncthreads();
  if (tid < 2 && a.change_slots) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) m = fmaxf(m, red[tid][i]);
    const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    atomicMax(a.change_slots + (size_t)tid * kAlphaSlots + (bid % kAlphaSlots), bits_of(m));
  }
}


template <int ND, int MODE, int LD, int CW>
cudaError_t launch_mag(const MagIterArgs& a, cudaStream_t st) {
  constexpr int TY = VT<ND, CW>::TY, TZ = 2 * VT<ND, CW>::TQ, kThreads = 32 * CW + 32;
  const size_t smem = mag_vec_smem_bytes<ND, CW>();
  auto kern = mag_iter_vec<ND, MODE, LD, CW>;
  static SmemOptIn opt;
  if (opt.first_time()) {
    cudaError_t e = opt.ensure(kern, smem);
    if (e != cudaSuccess) return e;
    // Leave the carveout to the driver (just enough shared memory for two blocks): the remaining L1
    // tracks the loads in flight, and starving it costs ~45 % (measured: 0.47 ms with
    // cudaSharedmemCarveoutMaxShared, 0.32 ms without).
    static const int carve = [] { const char* e = getenv("AMX_VEC_CARVE"); return e ? atoi(e) : -1; }();
    if (carve >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  }
  dim3 grid;
  const int LA = ND - 1;
  grid.x = (a.g.n[LA] + TZ - 1) / TZ;
  const int ytiles = ND == 3 ? (a.g.n[1] + TY - 1) / TY : 1;
  const long long tiles = (long long)grid.x * ytiles;
  const int nch = pick_chunks(a.g.n[0], tiles, CW == 7 ? 2 : AMX_MAG_MINB8, 3);
  const int xchunk = (a.g.n[0] + nch - 1) / nch;
  const int nchunks = (a.g.n[0] + xchunk - 1) / xchunk;
  if (ND == 3) { grid.y = ytiles; grid.z = nchunks; }
  else { grid.y = nchunks; grid.z = 1; }
  kern<<<grid, kThreads, smem, st>>>(a, xchunk);
  return cudaGetLastError();
}

// ================================================================ unsplit hydro stage
// hydro_stage_vec<ND, RS, LIM, COMBINE>: one forward-Euler stage of the unsplit scheme
// (evolve_state.py:339-414) [+ RK2 combine :456-469] with the same thread layout as mag_iter_vec.
//   * planes are staged by cp.async (LDGSTS) straight into a 3-slot shared-memory window ring
//     (tile + 2-cell halo; halo rows / columns by the halo warp, boundary rules by index mapping),
//     so no registers are tied up by loads in flight;
//   * x direction: s(x), s(x+1), the x half-slope and the flux at x-1/2 are carried in registers;
//   * y direction: a thread reads rows y-2, y-1, y+1 of its pair from the window, computes the
//     half-slopes of rows y-1 and y and the flux at y-1/2, publishes it, and after the plane's
//     one barrier reads the flux at y+1/2 published by the row above;
//   * z direction (inside the pair and across it): the half-slopes of cells z-1, z, z+1 and the
//     fluxes at z-1/2 and z+1/2 are evaluated as packed pairs; the flux at z+3/2 is the first
//     half of the right neighbour's pair, read back after the barrier;
//   * the fluxes on the upper tile edges (top row, right column) are extra work of the halo warp.
// Every flux is computed exactly once; slopes of rows y-1 / cells z-1 are recomputed (cheap).
AMX_DI void cp_async8(float* smem_dst, const float* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
AMX_DI void cp_async4(float* smem_dst, const float* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
AMX_DI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
AMX_DI void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef AMX_HYDRO_CW
#define AMX_HYDRO_CW 7
#endif
#ifndef AMX_HYDRO_MINB
#define AMX_HYDRO_MINB 2
#endif
#ifndef AMX_TILE_GROUP
#define AMX_TILE_GROUP 1
#endif
// tile rows per group of the work-list order (hydro_stage_vec); 1 = z fastest.  Measured at 256^3 (stage 1 /
// stage 2 + combine, ms): 1: 0.338 / 0.362, 2: 0.337 / 0.371, 4: 0.340 / 0.378, 19 (y fastest): 0.339 / 0.376
constexpr int kTileGroup = AMX_TILE_GROUP;
constexpr int kHydroCW = AMX_HYDRO_CW;  // compute warps (+ 1 halo warp); AMX_HYDRO_MINB blocks per SM

template <int ND>
struct HG {
  static constexpr int NV = ND + 2;
  static constexpr int TY = ND == 3 ? 2 * kHydroCW : 1, TQ = ND == 3 ? 16 : 32 * kHydroCW, TZ = 2 * TQ;
  static constexpr int WY = ND == 3 ? TY + 4 : 1, YO = ND == 3 ? 2 : 0;
#ifndef AMX_HYDRO_RSW_PAD
#define AMX_HYDRO_RSW_PAD 8
#endif
  // window row: cells z0-2, z0-1 at columns 2, 3; own pairs from column 4.  3-D pitch 42: the halo warp
  // reads one column of consecutive rows per lane, and 42 mod 32 = 10 spreads 14 rows over 14 banks
  static constexpr int RSW = TZ + AMX_HYDRO_RSW_PAD;
  static constexpr int PLW = WY * RSW;     // floats per variable per window slot
  static constexpr int WSLOT = NV * PLW;
  static constexpr int PLFY = ND == 3 ? (TY + 1) * TZ : 0;  // y fluxes: lower interfaces of rows 0 .. TY
  static constexpr int PLFZ = TY * (TQ + 1);                // first halves of the lane-axis flux pairs + edge
  static constexpr int FSLOT = NV * (PLFY + PLFZ);
  // 3-D: every limited slope is evaluated once.  A thread publishes the upper y face state of its pair
  // (s + d_y, the L state of the interface above) one plane ahead: rows -1 .. TY-1 at row index 0 .. TY;
  // the upper z face state of the last cell left of a tile row (L state of the row's first interface)
  // comes from the halo warp.  2-D keeps one row of 7 warps, where a shuffle cannot cross the warps.
  static constexpr bool SHARE = ND == 3;
  static constexpr int PLPY = SHARE ? (TY + 1) * TZ : 0;   // floats per variable per plus slot
  static constexpr int PSLOT = NV * PLPY;
  static constexpr int ESLOT = SHARE ? NV * 16 : 0;        // left-edge L states: [NV][row]
  static constexpr size_t smem_bytes = (size_t)(3 * WSLOT + 2 * FSLOT + 2 * PSLOT + 2 * ESLOT) * sizeof(float);
};

template <int ND, int RS, int LIM, bool COMBINE>
__global__ void __launch_bounds__(32 * kHydroCW + 32, AMX_HYDRO_MINB) hydro_stage_vec(const StageArgs a, const int xchunk, const int work) {
  if (a.ctl_ro && a.ctl_ro->done) return;
  using G = HG<ND>;
  constexpr int NV = G::NV, TY = G::TY, TQ = G::TQ, TZ = G::TZ, YO = G::YO, RSW = G::RSW, PLW = G::PLW;
  constexpr int WSLOT = G::WSLOT, PLFY = G::PLFY, PLFZ = G::PLFZ, FSLOT = G::FSLOT;
  constexpr bool SHARE = G::SHARE;
  constexpr int PLPY = G::PLPY, PSLOT = G::PSLOT, ESLOT = G::ESLOT;
  constexpr int LA = ND - 1;  // lane (contiguous) axis
  constexpr int kCompute = 32 * kHydroCW;
  extern __shared__ __align__(16) float smem[];
  float* const win = smem;                 // [3][NV][WY][RSW]
  float* const flx = smem + 3 * WSLOT;     // [2][NV][PLFY + PLFZ]
  float* const ply = flx + 2 * FSLOT;      // [2][NV][TY + 1][TZ]  upper y face states, slot = plane parity
  float* const edl = ply + 2 * PSLOT;      // [2][NV][16]          upper z face state of cell z0-1 per row
  // Split barriers (one arrival per warp): barF "the fluxes of plane x are published" (arrive after the
  // y / z fluxes, wait before the update; the x flux and the next plane's y face states run in between),
  // barP "plane x+2 has landed, the face states of plane x+1 are published" (arrive after them, wait at the
  // start of the next plane; the update and the stores run in between).  A warp never runs ahead by more
  // than the work between an arrive and its wait, and nobody waits for the slowest warp in lockstep.
  __shared__ unsigned long long bars[2];
  unsigned long long* const barF = &bars[0];
  unsigned long long* const barP = &bars[1];

  const Consts<F2> c = make_consts2(a.np);
  const Geom& g = a.g;
  const int lim = a.np.lim;
  const int tid = threadIdx.x;
  const bool compute = tid < kCompute;
  if (tid == 0) {
    mbar_init(barF, kHydroCW + 1);
    mbar_init(barP, kHydroCW + 1);
  }
  const float dts = *static_cast<const float*>(a.dt) * (float)a.dt_scale;
  // common strides (vec_stage_supported); element offsets are 32-bit (the same check): one IMAD.WIDE per access
  const int sx = (int)a.in.s[0], sy = ND == 3 ? (int)a.in.s[1] : 0;
  const float* const fin = static_cast<const float*>(a.in.p);
  const int vin = (int)a.in.vs;
  F2 alpha_lf[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) alpha_lf[k] = RS == RS_LF ? F2(static_cast<const float*>(a.alpha_lf)[k]) : F2(0.f);

  // (the ring is filled one plane further ahead than a segment needs: on a slab rank (BC_HALO) that plane would lie
  // beyond the two resident ghost planes, possibly beyond the allocation, so it is clamped to the last ghost plane)
  auto plane_map = [&](int p, bool& fx) {
    return min(max(map_index(p, g.n[0], g.bc[0][0], g.bc[0][1], fx), -2), g.n[0] + 1);
  };
  // conservative update of one variable: U + ((1/dx) * (Fl - Fr)) * dt (evolve_state.py:398-403)
#if AMX_STRICT
  const F2 dt2 = F2(dts);
  auto upd = [&](F2 U, F2 fl, F2 fr) -> F2 { return U + (c.inv_dx * (fl - fr)) * dt2; };
#else
  const F2 dtdx = c.inv_dx * F2(dts);
  auto upd = [&](F2 U, F2 fl, F2 fr) -> F2 { return U + dtdx * (fl - fr); };
#endif
  // flux at the interface between cells m1 | s0 from the four cells m2, m1, s0, p1 along `ax`
  auto edge_flux = [&](const F2 (&m2)[NV], const F2 (&m1)[NV], const F2 (&s0)[NV], const F2 (&p1)[NV], int ax,
                       F2 alpha) -> Gas<F2, ND> {
    Gas<F2, ND> L, R;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const F2 d0 = s0[k] - m1[k];
      L.q[k] = face_hi<LIM>(m1[k], slope_q<LIM>(m1[k] - m2[k], d0, lim, c));
      R.q[k] = face_lo<LIM>(s0[k], slope_q<LIM>(d0, p1[k] - s0[k], lim, c));
    }
    return riemann<F2, ND, RS>(L, R, ax, alpha, c);
  };

  float wbx = 0.f, wby = 0.f, wbz = 0.f;  // running max of |u_axis| + c (axes x, y, z; 2-D: x, y in wbx, wby)

  // Work list: the plane-tiles of the launch in (x chunk, tile, plane) order, cut into equal shares, one per
  // block (launch_stage_vec: as many blocks as the GPU holds at once), so that no SM idles behind a partial
  // last wave.  A share is a few segments [xs, xe) of consecutive tiles; neighbouring tiles stay a few planes
  // apart in x, so their halo rows meet in L2.
  const int ztiles = (g.n[LA] + TZ - 1) / TZ;
  const int ytiles = ND == 3 ? (g.n[1] + TY - 1) / TY : 1;
  const int ntiles = ztiles * ytiles;
  const int nchunks = (g.n[0] + xchunk - 1) / xchunk;
  const int u_end = min((int)(blockIdx.x + 1) * work, nchunks * ntiles * xchunk);  // (launcher: fits an int)
  unsigned pb = 0;  // parity of the mbarrier phases completed by earlier segments
  for (int u = blockIdx.x * work; u < u_end;) {
  const int tl = u / xchunk, xoff = u - tl * xchunk;
  const int chunk = tl / ntiles, tile = tl - chunk * ntiles;
  // tile order: groups of kTileGroup tile rows, inside a group y fastest -- neighbouring tiles in y sit next to
  // each other in the list (a few planes apart in x), neighbours in z kTileGroup entries apart
  int tzi, tyi;
  if constexpr (ND == 3) {
    const int per = kTileGroup * ztiles, grp = tile / per, rem = tile - grp * per;
    const int rows = min(kTileGroup, ytiles - grp * kTileGroup);
    tzi = rem / rows;
    tyi = grp * kTileGroup + (rem - tzi * rows);
  } else {
    tzi = tile; tyi = 0;
  }
  const int z0 = tzi * TZ;
  const int y0 = tyi * TY;
  const int xs = chunk * xchunk + xoff;
  const int ce = min(chunk * xchunk + xchunk, g.n[0]);
  const int xe = min(ce, xs + (u_end - u));
  u = xe >= ce ? (tl + 1) * xchunk : u_end;
  if (xs >= xe) continue;
  const int ta = ND == 3 ? min(TY, g.n[1] - y0) : 1;   // active rows / pairs of a ragged tile
  const int tqa = min(TQ, (g.n[LA] - z0) >> 1);

  if (!compute) {
    // ================= halo warp
    const int l = tid - kCompute;
    // items: A0 / A1 = y-halo pairs (3-D), B = lane-axis halo pair of one row (two mapped cells)
    int colA0 = 0, colA1 = 0, colB0 = 0, colB1 = 0;
    bool haveA = false, haveB = false, fA0 = false, fA1 = false, fB0 = false, fB1 = false;
    int smA0 = 0, smA1 = 0, smB = 0;
    if constexpr (ND == 3) {
      const int q = l % TQ, jr = l / TQ;  // jr = 0, 1
      haveA = q < tqa;
      const int my0 = map_index(y0 - 2 + jr, g.n[1], g.bc[1][0], g.bc[1][1], fA0);
      const int my1 = map_index(y0 + ta + jr, g.n[1], g.bc[1][0], g.bc[1][1], fA1);
      colA0 = z0 + 2 * q + my0 * sy;
      colA1 = z0 + 2 * q + my1 * sy;
      smA0 = jr * RSW + 4 + 2 * q;
      smA1 = (ta + 2 + jr) * RSW + 4 + 2 * q;
      const int r = l % 16, side = l / 16;
      haveB = r < ta;
      const int zc = side == 0 ? z0 - 2 : z0 + 2 * tqa;
      const int mz0 = map_index(zc, g.n[2], g.bc[2][0], g.bc[2][1], fB0);
      const int mz1 = map_index(zc + 1, g.n[2], g.bc[2][0], g.bc[2][1], fB1);
      colB0 = mz0 + (y0 + r) * sy;
      colB1 = mz1 + (y0 + r) * sy;
      smB = (r + YO) * RSW + (side == 0 ? 2 : 4 + 2 * tqa);
    } else {
      haveB = l < 2;
      const int zc = l == 0 ? z0 - 2 : z0 + 2 * tqa;
      const int mz0 = map_index(zc, g.n[1], g.bc[1][0], g.bc[1][1], fB0);
      const int mz1 = map_index(zc + 1, g.n[1], g.bc[1][0], g.bc[1][1], fB1);
      colB0 = mz0;
      colB1 = mz1;
      smB = l == 0 ? 2 : 4 + 2 * tqa;
    }
    auto issue = [&](int p, int slot) {
      bool fx;
      const int o = plane_map(p, fx) * sx;
      float* w = win + slot * WSLOT;
      if (haveA) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          cp_async8(w + k * PLW + smA0, fin + (k * vin + o + colA0));
          cp_async8(w + k * PLW + smA1, fin + (k * vin + o + colA1));
        }
      }
      if (haveB) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          cp_async4(w + k * PLW + smB, fin + (k * vin + o + colB0));
          cp_async4(w + k * PLW + smB + 1, fin + (k * vin + o + colB1));
        }
      }
      cp_async_commit();
    };
    // sign flips of reflective sides on the landed halo entries (own entries only)
    auto fixup = [&](int p, int slot) {
      bool fx;
      (void)plane_map(p, fx);
      float* w = win + slot * WSLOT;
      if (haveA) {
        if (fx) { w[PLW + smA0] = -w[PLW + smA0]; w[PLW + smA0 + 1] = -w[PLW + smA0 + 1];
                  w[PLW + smA1] = -w[PLW + smA1]; w[PLW + smA1 + 1] = -w[PLW + smA1 + 1]; }
        if (fA0) { w[2 * PLW + smA0] = -w[2 * PLW + smA0]; w[2 * PLW + smA0 + 1] = -w[2 * PLW + smA0 + 1]; }
        if (fA1) { w[2 * PLW + smA1] = -w[2 * PLW + smA1]; w[2 * PLW + smA1 + 1] = -w[2 * PLW + smA1 + 1]; }
      }
      if (haveB) {
        if (fx) { w[PLW + smB] = -w[PLW + smB]; w[PLW + smB + 1] = -w[PLW + smB + 1]; }
        if (fB0) w[(LA + 1) * PLW + smB] = -w[(LA + 1) * PLW + smB];
        if (fB1) w[(LA + 1) * PLW + smB + 1] = -w[(LA + 1) * PLW + smB + 1];
      }
    };
    // fluxes on the upper tile edges of plane x (window slot s complete and visible)
    auto edge_fluxes = [&](int s, int fs, int par) {
      const float* w = win + s * WSLOT;
      float* f = flx + fs * FSLOT;
      if constexpr (ND == 3) {
        if (l < TQ) {
          if (l < tqa) {  // interface below row ta (pair l)
            const float* wp = w + (ta + 2) * RSW + 4 + 2 * l;
            Gas<F2, ND> F;
            if constexpr (SHARE) {
              // L = upper face state of tile row ta-1 (published by its thread), R = lower face state of halo row ta
              const float* pl = ply + par * PSLOT + ta * TZ + 2 * l;
              Gas<F2, ND> L, R;
#pragma unroll
              for (int k = 0; k < NV; ++k) {
                const F2 m1 = lds2(wp + k * PLW - RSW), s0 = lds2(wp + k * PLW), p1 = lds2(wp + k * PLW + RSW);
                L.q[k] = lds2(pl + k * PLPY);
                R.q[k] = face_lo<LIM>(s0, slope_q<LIM>(s0 - m1, p1 - s0, lim, c));
              }
              F = riemann<F2, ND, RS>(L, R, 2, alpha_lf[1], c);
            } else {
              F2 m2[NV], m1[NV], s0[NV], p1[NV];
#pragma unroll
              for (int k = 0; k < NV; ++k) {
                m2[k] = lds2(wp + k * PLW - 2 * RSW);
                m1[k] = lds2(wp + k * PLW - RSW);
                s0[k] = lds2(wp + k * PLW);
                p1[k] = lds2(wp + k * PLW + RSW);
              }
              F = edge_flux(m2, m1, s0, p1, 2, alpha_lf[1]);
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) sts2(f + k * (PLFY + PLFZ) + ta * TZ + 2 * l, F.q[k]);
          }
          return;
        }
      }
      const int r = ND == 3 ? l - TQ : l;
      if (r < ta && (ND == 3 || l == 0)) {  // interface left of the first right-halo cell of row r
        const float* wp = w + (r + YO) * RSW + 4 + 2 * tqa;
        F2 m2[NV], m1[NV], s0[NV], p1[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          m2[k] = F2(wp[k * PLW - 2]);
          m1[k] = F2(wp[k * PLW - 1]);
          s0[k] = F2(wp[k * PLW]);
          p1[k] = F2(wp[k * PLW + 1]);
        }
        const Gas<F2, ND> F = edge_flux(m2, m1, s0, p1, LA + 1, alpha_lf[LA]);
#pragma unroll
        for (int k = 0; k < NV; ++k) f[k * (PLFY + PLFZ) + PLFY + r * (TQ + 1) + tqa] = F.q[k].x;
      }
    };
    // SHARE: face states that compute threads cannot get from a neighbour thread, from the complete window
    // slot s into parity slot q: the upper y face state of halo row y0-1 (pair l) and the upper z face
    // state of cell z0-1 of tile row r (the L states of the tile's lowest y / first z interfaces)
    auto halo_plus = [&](int s, int q) {
      if constexpr (SHARE) {
        const float* w = win + s * WSLOT;
        if (l < TQ) {
          if (l < tqa) {
            const float* wp = w + RSW + 4 + 2 * l;  // window row 1 = y0-1
            float* pp = ply + q * PSLOT + 2 * l;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const F2 m1 = lds2(wp + k * PLW - RSW), s0 = lds2(wp + k * PLW), p1 = lds2(wp + k * PLW + RSW);
              sts2(pp + k * PLPY, face_hi<LIM>(s0, slope_q<LIM>(s0 - m1, p1 - s0, lim, c)));
            }
          }
        } else if (l - TQ < ta) {
          const int r = l - TQ;
          const float* wp = w + (r + YO) * RSW + 3;  // cell z0-1
          float* pe = edl + q * ESLOT + r;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const F2 m1 = F2(wp[k * PLW - 1]), s0 = F2(wp[k * PLW]), p1 = F2(wp[k * PLW + 1]);
            pe[k * 16] = face_hi<LIM>(s0, slope_q<LIM>(s0 - m1, p1 - s0, lim, c)).x;
          }
        }
      }
    };
    issue(xs, 1);
    issue(xs + 1, 2);
    issue(xs + 2, 0);
    cp_async_wait<1>();
    fixup(xs, 1);
    fixup(xs + 1, 2);
    __syncthreads();
    if constexpr (SHARE) {
      halo_plus(1, 0);
      __syncthreads();
    }
    int sC = 1, sN = 2, sF = 0, fs = 0;
    for (int x = xs; x < xe; ++x) {
      if (x > xs) mbar_wait(barP, fs ^ 1 ^ pb);
      edge_fluxes(sC, fs, fs);
      mbar_arrive_warp(barF);
      halo_plus(sN, fs ^ 1);
      cp_async_wait<0>();
      fixup(x + 2, sF);
      mbar_arrive_warp(barP);
      mbar_wait(barF, fs ^ pb);
      issue(x + 3, sC);
      const int t = sC; sC = sN; sN = sF; sF = t;
      fs ^= 1;
    }
    cp_async_wait<0>();
  } else {
    // ================= compute threads: pair (ty, tq)
    const int tq = tid % TQ, ty = tid / TQ;
    const bool active = tq < tqa && ty < ta;
    const int col = z0 + 2 * tq + (ND == 3 ? (y0 + ty) * sy : 0);
    const int smW = (ty + YO) * RSW + 4 + 2 * tq;  // own pair inside a window plane
    float* const fout = static_cast<float*>(a.out.p);
    const float* const fin0 = static_cast<const float*>(a.in0.p);
    const int vout = (int)a.out.vs, vin0 = (int)a.in0.vs;

    auto issue = [&](int p, int slot) {
      if (active) {
        bool fx;
        const int src = plane_map(p, fx) * sx + col;
        float* w = win + slot * WSLOT + smW;
#pragma unroll
        for (int k = 0; k < NV; ++k) cp_async8(w + k * PLW, fin + (src + k * vin));
      }
      cp_async_commit();
    };
    auto fixup = [&](int p, int slot) {  // reflective x side: v_x of the mirrored plane changes sign
      bool fx;
      (void)plane_map(p, fx);
      if (fx && active) {
        float* w = win + slot * WSLOT + smW + PLW;
        w[0] = -w[0];
        w[1] = -w[1];
      }
    };
    auto load_direct = [&](int p, F2 (&v)[NV]) {
      bool fx;
      const int src = plane_map(p, fx) * sx + col;
#pragma unroll
      for (int k = 0; k < NV; ++k) v[k] = active ? ldg2<0>(fin + (src + k * vin)) : F2(1.f);
      if (fx) v[1] = -v[1];
    };
    auto read_own = [&](int slot, F2 (&v)[NV]) {
      const float* w = win + slot * WSLOT + smW;
#pragma unroll
      for (int k = 0; k < NV; ++k) v[k] = lds2(w + k * PLW);  // (rows / pairs beyond a ragged tile: never stored)
    };

    // ---- prologue: planes xs-2, xs-1 in registers; xs, xs+1 in the ring; xs+2 in flight.
    // Across steps a thread carries only the x half-slope of the centre plane and the flux at
    // x-1/2 (ping-pong register sets, the loop is unrolled by two); the centre / next planes of
    // its own pair are re-read from the window ring.
    F2 d0[NV], F0[NV], d1[NV], F1[NV];
    issue(xs, 1);
    issue(xs + 1, 2);
    issue(xs + 2, 0);
    {
      F2 sM2[NV], sM1[NV], sC[NV], sP[NV];
      load_direct(xs - 2, sM2);
      load_direct(xs - 1, sM1);
      cp_async_wait<1>();
      fixup(xs, 1);
      fixup(xs + 1, 2);
      read_own(1, sC);
      read_own(2, sP);
      Gas<F2, ND> L, R;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const F2 dd = sC[k] - sM1[k];
        d0[k] = slope_q<LIM>(dd, sP[k] - sC[k], lim, c);
        L.q[k] = face_hi<LIM>(sM1[k], slope_q<LIM>(sM1[k] - sM2[k], dd, lim, c));
        R.q[k] = face_lo<LIM>(sC[k], d0[k]);
      }
      const Gas<F2, ND> F = riemann<F2, ND, RS>(L, R, 1, alpha_lf[0], c);
#pragma unroll
      for (int k = 0; k < NV; ++k) F0[k] = F.q[k];
    }
    __syncthreads();
    // SHARE: lower / upper y face states of the own pair in plane p from the complete window slot s; the
    // upper one is published for the row above (plus ring, parity slot q), the lower one stays here
    auto y_faces = [&](int s, int q, const F2 (&sv)[NV], F2 (&lower)[NV]) {
      const float* wn = win + s * WSLOT + smW;
      float* pp = ply + q * PSLOT + (ty + 1) * TZ + 2 * tq;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const F2 m1 = lds2(wn + k * PLW - RSW), p1 = lds2(wn + k * PLW + RSW);
        const F2 qy = slope_q<LIM>(sv[k] - m1, p1 - sv[k], lim, c);
        if (active) sts2(pp + k * PLPY, face_hi<LIM>(sv[k], qy));
        lower[k] = face_lo<LIM>(sv[k], qy);
      }
    };
    F2 mY0[NV], mY1[NV];
    if constexpr (SHARE) {
      F2 sv[NV];
      read_own(1, sv);
      y_faces(1, 0, sv, mY0);
      __syncthreads();
    }

    int sCs = 1, sNs = 2, sFs = 0, fs = 0;
    // one plane: (dC, FxL, mYc) in, (dP, FxR, mYn) out
    auto step = [&](int x, const F2 (&dC)[NV], const F2 (&FxL)[NV], const F2 (&mYc)[NV], F2 (&dP)[NV],
                    F2 (&FxR)[NV], F2 (&mYn)[NV]) {
      const float* w = win + sCs * WSLOT + smW;
      float* f = flx + fs * FSLOT;
      if (x > xs) mbar_wait(barP, fs ^ 1 ^ pb);  // window of plane x+1 complete, face states of plane x published
      F2 sC[NV];
      read_own(sCs, sC);
      // ---- phase 1a: flux at the lower y interface
      if constexpr (ND == 3) {
        Gas<F2, ND> L, R;
        if constexpr (SHARE) {  // L: upper face state of the row below (plus ring), R: own lower face state
          const float* pl = ply + fs * PSLOT + ty * TZ + 2 * tq;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            L.q[k] = lds2(pl + k * PLPY);
            R.q[k] = mYc[k];
          }
        } else {  // rows y-2, y-1, y, y+1 of the window
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const F2 m2 = lds2(w + k * PLW - 2 * RSW), m1 = lds2(w + k * PLW - RSW), p1 = lds2(w + k * PLW + RSW);
            const F2 dd = sC[k] - m1;
            L.q[k] = face_hi<LIM>(m1, slope_q<LIM>(m1 - m2, dd, lim, c));
            R.q[k] = face_lo<LIM>(sC[k], slope_q<LIM>(dd, p1 - sC[k], lim, c));
          }
        }
        const Gas<F2, ND> F = riemann<F2, ND, RS>(L, R, 2, alpha_lf[1], c);
        if (active) {
#pragma unroll
          for (int k = 0; k < NV; ++k) sts2(f + k * (PLFY + PLFZ) + ty * TZ + 2 * tq, F.q[k]);
        }
      }
      // ---- phase 1b: fluxes at the two lower lane-axis interfaces of the pair (a, b):
      // cells l2 l1 | a b | r1;  L = (l1, a) + (d_l1, d_a),  R = (a, b) - (d_a, d_b)
      F2 Fz[NV];
      {
        Gas<F2, ND> L, R;
        if constexpr (SHARE) {
          // (d_a, d_b) once; l1 + d_l1 is the .y half of the left neighbour's upper face pair (shuffle
          // inside the 16 pairs of a tile row; for the first pair it comes from the halo warp)
          const float* pe = edl + fs * ESLOT + ty;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const float l1 = w[k * PLW - 1], r1 = w[k * PLW + 2];
            const F2 D2 = sC[k] - mk2(l1, sC[k].x), D3 = mk2(sC[k].y, r1) - sC[k];
            const F2 qz = slope_q<LIM>(D2, D3, lim, c);
            const F2 pz = face_hi<LIM>(sC[k], qz);
            R.q[k] = face_lo<LIM>(sC[k], qz);
            float nb = __shfl_up_sync(0xffffffffu, pz.y, 1, 16);
            if (tq == 0) nb = pe[k * 16];
            L.q[k] = mk2(nb, pz.x);
          }
        } else {
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const F2 lp = lds2(w + k * PLW - 2);
            const float r1 = w[k * PLW + 2];
            const F2 la = mk2(lp.y, sC[k].x);
            const F2 D1 = la - lp, D2 = sC[k] - la, D3 = mk2(sC[k].y, r1) - sC[k];
            L.q[k] = face_hi<LIM>(la, slope_q<LIM>(D1, D2, lim, c));
            R.q[k] = face_lo<LIM>(sC[k], slope_q<LIM>(D2, D3, lim, c));
          }
        }
        const Gas<F2, ND> F = riemann<F2, ND, RS>(L, R, LA + 1, alpha_lf[LA], c);
#pragma unroll
        for (int k = 0; k < NV; ++k) Fz[k] = F.q[k];
        if (active) {
#pragma unroll
          for (int k = 0; k < NV; ++k) f[k * (PLFY + PLFZ) + PLFY + ty * (TQ + 1) + tq] = F.q[k].x;
        }
      }
      mbar_arrive_warp(barF);  // y / z fluxes of plane x published
      // ---- phase 1c: x direction; plane x+2 has landed in ring slot sFs
      cp_async_wait<0>();
      fixup(x + 2, sFs);
      {
        F2 sP[NV], sN[NV];
        read_own(sNs, sP);
        read_own(sFs, sN);
        Gas<F2, ND> L, R;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          dP[k] = slope_q<LIM>(sP[k] - sC[k], sN[k] - sP[k], lim, c);
          L.q[k] = face_hi<LIM>(sC[k], dC[k]);
          R.q[k] = face_lo<LIM>(sP[k], dP[k]);
        }
        const Gas<F2, ND> F = riemann<F2, ND, RS>(L, R, 1, alpha_lf[0], c);
#pragma unroll
        for (int k = 0; k < NV; ++k) FxR[k] = F.q[k];
        // y face states of plane x+1 (its window slot has been complete since the previous barrier)
        if constexpr (SHARE) y_faces(sNs, fs ^ 1, sP, mYn);
      }
      // RK2 combine input (centre pair of s0), consumed after the barrier
      F2 s0v[NV];
      if constexpr (COMBINE) {
        const int p0 = x * sx + col;
#pragma unroll
        for (int k = 0; k < NV; ++k) s0v[k] = active ? ldg2<0>(fin0 + (p0 + k * vin0)) : F2(1.f);
      }
      mbar_arrive_warp(barP);  // own part of plane x+2 landed, y face states of plane x+1 published
      mbar_wait(barF, fs ^ pb);     // fluxes of plane x published by everybody; nobody reads window slot sCs any more
      issue(x + 3, sCs);

      // ---- phase 2: conservative update axis by axis like the reference, prim (+ RK2 combine)
      {
        Gas<F2, ND> sc;
#pragma unroll
        for (int k = 0; k < NV; ++k) sc.q[k] = sC[k];
        Gas<F2, ND> U = cons_from_prim(sc, c);
#pragma unroll
        for (int k = 0; k < NV; ++k) U.q[k] = upd(U.q[k], FxL[k], FxR[k]);
        if constexpr (ND == 3) {
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const float* fy = f + k * (PLFY + PLFZ) + ty * TZ + 2 * tq;
            U.q[k] = upd(U.q[k], lds2(fy), lds2(fy + TZ));
          }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const float fr = f[k * (PLFY + PLFZ) + PLFY + ty * (TQ + 1) + tq + 1];
          U.q[k] = upd(U.q[k], Fz[k], mk2(Fz[k].y, fr));
        }
        Gas<F2, ND> s = prim_from_cons(U, c);
        if (active) {
          if constexpr (COMBINE) {
            Gas<F2, ND> s0;
#pragma unroll
            for (int k = 0; k < NV; ++k) s0.q[k] = s0v[k];
            const Gas<F2, ND> U0 = cons_from_prim(s0, c), U2 = cons_from_prim(s, c);
            Gas<F2, ND> Um;
#pragma unroll
            for (int k = 0; k < NV; ++k) Um.q[k] = F2(0.5f) * (U0.q[k] + U2.q[k]);
            s = prim_from_cons(Um, c);
          }
          const int po32 = x * sx + col;
          float* po = fout + po32;
#pragma unroll
          for (int k = 0; k < NV; ++k) stg2(fout + (po32 + k * vout), s.q[k]);
          if (a.peers.flags) {  // slab rank: the edge planes also go into the neighbours' ghost planes (NVLink stores)
#pragma unroll
            for (int side = 0; side < 2; ++side) {
              long long bd;
              int ps;
              if (peer_target(a.peers, x, g.n[0], side, bd, ps)) {
                float* pp = reinterpret_cast<float*>(reinterpret_cast<char*>(po) + bd) + ps * sx;
#pragma unroll
                for (int k = 0; k < NV; ++k) stg2(pp + k * vout, s.q[k]);
              }
            }
          }
          if (a.alpha_slots) {
            const F2 bx = wave_bound(s, 1, c), by = wave_bound(s, 2, c);
            wbx = fmaxf(wbx, fmaxf(bx.x, bx.y));
            wby = fmaxf(wby, fmaxf(by.x, by.y));
            if constexpr (ND == 3) {
              const F2 bz = wave_bound(s, 3, c);
              wbz = fmaxf(wbz, fmaxf(bz.x, bz.y));
            }
          }
        }
      }
      const int t = sCs; sCs = sNs; sNs = sFs; sFs = t;
      fs ^= 1;
    };
    for (int x = xs;;) {
      step(x, d0, F0, mY0, d1, F1, mY1);
      if (++x >= xe) break;
      step(x, d1, F1, mY1, d0, F0, mY0);
      if (++x >= xe) break;
    }
    cp_async_wait<0>();
  }
  __syncthreads();  // the rings are free for the next segment
  pb ^= (unsigned)(xe - xs) & 1u;
  }  // segments

  if (a.alpha_slots) {
    float wb[3] = {wbx, wby, wbz};
    __shared__ float red[3][8];
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wb[k] = fmaxf(wb[k], __shfl_xor_sync(0xffffffffu, wb[k], o));
      if (lane == 0) red[k][wid] = wb[k];
    }
    __syncthreads();
    if (tid < ND) {
      float m = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) m = fmaxf(m, red[tid][i]);
      const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
      atomicMax(a.alpha_slots + (size_t)tid * kAlphaSlots + (bid % kAlphaSlots), bits_of(m));
    }
  }
}

template <int ND, int RS, int LIM, bool COMBINE>
cudaError_t launch_stage_vec(const StageArgs& a, cudaStream_t st) {
  using G = HG<ND>;
  auto kern = hydro_stage_vec<ND, RS, LIM, COMBINE>;
  static SmemOptIn opt;
  {
    cudaError_t e = opt.ensure(kern, G::smem_bytes);
    if (e != cudaSuccess) return e;
  }
  const int LA = ND - 1;
  const long long tiles = (long long)((a.g.n[LA] + G::TZ - 1) / G::TZ) * (ND == 3 ? (a.g.n[1] + G::TY - 1) / G::TY : 1);
  int xchunk, work;
  const int nblocks = pick_shares(a.g.n[0], tiles, AMX_HYDRO_MINB, xchunk, work);
  const dim3 grid(nblocks);
  kern<<<grid, 32 * kHydroCW + 32, G::smem_bytes, st>>>(a, xchunk, work);
  return cudaGetLastError();
}


}  // namespace vec

// fp32, 2-D / 3-D, contiguous 8-byte-aligned rows, even extent along the contiguous axis
bool vec_mag_supported(const MagIterArgs& a) {
  if (a.g.jet_hi > a.g.jet_lo) return false;  // MHD jet inflow cells: generic kernels only
  if (a.np.dtype != 0) return false;
  const int nd = a.g.nd;
  if (nd != 2 && nd != 3) return false;
  if (!vec::view_ok(a.B0, nd) || !vec::view_ok(a.gas0, nd) || !vec::view_ok(a.Bk1, nd) || !vec::view_ok(a.vk1, nd)) return false;
  if (a.Bk.p && (!vec::view_ok(a.Bk, nd) || !vec::view_ok(a.vk, nd))) return false;
  for (int ax = 0; ax < nd - 1; ++ax) {
    const long long s0 = a.B0.s[ax];
    if (a.gas0.s[ax] != s0 || a.Bk1.s[ax] != s0 || a.vk1.s[ax] != s0) return false;
    if (a.Bk.p && (a.Bk.s[ax] != s0 || a.vk.s[ax] != s0)) return false;
  }
  if (a.g.n[0] < 4 || (a.g.n[nd - 1] & 1) || a.g.n[nd - 1] < 8) return false;
  return nd == 2 || a.g.n[1] >= 4;
}

template <int ND, int LD, int CW>
static cudaError_t launch_mag_ld(const MagIterArgs& a, cudaStream_t st) {
  if (!a.Bk.p) return vec::launch_mag<ND, 0, LD, CW>(a, st);
  return a.midpoint ? vec::launch_mag<ND, 1, LD, CW>(a, st) : vec::launch_mag<ND, 2, LD, CW>(a, st);
}
template <int ND>
static cudaError_t launch_mag_mode(const MagIterArgs& a, cudaStream_t st) {
  static const int cw = [] { const char* e = getenv("AMX_VEC_CW"); return e ? atoi(e) : 7; }();
  if (cw == 8) return launch_mag_ld<ND, 0, 8>(a, st);
  return launch_mag_ld<ND, 0, 7>(a, st);
}

cudaError_t launch_mag_iter_vec(const MagIterArgs& a, cudaStream_t st) {
  if (a.g.nd == 3) return launch_mag_mode<3>(a, st);
  if (a.g.nd == 2) return launch_mag_mode<2>(a, st);
  return cudaErrorInvalidValue;
}

// fp32, 2-D / 3-D, in / out / in0 with the same 8-byte-aligned contiguous layout, even lane extent
bool vec_stage_supported(const StageArgs& a) {
  if (a.g.jet_hi > a.g.jet_lo) return false;  // MHD jet inflow cells: generic kernels only
  if (a.np.lim == LIM_VAN_ALBADA_PP) return false;  // multi-axis limiter: generic kernel (amx_kernels.cu)
  if (a.np.dtype != 0) return false;
  const int nd = a.g.nd;
  if (nd != 2 && nd != 3) return false;
  if (!vec::view_ok(a.in, nd) || !vec::view_ok(a.out, nd)) return false;
  if (a.in0.p && !vec::view_ok(a.in0, nd)) return false;
  for (int ax = 0; ax < nd - 1; ++ax) {
    if (a.out.s[ax] != a.in.s[ax]) return false;
    if (a.in0.p && a.in0.s[ax] != a.in.s[ax]) return false;
  }
  if (a.g.n[0] < 4 || (a.g.n[nd - 1] & 1) || a.g.n[nd - 1] < 8) return false;
  if (nd == 3 && a.g.n[1] < 4) return false;
  if (a.np.rs != RS_HLL && a.np.rs != RS_HLLC && a.np.rs != RS_LF && a.np.rs != RS_HLLC_LM) return false;
  {  // 32-bit element offsets inside the kernel: rows * row stride (+ look-ahead planes) of every array below 2^31
    auto fits = [&](const View& v) {
      const long long ext = (long long)(nd + 1) * v.vs + (long long)(a.g.n[0] + 4) * v.s[0];
      return v.vs >= 0 && ext < 2147483647LL;
    };
    if (!fits(a.in) || !fits(a.out) || (a.in0.p && !fits(a.in0))) return false;
  }
  if (a.np.rs == RS_LF && !a.alpha_lf) return false;
  return true;
}

template <int ND, int RS, int LIM>
static cudaError_t launch_stage_vec_c(const StageArgs& a, cudaStream_t st) {
  return a.in0.p ? vec::launch_stage_vec<ND, RS, LIM, true>(a, st) : vec::launch_stage_vec<ND, RS, LIM, false>(a, st);
}
template <int ND, int RS>
static cudaError_t launch_stage_vec_l(const StageArgs& a, cudaStream_t st) {
  return a.np.lim == LIM_MINMOD ? launch_stage_vec_c<ND, RS, LIM_MINMOD>(a, st) : launch_stage_vec_c<ND, RS, -1>(a, st);
}

cudaError_t launch_hydro_stage_vec(const StageArgs& a, cudaStream_t st) {
  // (Lax-Friedrichs: the flux of the reference's published fv_mhd benchmark, arena/fv_mhd.py:36-49; the global
  // alpha of each axis comes in through StageArgs::alpha_lf)
  if (a.g.nd == 3) {
    if (a.np.rs == RS_LF) return launch_stage_vec_l<3, RS_LF>(a, st);
    if (a.np.rs == RS_HLLC_LM) return launch_stage_vec_l<3, RS_HLLC_LM>(a, st);
    return a.np.rs == RS_HLL ? launch_stage_vec_l<3, RS_HLL>(a, st) : launch_stage_vec_l<3, RS_HLLC>(a, st);
  }
  if (a.g.nd == 2) {
    if (a.np.rs == RS_LF) return launch_stage_vec_l<2, RS_LF>(a, st);
    if (a.np.rs == RS_HLLC_LM) return launch_stage_vec_l<2, RS_HLLC_LM>(a, st);
    return a.np.rs == RS_HLL ? launch_stage_vec_l<2, RS_HLL>(a, st) : launch_stage_vec_l<2, RS_HLLC>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace AMX_NS
