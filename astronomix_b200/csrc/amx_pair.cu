// amx_pair.cu -- TWO consecutive evaluations of the Pang-Wu fixed point in one launch (fp32, 3-D).
//
//   R_{k+1} = R_0 - dt * Psi(arg(R_k,     R_0))        (_magnetic_field_update.py:82-165)
//   R_{k+2} = R_0 - dt * Psi(arg(R_{k+1}, R_0))
//
// The fixed point is memory-bound: one evaluation streams R_0, rho, R_k (13 words per cell) and
// writes R_{k+1} (6 words).  Fusing two evaluations keeps R_{k+1} on chip: 19 words per TWO
// evaluations instead of 38 (temporal blocking).  The intermediate iterate is never
// written; both change norms max|R_k - R_{k+1}| and max|R_{k+1} - R_{k+2}| are produced, so the
// reference's while-condition (:127-141) is evaluated after every evaluation exactly as before.
// If the loop turns out to end on the FIRST evaluation of a pair, amx_api.cu re-runs that single
// evaluation from R_k (still intact in the other iterate buffer) with mag_iter_vec.
//
// Layout: a block owns a TY (y) x 64 (z) tile of columns and marches along x.  One warp per
// window row (32 pairs of z-adjacent cells per row), the window being the tile plus a 2-cell
// frame: rows y0-2 .. y0+TY+1, cells z0-2 .. z0+65.
//   warps 0 .. TY-1    tile rows            load, argument 1, evaluation 1, evaluation 2
//   warps TY, TY+1     rows y0-1, y0+TY     load, argument 1, evaluation 1 (frame of evaluation 2)
//   warps TY+2, TY+3   rows y0-2, y0+TY+1   load, argument 1
//   warp  TY+4         the two side pairs (z0-2, z0-1) and (z0+64, z0+65) of every window row
// Per plane step i (one __syncthreads each):
//   P1(i+2)  the loads in flight become plane i+2: argument (B, v), E = B x v -> ring A
//   E1(i+1)  Psi from ring A (planes i, i+1, i+2), R_{k+1}(i+1), norm 1, argument 2 -> ring B;
//            R_0, rho, R_{k+1} of the own pair -> thread-private stash; loads of plane i+3 start
//   E2(i)    Psi from ring B (planes i-1, i, i+1), R_{k+2}(i) -> global memory, norm 2
// Rings hold 3 planes of (B, E) (6 fields); the x-1 neighbour is the own column in another slot, the
// x+1 neighbour was computed by the thread itself a moment earlier (registers), y / z neighbours
// come from the centre slot.  y and z must be periodic (a ghost cell of an
// open / reflective side is a COPY of an interior R_{k+1}, not a cell with its own stencil);
// x is periodic or resident halo planes (slab ranks).  Anything else runs as two single launches.
#include "amx_vec_common.cuh"
#include "amx_tma.h"

#include <cstdlib>

namespace AMX_NS {
using namespace amx;

namespace vec {

// Ring layout (per slot): [group g = 0..2][window row][chunk][4 floats] with one 16-byte chunk per
// pair holding (B_g.x, E_g.x, B_g.y, E_g.y): every shared-memory access of an evaluation is a
// 128-bit one (own chunk, y / x neighbours) or a 64-bit one that returns the B AND E value of a
// z neighbour (the .y halves of the chunk on the left, the .x halves of the chunk on the right):
// 15 loads + 3 stores per evaluation instead of 30 + 6 (the kernel is bound by the number of
// LSU instructions in flight, not by their bytes).
template <int TY>
struct PG {
  static constexpr int WYA = TY + 4, WYB = TY + 2;
  static constexpr int RS = 144;  // window row: 36 chunks; window cell column c (0 .. 66, even) is chunk c/2 + 1
  static constexpr int PLA = WYA * RS, PLB = WYB * RS;
  static constexpr int SLOTA = 3 * PLA, SLOTB = 3 * PLB;
  static constexpr int NT = 32 * TY;   // tile threads
  static constexpr int SP = 4 * NT;    // floats per stashed float4 field
  static constexpr int SS = 7 * SP;    // floats per stash slot: 13 pairs packed into 7 float4
  static constexpr int kWarps = TY + 5, kThreads = 32 * kWarps;
  static constexpr size_t smem_bytes = (size_t)(3 * SLOTA + 3 * SLOTB + 2 * SS) * sizeof(float);
};

struct RawP {
  F2 B0[3], v0[3], rho, Bk[3], vk[3];
};

// Tensor maps of the four input arrays (boxes = the block's window of one plane: 72 x (TY + 4) cells x
// all rows of the array): ONE thread pulls the window of a plane that will be loaded two steps later into
// L2 with four cp.async.bulk.prefetch.tensor instructions (SASS UTMAPF) instead of every 16th lane
// prefetching its 13 lines.  on = 0: the arrays cannot be described by a map (per-lane prefetch).
struct PairMaps {
  CUtensorMap B0, gas, Bk, vk;
  int on, xpad;
};

AMX_DI float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
AMX_DI void sts4(float* p, F2 a, F2 b) { *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y); }

// Psi of the argument whose (B, E) sit in a 3-slot ring: curl E and B x curl B of the own pair.
// w / wm point at group 0 of the own chunk in the centre / x-1 slots; the x+1 neighbours of the
// own column (components 1, 2) were computed by this thread in this step and come in registers
// (_vector_maths.py:55-78).
template <int PL, int RS>
AMX_DI void psi_ring(const float* w, const float* wm, const F2 (&Bp)[3], const F2 (&Ep)[3], float dpar,
                     F2 (&curlE)[3], F2 (&cr)[3]) {
  auto cd = [&](F2 fp, F2 fm) -> F2 {
#if AMX_STRICT
    const F2 h = bc2(0.5f) * (fp - fm);
    return mk2(h.x / dpar, h.y / dpar);
#else
    return bc2(dpar) * (fp - fm);
#endif
  };
  F2 BC[3], EC[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 o = lds4(w + k * PL);
    BC[k] = mk2(o.x, o.z);
    EC[k] = mk2(o.y, o.w);
  }
  // z neighbours of groups 0, 1: (B.y, E.y) of the chunk on the left, (B.x, E.x) of the chunk on the right
  F2 dlB[2], dlE[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const F2 l = lds2(w + k * PL - 2), r = lds2(w + k * PL + 4);
    dlB[k] = cd(mk2(BC[k].y, r.x), mk2(l.x, BC[k].x));
    dlE[k] = cd(mk2(EC[k].y, r.y), mk2(l.y, EC[k].x));
  }
  // y neighbours of groups 0, 2
  F2 dyB[2], dyE[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float4 u = lds4(w + 2 * k * PL + RS), d = lds4(w + 2 * k * PL - RS);
    dyB[k] = cd(mk2(u.x, u.z), mk2(d.x, d.z));
    dyE[k] = cd(mk2(u.y, u.w), mk2(d.y, d.w));
  }
  // x neighbours of groups 1, 2
  F2 dxB[2], dxE[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float4 m = lds4(wm + (1 + k) * PL);
    dxB[k] = cd(Bp[1 + k], mk2(m.x, m.z));
    dxE[k] = cd(Ep[1 + k], mk2(m.y, m.w));
  }
  F2 curlB[3];
  curlB[0] = dyB[1] - dlB[1];
  curlB[1] = dlB[0] - dxB[1];
  curlB[2] = dxB[0] - dyB[0];
  curlE[0] = dyE[1] - dlE[1];
  curlE[1] = dlE[0] - dxE[1];
  curlE[2] = dxE[0] - dyE[0];
  cr[0] = BC[1] * curlB[2] - BC[2] * curlB[1];
  cr[1] = BC[2] * curlB[0] - BC[0] * curlB[2];
  cr[2] = BC[0] * curlB[1] - BC[1] * curlB[0];
}

// FIRST: evaluation 1 starts from R_k = R_0 (launch 0 of the fixed point).  MID: implicit midpoint
// (argument (R + R_0)/2) or implicit Euler (argument R).
template <bool FIRST, bool MID, int TY, bool EMIT>
__global__ void __launch_bounds__(PG<TY>::kThreads, 1)
mag_pair_vec(const MagIterArgs a, const int xchunk, const int pf, const __grid_constant__ PairMaps maps) {
  if (a.ctl_ro && (a.ctl_ro->done || (!a.first && !a.ctl_ro->mag_active))) return;
  using G = PG<TY>;
  constexpr int RS = G::RS, PLA = G::PLA, PLB = G::PLB, SLOTA = G::SLOTA, SLOTB = G::SLOTB;
  constexpr int SP = G::SP, SS = G::SS;
  extern __shared__ __align__(16) float smem[];
  float* const ringA = smem;                    // [3][3][TY+4][RS]
  float* const ringB = smem + 3 * SLOTA;        // [3][3][TY+2][RS]
  float* const stash0 = ringB + 3 * SLOTB;      // [2][7][NT] float4

  const Geom& g = a.g;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const float dt = *static_cast<const float*>(a.dt);
#if AMX_STRICT
  const float dpar = (float)a.np.dx;
#else
  const float dpar = 0.5f * (float)(1.0 / a.np.dx);
#endif
  // element offsets are 32-bit here (pair_mag_supported checks that every array fits): one IMAD.WIDE per global
  // access instead of a 64-bit add chain -- integer work was 36 % of the executed instructions of this kernel
  const int sx = (int)a.B0.s[0], sy = (int)a.B0.s[1];
  const float* const fB0 = static_cast<const float*>(a.B0.p);
  const float* const fG = static_cast<const float*>(a.gas0.p);
  const float* const fBk = static_cast<const float*>(a.Bk.p);
  const float* const fvk = static_cast<const float*>(a.vk.p);
  const int vB0 = (int)a.B0.vs, vG = (int)a.gas0.vs, vBk = (int)a.Bk.vs, vvk = (int)a.vk.vs;
  float* const fB1 = static_cast<float*>(a.Bk1.p);
  float* const fv1 = static_cast<float*>(a.vk1.p);
  const int vB1 = (int)a.Bk1.vs, vv1 = (int)a.vk1.vs;
  const bool xwrap = g.bc[0][0] == BC_PERIODIC;

  F2 n1B = bc2(0.f), n1v = bc2(0.f), n2B = bc2(0.f), n2v = bc2(0.f);
  // (A work list of equal shares as in hydro_stage_vec was measured and dropped: 96 tiles on 148 one-block
  // SMs put neighbouring tiles 40 planes apart in x, their frame rows no longer met in L2, 0.51 -> 0.70 ms.)
  const int z0 = blockIdx.x * 64;
  const int y0 = blockIdx.y * TY;
  const int xs = blockIdx.z * xchunk;
  const int xe = min(xs + xchunk, g.n[0]);
  if (xs >= xe) return;
  const int ta = min(TY, g.n[1] - y0);
  const int tqa = min(32, (g.n[2] - z0) >> 1);
  // ---- role of this thread: window row j, window cell column c, what it computes
  int j, c;
  bool exists, e1, e2 = false, need_rho = true;
  if (wid < TY) {
    j = 2 + wid; c = 2 + 2 * lane;
    exists = wid < ta && lane < tqa;
    e1 = e2 = exists;
  } else if (wid < TY + 2) {
    j = wid == TY ? 1 : ta + 2; c = 2 + 2 * lane;
    exists = lane < tqa;
    e1 = exists;
  } else if (wid < TY + 4) {
    j = wid == TY + 2 ? 0 : ta + 3; c = 2 + 2 * lane;
    exists = lane < tqa;
    e1 = false;
    need_rho = false;
  } else {
    j = lane >> 1; c = (lane & 1) ? 2 + 2 * tqa : 0;
    exists = j < ta + 4;
    e1 = exists && j >= 1 && j <= ta + 2;
  }
  int gy = y0 - 2 + j, gz = z0 - 2 + c;
  gy += gy < 0 ? g.n[1] : (gy >= g.n[1] ? -g.n[1] : 0);
  gz += gz < 0 ? g.n[2] : (gz >= g.n[2] ? -g.n[2] : 0);
  const int col = gz + gy * sy;
  const int smA = j * RS + 2 * c + 4;        // chunk c/2 + 1
  const int smB = (j - 1) * RS + 2 * c + 4;
  float* const stash = stash0 + 4 * (wid * 32 + lane);  // tile threads only

  auto plane_off = [&](int p) -> int {
    if (xwrap) p += p < 0 ? g.n[0] : (p >= g.n[0] ? -g.n[0] : 0);
    return col + p * sx;
  };
  // One block per SM and one barrier per plane leave nothing to overlap a DRAM miss with: the
  // lines of plane p + pf are pulled into L2 while plane p is loaded into registers (one lane per
  // 128-byte line asks; the side warp's pairs each sit in a line of their own).
  const bool pf_lane = !maps.on && exists && (wid >= TY + 4 || (lane & 15) == 0);
  auto issue = [&](int p, RawP& r) {
    if (maps.on && pf > 0 && tid == 32 * (TY + 4) && p + pf <= xe + 1) {
      int q = p + pf;
      if (xwrap) q += q < 0 ? g.n[0] : (q >= g.n[0] ? -g.n[0] : 0);
      q += maps.xpad;
      // The innermost box coordinate must be a multiple of 4 floats (every box row starts at a 16-byte
      // aligned address, else the instruction faults: tools/micro/tma_probe.cu); negative coordinates and
      // boxes that stick out of the array are fine (those parts are skipped).  Window cells z0-2 .. z0+65
      // sit inside the 72-cell box that starts at z0-4.
      const int pz = z0 - 4, py = y0 - 2;
      tma_prefetch_4d(&maps.B0, pz, py, q, 0);
      tma_prefetch_4d(&maps.gas, pz, py, q, 0);
      if constexpr (!FIRST) {
        tma_prefetch_4d(&maps.Bk, pz, py, q, 0);
        tma_prefetch_4d(&maps.vk, pz, py, q, 0);
      }
    }
    if (!exists) return;
    if (pf > 0 && pf_lane && p + pf <= xe + 1) {
      const int q = plane_off(p + pf);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(fB0 + (q + k * vB0)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(fG + (q + (k + 1) * vG)));
      }
      if (need_rho) asm volatile("prefetch.global.L2 [%0];" ::"l"(fG + q));
      if constexpr (!FIRST) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(fBk + (q + k * vBk)));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(fvk + (q + k * vvk)));
        }
      }
    }
    const int o = plane_off(p);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      r.B0[k] = ldg2<0>(fB0 + (o + k * vB0));
      r.v0[k] = ldg2<0>(fG + (o + (k + 1) * vG));
    }
    if (need_rho) r.rho = ldg2<0>(fG + o);
    if constexpr (!FIRST) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        r.Bk[k] = ldg2<0>(fBk + (o + k * vBk));
        r.vk[k] = ldg2<0>(fvk + (o + k * vvk));
      }
    }
  };
  // argument of an evaluation from the iterate (Bi, vi) and R_0; E = B x v (_vector_maths.py:6-12);
  // published to the ring as 3 chunks, returned for the x+1 neighbour of the next evaluation
  auto publish_arg = [&](float* dst, int PL, const F2 (&Bi)[3], const F2 (&vi)[3], const F2 (&B0)[3],
                         const F2 (&v0)[3], bool from_r0, F2 (&B)[3], F2 (&E)[3]) {
    F2 v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (from_r0) { B[k] = B0[k]; v[k] = v0[k]; }
      else if constexpr (MID) { B[k] = (Bi[k] + B0[k]) * bc2(0.5f); v[k] = (vi[k] + v0[k]) * bc2(0.5f); }
      else { B[k] = Bi[k]; v[k] = vi[k]; }
    }
    E[0] = B[1] * v[2] - B[2] * v[1];
    E[1] = B[2] * v[0] - B[0] * v[2];
    E[2] = B[0] * v[1] - B[1] * v[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) sts4(dst + k * PL, mk2(B[k].x, E[k].x), mk2(B[k].y, E[k].y));
  };
  // R_new = R_0 - dt * Psi
  auto advance = [&](const F2 (&curlE)[3], const F2 (&cr)[3], const F2 (&B0)[3], const F2 (&v0)[3], F2 rho,
                     F2 (&Bn)[3], F2 (&vn)[3]) {
#if !AMX_STRICT
    const F2 ir = mk2(rdiv(1.f, rho.x), rdiv(1.f, rho.y));
#endif
    const F2 dt2 = bc2(dt);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#if AMX_STRICT
      const F2 phiB = mk2(cr[k].x / rho.x, cr[k].y / rho.y);
#else
      const F2 phiB = cr[k] * ir;
#endif
      Bn[k] = B0[k] - dt2 * curlE[k];
      vn[k] = v0[k] - dt2 * phiB;
    }
  };
  // running maxima of the squared change norms against (Bp, vp)
  auto change = [&](const F2 (&Bp)[3], const F2 (&vp)[3], const F2 (&Bn)[3], const F2 (&vn)[3], F2& mB, F2& mv) {
    F2 dB = bc2(0.f), dv = bc2(0.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const F2 eB = abs2(Bp[k] - Bn[k]), ev = abs2(vp[k] - vn[k]);
      dB = dB + eB * eB;
      dv = dv + ev * ev;
    }
    mB = max2(mB, dB);
    mv = max2(mv, dv);
  };

  RawP rawX, rawY;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    rawX.B0[k] = rawX.v0[k] = rawX.Bk[k] = rawX.vk[k] = rawY.B0[k] = rawY.v0[k] = rawY.Bk[k] = rawY.vk[k] = bc2(0.f);
  rawX.rho = rawY.rho = bc2(1.f);

  // ring slot of plane p: (p - xs + 6) % 3.
  // sM / sC / sP: slots of planes i-1, i, i+1; plane i+2 shares the slot of plane i-1.
  // At i = xs - 4 these are the slots of planes xs-5, xs-4, xs-3.
  int sM = 1, sC = 2, sP = 0;

  // rawN: plane i+2 (landing); rawC: plane i+1 (argument already published)
  auto step = [&](int i, RawP& rawN, RawP& rawC) {
    // ---- P1(i+2)
    F2 Bx[3], Ex[3];  // the latest argument of the own pair: x+1 neighbour of the evaluation that follows
#pragma unroll
    for (int k = 0; k < 3; ++k) Bx[k] = Ex[k] = bc2(0.f);
    if (exists) publish_arg(ringA + sM * SLOTA + smA, PLA, rawN.Bk, rawN.vk, rawN.B0, rawN.v0, FIRST, Bx, Ex);
    // ---- E1(i+1): ring A planes i (sC), i+1 (sP); plane i+2 in (Bx, Ex)
    if (i >= xs - 2 && e1) {
      F2 curlE[3], cr[3], Bn[3], vn[3];
      psi_ring<PLA, RS>(ringA + sP * SLOTA + smA, ringA + sC * SLOTA + smA, Bx, Ex, dpar, curlE, cr);
      advance(curlE, cr, rawC.B0, rawC.v0, rawC.rho, Bn, vn);
      publish_arg(ringB + sP * SLOTB + smB, PLB, Bn, vn, rawC.B0, rawC.v0, false, Bx, Ex);
      if (e2) {
        if constexpr (FIRST) change(rawC.B0, rawC.v0, Bn, vn, n1B, n1v);
        else change(rawC.Bk, rawC.vk, Bn, vn, n1B, n1v);
        float* st = stash + ((i + 1) & 1) * SS;
        sts4(st, rawC.B0[0], rawC.B0[1]);
        sts4(st + SP, rawC.B0[2], rawC.v0[0]);
        sts4(st + 2 * SP, rawC.v0[1], rawC.v0[2]);
        sts4(st + 3 * SP, rawC.rho, Bn[0]);
        sts4(st + 4 * SP, Bn[1], Bn[2]);
        sts4(st + 5 * SP, vn[0], vn[1]);
        sts4(st + 6 * SP, vn[2], vn[2]);
      }
    }
    if (i + 3 <= xe + 1) issue(i + 3, rawC);
    // ---- E2(i): ring B planes i-1 (sM), i (sC); plane i+1 in (Bx, Ex)
    if (i >= xs && e2) {
      const float* st = stash + (i & 1) * SS;
      F2 B0[3], v0[3], B1[3], v1[3], curlE[3], cr[3], Bn[3], vn[3];
      const float4 q0 = lds4(st), q1 = lds4(st + SP), q2 = lds4(st + 2 * SP), q3 = lds4(st + 3 * SP),
                   q4 = lds4(st + 4 * SP), q5 = lds4(st + 5 * SP), q6 = lds4(st + 6 * SP);
      B0[0] = mk2(q0.x, q0.y); B0[1] = mk2(q0.z, q0.w); B0[2] = mk2(q1.x, q1.y);
      v0[0] = mk2(q1.z, q1.w); v0[1] = mk2(q2.x, q2.y); v0[2] = mk2(q2.z, q2.w);
      const F2 rho = mk2(q3.x, q3.y);
      B1[0] = mk2(q3.z, q3.w); B1[1] = mk2(q4.x, q4.y); B1[2] = mk2(q4.z, q4.w);
      v1[0] = mk2(q5.x, q5.y); v1[1] = mk2(q5.z, q5.w); v1[2] = mk2(q6.x, q6.y);
      psi_ring<PLB, RS>(ringB + sC * SLOTB + smB, ringB + sM * SLOTB + smB, Bx, Ex, dpar, curlE, cr);
      advance(curlE, cr, B0, v0, rho, Bn, vn);
      change(B1, v1, Bn, vn, n2B, n2v);
      const int o = col + i * sx;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        stg2(fB1 + (o + k * vB1), Bn[k]);
        stg2(fv1 + (o + k * vv1), vn[k]);
      }
      // slab rank: the edge planes also go into the neighbours' ghost planes (stores over NVLink)
      long long pbd[2];
      int pps[2];
      bool pon[2] = {false, false};
      if (a.peers.flags) {
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          pon[side] = peer_target(a.peers, i, g.n[0], side, pbd[side], pps[side]);
          if (pon[side]) {
            const int po = o + pps[side] * sx;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              stg2(reinterpret_cast<float*>(reinterpret_cast<char*>(fB1 + (po + k * vB1)) + pbd[side]), Bn[k]);
              stg2(reinterpret_cast<float*>(reinterpret_cast<char*>(fv1 + (po + k * vv1)) + pbd[side]), vn[k]);
            }
          }
        }
      }
      if constexpr (EMIT) {
        // this launch is expected to end the loop: write what mag_final would from R_{k+2}
        // (rho, v_n, p_2 -> emit_gas; B_n -> emit_B); valid only if the while-condition agrees
        // (Ctl::final_emitted), else mag_final runs as usual
        const F2 pr = ldg2<0>(fG + (o + 4 * vG));
        F2 n0 = bc2(0.f), nn = bc2(0.f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          n0 = n0 + abs2(v0[k]) * abs2(v0[k]);
          nn = nn + abs2(vn[k]) * abs2(vn[k]);
        }
        const F2 p2 = mk2(magnetic_writeback_pressure(rho.x, pr.x, n0.x, nn.x),
                          magnetic_writeback_pressure(rho.y, pr.y, n0.y, nn.y));
        float* const eg0 = static_cast<float*>(a.emit_gas.p);
        float* const eb0 = static_cast<float*>(a.emit_B.p);
        float* const eg = eg0 + o;
        float* const eb = eb0 + o;
        const int veg = (int)a.emit_gas.vs, veb = (int)a.emit_B.vs;
        stg2(eg, rho);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          stg2(eg0 + (o + (k + 1) * veg), vn[k]);
          stg2(eb0 + (o + k * veb), Bn[k]);
        }
        stg2(eg0 + (o + 4 * veg), p2);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          if (pon[side]) {
            float* const pg = reinterpret_cast<float*>(reinterpret_cast<char*>(eg) + pbd[side]) + pps[side] * sx;
            float* const pb = reinterpret_cast<float*>(reinterpret_cast<char*>(eb) + pbd[side]) + pps[side] * sx;
            stg2(pg, rho);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              stg2(pg + (k + 1) * veg, vn[k]);
              stg2(pb + k * veb, Bn[k]);
            }
            stg2(pg + 4 * veg, p2);
          }
        }
      }
    }
    __syncthreads();
    const int t = sM; sM = sC; sC = sP; sP = t;
  };

  issue(xs - 2, rawX);
  for (int i = xs - 4;;) {
    step(i, rawX, rawY);
    if (++i >= xe) break;
    step(i, rawY, rawX);
    if (++i >= xe) break;
  }

  // ---- change norms: sqrt of the block maxima of the squared 3-norms -> spread slots
  float chs[4] = {rsqrt_(fmaxf(n1B.x, n1B.y)), rsqrt_(fmaxf(n1v.x, n1v.y)), rsqrt_(fmaxf(n2B.x, n2B.y)),
                  rsqrt_(fmaxf(n2v.x, n2v.y))};
  __shared__ float red[4][G::kWarps];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chs[k] = fmaxf(chs[k], __shfl_xor_sync(0xffffffffu, chs[k], o));
    if (lane == 0) red[k][wid] = chs[k];
  }
  __syncthreads();
  if (tid < 4) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < G::kWarps; ++i) m = fmaxf(m, red[tid][i]);
    const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    unsigned long long* slots = tid < 2 ? a.change_slots : a.change_slots2;
    atomicMax(slots + (size_t)(tid & 1) * kAlphaSlots + (bid % kAlphaSlots), bits_of(m));
  }
}

template <bool FIRST, bool MID, int TY, bool EMIT>
cudaError_t launch_pair_e(const MagIterArgs& a, cudaStream_t st) {
  using G = PG<TY>;
  auto kern = mag_pair_vec<FIRST, MID, TY, EMIT>;
  static SmemOptIn opt;
  {
    cudaError_t e = opt.ensure(kern, G::smem_bytes);
    if (e != cudaSuccess) return e;
  }
  dim3 grid;
  grid.x = (a.g.n[2] + 63) / 64;
  grid.y = (a.g.n[1] + TY - 1) / TY;
  const int nch = pick_chunks(a.g.n[0], (long long)grid.x * grid.y, 1, 3);
  const int xchunk = (a.g.n[0] + nch - 1) / nch;
  grid.z = (a.g.n[0] + xchunk - 1) / xchunk;
  // AMX_PAIR_PF: planes ahead whose lines are pulled into L2 (0: off)
  static const int pf = [] { const char* e = getenv("AMX_PAIR_PF"); return e ? atoi(e) : 2; }();
  // AMX_PAIR_TMAPF=0: per-lane prefetch instructions instead of tensor-map prefetches
  static const bool tmapf = [] { const char* e = getenv("AMX_PAIR_TMAPF"); return !(e && e[0] == '0'); }();
  PairMaps maps;
  maps.on = 0;
  maps.xpad = a.g.bc[0][0] == BC_HALO ? 2 : 0;
  if (tmapf && pf > 0) {
    // (a box must fit into the array: small grids prefetch a clipped window)
    const unsigned bz = a.g.n[2] < 72 ? (unsigned)a.g.n[2] : 72u, by = a.g.n[1] < TY + 4 ? (unsigned)a.g.n[1] : (unsigned)(TY + 4);
    bool ok = tma_map_rows_3d(&maps.B0, a.B0, a.g, 3, maps.xpad, bz, by, 3) &&
              tma_map_rows_3d(&maps.gas, a.gas0, a.g, 4, maps.xpad, bz, by, 4);
    if (ok && !FIRST)
      ok = tma_map_rows_3d(&maps.Bk, a.Bk, a.g, 3, maps.xpad, bz, by, 3) &&
           tma_map_rows_3d(&maps.vk, a.vk, a.g, 3, maps.xpad, bz, by, 3);
    maps.on = ok ? 1 : 0;
  }
  kern<<<grid, G::kThreads, G::smem_bytes, st>>>(a, xchunk, pf, maps);
  return cudaGetLastError();
}

template <bool FIRST, bool MID, int TY>
cudaError_t launch_pair(const MagIterArgs& a, cudaStream_t st) {
  return a.emit ? launch_pair_e<FIRST, MID, TY, true>(a, st) : launch_pair_e<FIRST, MID, TY, false>(a, st);
}

constexpr int kPairTY = 11;

}  // namespace vec

// fp32, 3-D, y and z periodic, x periodic or resident halo planes, same layout rules as mag_iter_vec
bool pair_mag_supported(const MagIterArgs& a) {
  if (!vec_mag_supported(a)) return false;
  if (a.g.nd != 3 || !a.change_slots2) return false;
  if (a.g.bc[1][0] != BC_PERIODIC || a.g.bc[1][1] != BC_PERIODIC) return false;
  if (a.g.bc[2][0] != BC_PERIODIC || a.g.bc[2][1] != BC_PERIODIC) return false;
  const bool xper = a.g.bc[0][0] == BC_PERIODIC && a.g.bc[0][1] == BC_PERIODIC;
  const bool xhalo = a.g.bc[0][0] == BC_HALO && a.g.bc[0][1] == BC_HALO;
  if (!xper && !xhalo) return false;
  if (a.emit) {
    if (!vec::view_ok(a.emit_gas, 3) || !vec::view_ok(a.emit_B, 3)) return false;
    for (int ax = 0; ax < 2; ++ax)
      if (a.emit_gas.s[ax] != a.B0.s[ax] || a.emit_B.s[ax] != a.B0.s[ax]) return false;
  }
  // 32-bit element offsets inside the kernel: rows * row stride (+ the ghost planes) of every array below 2^31
  auto fits = [&](const View& v, int rows) {
    const long long ext = (long long)(rows - 1) * v.vs + (long long)(a.g.n[0] + 4) * v.s[0];
    return v.vs >= 0 && ext < 2147483647LL && (long long)rows * v.vs < 2147483647LL;
  };
  if (!fits(a.B0, 3) || !fits(a.gas0, 5) || !fits(a.Bk1, 3) || !fits(a.vk1, 3)) return false;
  if (a.Bk.p && (!fits(a.Bk, 3) || !fits(a.vk, 3))) return false;
  if (a.emit && (!fits(a.emit_gas, 5) || !fits(a.emit_B, 3))) return false;
  return a.g.n[0] >= 4 && a.g.n[1] >= 4 && a.g.n[2] >= 8;
}

cudaError_t launch_mag_pair_vec(const MagIterArgs& a, cudaStream_t st) {
  constexpr int TY = vec::kPairTY;
  if (!a.Bk.p) return a.midpoint ? vec::launch_pair<true, true, TY>(a, st) : vec::launch_pair<true, false, TY>(a, st);
  return a.midpoint ? vec::launch_pair<false, true, TY>(a, st) : vec::launch_pair<false, false, TY>(a, st);
}

}  // namespace AMX_NS
