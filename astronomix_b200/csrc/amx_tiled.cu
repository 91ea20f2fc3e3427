// amx_tiled.cu -- plane-streaming shared-memory kernels for the 2-D / 3-D hot configurations.
//
// hydro_stage_tiled: one unsplit forward-Euler stage (evolve_state.py:339-414) [+ RK2 combine
// :456-469], every interface flux and every limited slope computed exactly once per block:
//   * a block owns a TY x TZ column tile (3-D: 8 x 32 cells, z contiguous; 2-D: 1 x 256) and
//     marches along x over a chunk of planes;
//   * x direction: the thread carries s(x), s(x+1), the x half-slope of cell x, the difference
//     across x+1/2 and the flux at x-1/2 in registers -- one new plane load per step, no
//     shared memory, no recomputation;
//   * y / z directions: the plane window (tile + 2-cell halo, boundary rules applied while
//     loading) sits in a 3-deep shared-memory ring; each thread computes the half-slopes of its
//     own cell, publishes them, computes the fluxes at its two lower interfaces from its own and
//     its neighbours' published slopes, publishes those, and reads the upper-interface fluxes
//     back.  The tile-edge slopes / fluxes are extra work items spread over otherwise idle lanes.
//   * conservative update, cons->prim, optional RK2 combine with s0, store, running max of
//     |u|+c for the next CFL / Lax-Friedrichs alpha: all in the same pass.
// Algorithmic traffic: read ng (+ng for the combine), write ng words per cell; the window halo
// (3-D: 12x36 / 8x32 = 1.69x of the read) is served by L2.
//
// mag_iter_tiled: one evaluation of the Pang-Wu fixed point (_magnetic_field_update.py:82-165)
// with the same marching scheme (1-cell halo, arguments averaged once per cell, E = B x v
// computed once per cell).
//
// Compiled twice like amx_kernels.cu (fast / strict arithmetic).
#include "amx_internal.h"
#include "amx_math.cuh"

#if AMX_STRICT
#define AMX_NS amx_strict
#else
#define AMX_NS amx_fast
#endif

namespace AMX_NS {
using namespace amx;

namespace tiled {

template <int ND>
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

template <typename T>
AMX_DI Consts<T> make_consts(const NumPar& np) {
  Consts<T> c;
  c.gamma = (T)np.gamma;
  c.gm1 = c.gamma - T(1);
  c.inv_gm1 = T(1) / c.gm1;
  c.dx = (T)np.dx;
  c.inv_dx = (T)(1.0 / np.dx);
  c.half_dx = (T)(np.dx / 2);
  c.va_eps = (T)(3 * np.dx);
  return c;
}

AMX_DI unsigned long long bits_of(float v) { return (unsigned long long)__float_as_uint(v); }
AMX_DI unsigned long long bits_of(double v) { return (unsigned long long)__double_as_longlong(v); }

// limited HALF difference d = (limiter(dm/dx, dp/dx) * dx) / 2 of one variable from the
// differences across the cell's lower (dm) and upper (dp) interface
// (limited_gradients.py:50-60 + reconstruction.py:409-412)
template <typename T, int LIM>
AMX_DI T half_slope(T dm, T dp, int lim, const Consts<T>& c) {
#if AMX_STRICT
  const T a = dm / c.dx, b = dp / c.dx;
  const T g = (LIM == LIM_MINMOD) ? minmod(a, b) : limited_slope(a, b, lim, c);
  return (g * c.dx) / T(2);
#else
  if constexpr (LIM == LIM_MINMOD) {
    return T(0.5) * minmod(dm, dp);
  } else {
    const T g = limited_slope(dm * c.inv_dx, dp * c.inv_dx, lim, c);
    return (g * c.dx) * T(0.5);
  }
#endif
}

template <int ND> struct Tile;
template <> struct Tile<3> { static constexpr int TY = 8, TZ = 32; };
template <> struct Tile<2> { static constexpr int TY = 1, TZ = 256; };

template <typename T, int ND>
constexpr size_t hydro_smem_bytes() {
  constexpr int NV = ND + 2, TY = Tile<ND>::TY, TZ = Tile<ND>::TZ;
  constexpr int WY = ND == 3 ? TY + 4 : 1, WZ = TZ + 4;
  size_t n = 3 * NV * WY * WZ + NV * TY * (TZ + 2) + NV * TY * (TZ + 1);
  if (ND == 3) n += NV * (TY + 2) * TZ + NV * (TY + 1) * TZ;
  return n * sizeof(T);
}

template <typename T, int ND, int RS, int LIM>
__global__ void __launch_bounds__(256) hydro_stage_tiled(const StageArgs a, const int xchunk) {
  if (a.ctl_ro && a.ctl_ro->done) return;
  constexpr int NV = ND + 2;
  constexpr int TY = Tile<ND>::TY, TZ = Tile<ND>::TZ;
  constexpr int WY = ND == 3 ? TY + 4 : 1, WZ = TZ + 4;
  constexpr int YO = ND == 3 ? 2 : 0;  // row offset of the tile inside the window
  constexpr int LA = ND - 1;           // lane (contiguous) axis
  constexpr int PL = WY * WZ;          // window plane size per variable
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* const ring = reinterpret_cast<T*>(smem_raw);    // [3][NV][WY][WZ]
  T* const Dz = ring + 3 * NV * PL;                   // [NV][TY][TZ+2]  half-slopes along the lane axis
  T* const Fz = Dz + NV * TY * (TZ + 2);              // [NV][TY][TZ+1]  fluxes at lane-axis interfaces
  T* const Dy = Fz + NV * TY * (TZ + 1);              // [NV][TY+2][TZ]  (3-D) half-slopes along y
  T* const Fy = Dy + NV * (TY + 2) * TZ;              // [NV][TY+1][TZ]  (3-D) fluxes at y interfaces

  const Consts<T> c = make_consts<T>(a.np);
  const Geom& g = a.g;
  const int tid = threadIdx.x;
  const int tz = tid % TZ, ty = tid / TZ;
  const int z0 = blockIdx.x * TZ;
  const int y0 = ND == 3 ? blockIdx.y * TY : 0;
  const int chunk = ND == 3 ? blockIdx.z : blockIdx.y;
  const int xs = chunk * xchunk;
  const int xe = min(xs + xchunk, g.n[0]);
  if (xs >= xe) return;
  const T dt = *static_cast<const T*>(a.dt) * (T)a.dt_scale;

  // ---- this thread's own column and its extra window-halo cell, boundary rules applied
  bool own_fy = false, own_fz = false, h_fy = false, h_fz = false;
  const int gz = z0 + tz, gy = y0 + ty;
  const bool col_active = gz < g.n[LA] && (ND == 2 || gy < g.n[1]);
  long long own_off;  // offset of the (mapped) own column in `in`, plane 0
  {
    const int mz = map_index<ND>(gz, g.n[LA], g.bc[LA][0], g.bc[LA][1], own_fz);
    own_off = (long long)mz * a.in.s[LA];
    if constexpr (ND == 3) {
      const int my = map_index<ND>(gy, g.n[1], g.bc[1][0], g.bc[1][1], own_fy);
      own_off += (long long)my * a.in.s[1];
    }
  }
  constexpr int NHALO = ND == 3 ? (WY * WZ - TY * TZ) : 4;
  int hpos = -1;  // window position (j * WZ + k) of the halo cell this thread loads
  long long h_off = 0;
  if (tid < NHALO) {
    int j, k;
    if constexpr (ND == 3) {
      if (tid < 4 * WZ) {
        const int r = tid / WZ;
        j = r < 2 ? r : (TY + r);  // rows 0, 1, TY+2, TY+3
        k = tid % WZ;
      } else {
        const int h2 = tid - 4 * WZ;
        j = 2 + h2 / 4;
        const int q = h2 % 4;
        k = q < 2 ? q : (TZ + q);  // columns 0, 1, TZ+2, TZ+3
      }
    } else {
      j = 0;
      k = tid < 2 ? tid : (TZ + tid);
    }
    hpos = j * WZ + k;
    const int mz = map_index<ND>(z0 - 2 + k, g.n[LA], g.bc[LA][0], g.bc[LA][1], h_fz);
    h_off = (long long)mz * a.in.s[LA];
    if constexpr (ND == 3) {
      const int my = map_index<ND>(y0 - 2 + j, g.n[1], g.bc[1][0], g.bc[1][1], h_fy);
      h_off += (long long)my * a.in.s[1];
    }
  }
  const T* const in = static_cast<const T*>(a.in.p);
  const long long vs = a.in.vs;

  // load the (mapped) plane xp: own cell into v[], halo cell into hv[]
  auto load_plane = [&](int xp, T (&v)[NV], T (&hv)[NV], bool want_halo) {
    bool fx;
    const int mx = map_index<ND>(xp, g.n[0], g.bc[0][0], g.bc[0][1], fx);
    const long long po = (long long)mx * a.in.s[0];
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = __ldg(in + k * vs + po + own_off);
    if (fx) v[1] = -v[1];
    if (ND == 3 && own_fy) v[2] = -v[2];
    if (own_fz) v[LA + 1] = -v[LA + 1];
    if (want_halo && hpos >= 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) hv[k] = __ldg(in + k * vs + po + h_off);
      if (fx) hv[1] = -hv[1];
      if (ND == 3 && h_fy) hv[2] = -hv[2];
      if (h_fz) hv[LA + 1] = -hv[LA + 1];
    }
  };
  auto store_window = [&](int slot, const T (&v)[NV], const T (&hv)[NV]) {
    T* w = ring + slot * NV * PL;
#pragma unroll
    for (int k = 0; k < NV; ++k) w[k * PL + (ty + YO) * WZ + tz + 2] = v[k];
    if (hpos >= 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) w[k * PL + hpos] = hv[k];
    }
  };

  // conservative update of one variable: U + ((1/dx) * (Fl - Fr)) * dt (evolve_state.py:398-403)
#if AMX_STRICT
  auto upd = [&](T U, T fl, T fr) -> T { return U + (c.inv_dx * (fl - fr)) * dt; };
#else
  const T dtdx = c.inv_dx * dt;
  auto upd = [&](T U, T fl, T fr) -> T { return U + dtdx * (fl - fr); };
#endif

  // ---- prologue: planes xs-2 .. xs+1
  T sC[NV], sP[NV], dC[NV], FxL[NV], hv[NV];
  {
    T sM2[NV], sM1[NV];
    load_plane(xs - 2, sM2, hv, false);
    load_plane(xs - 1, sM1, hv, false);
    load_plane(xs, sC, hv, true);
    store_window(0, sC, hv);
    load_plane(xs + 1, sP, hv, true);
    store_window(1, sP, hv);
    Gas<T, ND> L, R;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const T d0 = sC[k] - sM1[k];
      const T dM1 = half_slope<T, LIM>(sM1[k] - sM2[k], d0, a.np.lim, c);
      dC[k] = half_slope<T, LIM>(d0, sP[k] - sC[k], a.np.lim, c);
      L.q[k] = sM1[k] + dM1;
      R.q[k] = sC[k] - dC[k];
    }
    T alpha = T(0);
    if constexpr (RS == RS_LF) alpha = static_cast<const T*>(a.alpha_lf)[0];
    const Gas<T, ND> F = riemann<T, ND, RS>(L, R, 1, alpha, c);
#pragma unroll
    for (int k = 0; k < NV; ++k) FxL[k] = F.q[k];
  }
  __syncthreads();

  T wb[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) wb[k] = T(0);
  T alpha_lf[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) alpha_lf[k] = RS == RS_LF ? static_cast<const T*>(a.alpha_lf)[k] : T(0);

  T* const out_col = static_cast<T*>(a.out.p) + (long long)gz * a.out.s[LA] + (ND == 3 ? (long long)gy * a.out.s[1] : 0);
  const T* const in0_col = a.in0.p ? static_cast<const T*>(a.in0.p) + (long long)gz * a.in0.s[LA] +
                                         (ND == 3 ? (long long)gy * a.in0.s[1] : 0)
                                   : nullptr;
  const int wc = (ty + YO) * WZ + tz + 2;  // own cell in the window
  const int dzo = ty * (TZ + 2) + tz + 1;  // own entry of Dz (per variable: + k*TY*(TZ+2))
  const int fzo = ty * (TZ + 1) + tz;      // own lower-interface entry of Fz
  const int dyo = (ty + 1) * TZ + tz;      // own entry of Dy
  const int fyo = ty * TZ + tz;            // own lower-interface entry of Fy

  int slotC = 0;
  for (int x = xs; x < xe; ++x) {
    const int slotN = slotC == 0 ? 2 : slotC - 1;  // (slotC + 2) % 3
    T sN[NV];
    load_plane(x + 2, sN, hv, true);  // consumed at the end of the step
    const T* w = ring + slotC * NV * PL;

    // ---- phase 1: half-slopes of the own cell along z (and y), published for the neighbours
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const T s0 = sC[k];
      Dz[k * TY * (TZ + 2) + dzo] = half_slope<T, LIM>(s0 - w[k * PL + wc - 1], w[k * PL + wc + 1] - s0, a.np.lim, c);
      if constexpr (ND == 3)
        Dy[k * (TY + 2) * TZ + dyo] = half_slope<T, LIM>(s0 - w[k * PL + wc - WZ], w[k * PL + wc + WZ] - s0, a.np.lim, c);
    }
    // tile-edge slopes: cells one outside the tile (rows -1 / TY, columns -1 / TZ)
    if constexpr (ND == 3) {
      if (tid < 2 * TZ) {
        const int row = tid < TZ ? -1 : TY;
        const int p = (row + 2) * WZ + (tid % TZ) + 2;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const T s0 = w[k * PL + p];
          Dy[(k * (TY + 2) + row + 1) * TZ + (tid % TZ)] =
              half_slope<T, LIM>(s0 - w[k * PL + p - WZ], w[k * PL + p + WZ] - s0, a.np.lim, c);
        }
      } else if (tid < 2 * TZ + 2 * TY) {
        const int i = tid - 2 * TZ;
        const int col = i < TY ? -1 : TZ, yy = i % TY;
        const int p = (yy + 2) * WZ + col + 2;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const T s0 = w[k * PL + p];
          Dz[(k * TY + yy) * (TZ + 2) + col + 1] =
              half_slope<T, LIM>(s0 - w[k * PL + p - 1], w[k * PL + p + 1] - s0, a.np.lim, c);
        }
      }
    } else {
      if (tid < 2) {
        const int col = tid == 0 ? -1 : TZ;
        const int p = col + 2;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const T s0 = w[k * PL + p];
          Dz[k * (TZ + 2) + col + 1] = half_slope<T, LIM>(s0 - w[k * PL + p - 1], w[k * PL + p + 1] - s0, a.np.lim, c);
        }
      }
    }
    __syncthreads();  // published slopes visible

    // ---- phase 2: fluxes at the own cell's lower z (and y) interfaces, published
    {
      Gas<T, ND> L, R;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        L.q[k] = w[k * PL + wc - 1] + Dz[k * TY * (TZ + 2) + dzo - 1];
        R.q[k] = sC[k] - Dz[k * TY * (TZ + 2) + dzo];
      }
      const Gas<T, ND> F = riemann<T, ND, RS>(L, R, LA + 1, alpha_lf[LA], c);
#pragma unroll
      for (int k = 0; k < NV; ++k) Fz[k * TY * (TZ + 1) + fzo] = F.q[k];
    }
    if constexpr (ND == 3) {
      Gas<T, ND> L, R;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        L.q[k] = w[k * PL + wc - WZ] + Dy[k * (TY + 2) * TZ + dyo - TZ];
        R.q[k] = sC[k] - Dy[k * (TY + 2) * TZ + dyo];
      }
      const Gas<T, ND> F = riemann<T, ND, RS>(L, R, 2, alpha_lf[1], c);
#pragma unroll
      for (int k = 0; k < NV; ++k) Fy[k * (TY + 1) * TZ + fyo] = F.q[k];
    }
    // tile-edge fluxes: the interfaces above the last row / column of the tile
    if constexpr (ND == 3) {
      if (tid >= 96 && tid < 96 + TZ) {
        const int zz = tid - 96;
        const int p = (TY + 1) * WZ + zz + 2;  // cell (TY-1, zz); the cell above is p + WZ
        Gas<T, ND> L, R;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          L.q[k] = w[k * PL + p] + Dy[(k * (TY + 2) + TY) * TZ + zz];
          R.q[k] = w[k * PL + p + WZ] - Dy[(k * (TY + 2) + TY + 1) * TZ + zz];
        }
        const Gas<T, ND> F = riemann<T, ND, RS>(L, R, 2, alpha_lf[1], c);
#pragma unroll
        for (int k = 0; k < NV; ++k) Fy[(k * (TY + 1) + TY) * TZ + zz] = F.q[k];
      } else if (tid >= 128 && tid < 128 + TY) {
        const int yy = tid - 128;
        const int p = (yy + 2) * WZ + TZ + 1;  // cell (yy, TZ-1); the next cell is p + 1
        Gas<T, ND> L, R;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          L.q[k] = w[k * PL + p] + Dz[(k * TY + yy) * (TZ + 2) + TZ];
          R.q[k] = w[k * PL + p + 1] - Dz[(k * TY + yy) * (TZ + 2) + TZ + 1];
        }
        const Gas<T, ND> F = riemann<T, ND, RS>(L, R, 3, alpha_lf[2], c);
#pragma unroll
        for (int k = 0; k < NV; ++k) Fz[(k * TY + yy) * (TZ + 1) + TZ] = F.q[k];
      }
    } else {
      if (tid == 32) {
        const int p = TZ + 1;
        Gas<T, ND> L, R;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          L.q[k] = w[k * PL + p] + Dz[k * (TZ + 2) + TZ];
          R.q[k] = w[k * PL + p + 1] - Dz[k * (TZ + 2) + TZ + 1];
        }
        const Gas<T, ND> F = riemann<T, ND, RS>(L, R, 2, alpha_lf[1], c);
#pragma unroll
        for (int k = 0; k < NV; ++k) Fz[k * (TZ + 1) + TZ] = F.q[k];
      }
    }

    // ---- phase 3a: x direction from registers -- flux at x+1/2 (first use of plane x+2)
    T dP[NV], FxR[NV];
    {
      Gas<T, ND> L, R;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        dP[k] = half_slope<T, LIM>(sP[k] - sC[k], sN[k] - sP[k], a.np.lim, c);
        L.q[k] = sC[k] + dC[k];
        R.q[k] = sP[k] - dP[k];
      }
      const Gas<T, ND> F = riemann<T, ND, RS>(L, R, 1, alpha_lf[0], c);
#pragma unroll
      for (int k = 0; k < NV; ++k) FxR[k] = F.q[k];
    }
    __syncthreads();  // published fluxes visible

    // ---- phase 3b: conservative update, axis by axis like the reference, prim (+ RK2 combine)
    {
      Gas<T, ND> sc;
#pragma unroll
      for (int k = 0; k < NV; ++k) sc.q[k] = sC[k];
      Gas<T, ND> U = cons_from_prim(sc, c);
#pragma unroll
      for (int k = 0; k < NV; ++k) U.q[k] = upd(U.q[k], FxL[k], FxR[k]);
      if constexpr (ND == 3) {
#pragma unroll
        for (int k = 0; k < NV; ++k)
          U.q[k] = upd(U.q[k], Fy[k * (TY + 1) * TZ + fyo], Fy[k * (TY + 1) * TZ + fyo + TZ]);
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) U.q[k] = upd(U.q[k], Fz[k * TY * (TZ + 1) + fzo], Fz[k * TY * (TZ + 1) + fzo + 1]);
      Gas<T, ND> s = prim_from_cons(U, c);
      if (col_active) {
        if (in0_col) {
          Gas<T, ND> s0;
          const T* p0 = in0_col + (long long)x * a.in0.s[0];
#pragma unroll
          for (int k = 0; k < NV; ++k) s0.q[k] = __ldg(p0 + k * a.in0.vs);
          const Gas<T, ND> U0 = cons_from_prim(s0, c), U2 = cons_from_prim(s, c);
          Gas<T, ND> Um;
#pragma unroll
          for (int k = 0; k < NV; ++k) Um.q[k] = T(0.5) * (U0.q[k] + U2.q[k]);
          s = prim_from_cons(Um, c);
        }
        T* po = out_col + (long long)x * a.out.s[0];
#pragma unroll
        for (int k = 0; k < NV; ++k) po[k * a.out.vs] = s.q[k];
        if (a.alpha_slots) {
#pragma unroll
          for (int k = 0; k < ND; ++k) wb[k] = rmax(wb[k], wave_bound(s, k + 1, c));
        }
      }
    }

    // ---- rotate: plane x+2 enters the ring, registers shift by one plane
    store_window(slotN, sN, hv);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      sC[k] = sP[k];
      sP[k] = sN[k];
      dC[k] = dP[k];
      FxL[k] = FxR[k];
    }
    slotC = slotC == 2 ? 0 : slotC + 1;
  }

  if (a.alpha_slots) {
    __shared__ T red[ND][8];
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wb[k] = rmax(wb[k], __shfl_xor_sync(0xffffffffu, wb[k], o));
      if (lane == 0) red[k][wid] = wb[k];
    }
    __syncthreads();
    if (wid == 0) {
      const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        T m = lane < 8 ? red[k][lane] : T(0);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = rmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(a.alpha_slots + (size_t)k * kAlphaSlots + (bid % kAlphaSlots), bits_of(m));
      }
    }
  }
}

inline int pick_xchunk(int nx, long long tiles) {
  const long long target = 148LL * 6;  // >= ~2 waves at 3 resident blocks per SM
  long long nch = (target + tiles - 1) / tiles;
  if (nch < 1) nch = 1;
  const long long maxch = nx / 16 > 0 ? nx / 16 : 1;  // keep the 3-plane prologue below ~20 %
  if (nch > maxch) nch = maxch;
  return (int)((nx + nch - 1) / nch);
}

template <typename T, int ND, int RS, int LIM>
cudaError_t launch_stage(const StageArgs& a, cudaStream_t st) {
  constexpr int TY = Tile<ND>::TY, TZ = Tile<ND>::TZ;
  const size_t smem = hydro_smem_bytes<T, ND>();
  auto kern = hydro_stage_tiled<T, ND, RS, LIM>;
  static SmemOptIn opt;
  {
    cudaError_t e = opt.ensure(kern, smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid;
  const int LA = ND - 1;
  grid.x = (a.g.n[LA] + TZ - 1) / TZ;
  const long long tiles = ND == 3 ? (long long)grid.x * ((a.g.n[1] + TY - 1) / TY) : grid.x;
  const int xchunk = pick_xchunk(a.g.n[0], tiles);
  const int nch = (a.g.n[0] + xchunk - 1) / xchunk;
  if (ND == 3) { grid.y = (a.g.n[1] + TY - 1) / TY; grid.z = nch; }
  else { grid.y = nch; grid.z = 1; }
  kern<<<grid, 256, smem, st>>>(a, xchunk);
  return cudaGetLastError();
}


// ================================================================ magnetic fixed point
// One evaluation R_{k+1} = R_0 - dt * Psi(arg(R_k, R_0)) (_magnetic_field_update.py:82-165).
// Per loaded plane every cell's argument (B, v) is averaged once, E = B x v is formed once,
// and (B, E) go into a 4-deep shared-memory ring (1-cell halo); the curls of plane x are then
// central differences over the ring (x-1, x, x+1).  R_0, R_k and rho of the own cell stay in
// registers between the load of a plane and its update one step later.
// 4-byte asynchronous global->shared copy (LDGSTS); the data never passes through registers
AMX_DI void cp_async4(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
AMX_DI void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
template <typename T>
AMX_DI void cp_async_elem(T* smem_dst, const T* gmem_src) {
  if constexpr (sizeof(T) == 4) cp_async4(smem_dst, gmem_src);
  else cp_async8(smem_dst, gmem_src);
}
AMX_DI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
AMX_DI void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int ND>
constexpr size_t mag_smem_bytes() {
  constexpr int TY = Tile<ND>::TY, TZ = Tile<ND>::TZ;
  constexpr int WY = ND == 3 ? TY + 2 : 1, WZ = TZ + 2;
  return (size_t)4 * (13 + 6) * WY * WZ * sizeof(T);
}

// Plane pipeline of mag_iter_tiled (per block, planes p of the x chunk):
//   cp.async raw(p)  ->  [2 planes later]  publish (B, E)(p) to the ring  ->  [2 planes later]  update(p)
// raw(p) = B0, v0, rho, Bk, vk of the own cell and of one halo cell per thread, staged in shared
// memory by LDGSTS (thread-private entries, so only cp.async.wait_group orders them); two planes
// of loads are always in flight per block.  One __syncthreads per plane orders the (B, E) ring.
template <typename T, int ND>
__global__ void __launch_bounds__(256) mag_iter_tiled(const MagIterArgs a, const int xchunk) {
  if (a.ctl_ro && (a.ctl_ro->done || (!a.first && !a.ctl_ro->mag_active))) return;
  constexpr int TY = Tile<ND>::TY, TZ = Tile<ND>::TZ;
  constexpr int WY = ND == 3 ? TY + 2 : 1, WZ = TZ + 2;
  constexpr int YO = ND == 3 ? 1 : 0;
  constexpr int LA = ND - 1;
  constexpr int PL = WY * WZ;
  constexpr int NRAW = 13;  // B0[3], v0[3], rho, Bk[3], vk[3]
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* const stage = reinterpret_cast<T*>(smem_raw);  // [4][13][PL]
  T* const ring = stage + 4 * NRAW * PL;            // [4][6][PL]: B (3) then E (3)

  const Consts<T> c = make_consts<T>(a.np);
  const Geom& g = a.g;
  const int tid = threadIdx.x;
  const int tz = tid % TZ, ty = tid / TZ;
  const int z0 = blockIdx.x * TZ;
  const int y0 = ND == 3 ? blockIdx.y * TY : 0;
  const int chunk = ND == 3 ? blockIdx.z : blockIdx.y;
  const int xs = chunk * xchunk;
  const int xe = min(xs + xchunk, g.n[0]);
  if (xs >= xe) return;
  const T dt = *static_cast<const T*>(a.dt);
  const bool have_k = a.Bk.p != nullptr;
  const bool midpoint = a.midpoint != 0;

  // Mapped own column / halo cell: one base pointer per array, flip flag.  Q12: a reflective
  // side of axis `ax` (1-based) flips 3-vector component index ax, i.e. component 1 (y) for the
  // x axis, component 2 (z) for the y axis and nothing for the z axis.
  struct Col { const T *B0, *gas, *Bk, *vk; bool f2; };
  auto make_col = [&](int wy, int wz, Col& col) {
    bool fz, fy = false;
    const int mz = map_index<ND>(z0 - 1 + wz, g.n[LA], g.bc[LA][0], g.bc[LA][1], fz);
    int my = 0;
    if constexpr (ND == 3) my = map_index<ND>(y0 - 1 + wy, g.n[1], g.bc[1][0], g.bc[1][1], fy);
    col.B0 = static_cast<const T*>(a.B0.p) + (long long)mz * a.B0.s[LA] + (ND == 3 ? (long long)my * a.B0.s[1] : 0);
    col.gas = static_cast<const T*>(a.gas0.p) + (long long)mz * a.gas0.s[LA] + (ND == 3 ? (long long)my * a.gas0.s[1] : 0);
    col.Bk = have_k ? static_cast<const T*>(a.Bk.p) + (long long)mz * a.Bk.s[LA] + (ND == 3 ? (long long)my * a.Bk.s[1] : 0) : nullptr;
    col.vk = have_k ? static_cast<const T*>(a.vk.p) + (long long)mz * a.vk.s[LA] + (ND == 3 ? (long long)my * a.vk.s[1] : 0) : nullptr;
    col.f2 = ND == 3 ? fy : fz;  // the y axis is the row axis in 3-D and the lane axis in 2-D
  };
  Col own, halo;
  make_col(ty + YO, tz + 1, own);
  constexpr int NHALO = ND == 3 ? (WY * WZ - TY * TZ) : 2;
  int hpos = -1;
  if (tid < NHALO) {
    int j, k;
    if constexpr (ND == 3) {
      if (tid < 2 * WZ) { j = tid < WZ ? 0 : TY + 1; k = tid % WZ; }
      else { const int h2 = tid - 2 * WZ; j = 1 + h2 / 2; k = (h2 & 1) ? TZ + 1 : 0; }
    } else {
      j = 0;
      k = tid == 0 ? 0 : TZ + 1;
    }
    hpos = j * WZ + k;
    make_col(j, k, halo);
  }
  const int gz = z0 + tz, gy = y0 + ty;
  const bool col_active = gz < g.n[LA] && (ND == 2 || gy < g.n[1]);
  const int wown = (ty + YO) * WZ + tz + 1;

  // stage one cell of plane p (asynchronously)
  auto issue_cell = [&](const Col& col, int mx, T* dst) {
    const T* pb = col.B0 + (long long)mx * a.B0.s[0];
    const T* pg = col.gas + (long long)mx * a.gas0.s[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) cp_async_elem(dst + k * PL, pb + k * a.B0.vs);
#pragma unroll
    for (int k = 0; k < ND; ++k) cp_async_elem(dst + (3 + k) * PL, pg + (k + 1) * a.gas0.vs);
    cp_async_elem(dst + 6 * PL, pg);
    if (have_k) {
      const T* pbk = col.Bk + (long long)mx * a.Bk.s[0];
      const T* pvk = col.vk + (long long)mx * a.vk.s[0];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        cp_async_elem(dst + (7 + k) * PL, pbk + k * a.Bk.vs);
        cp_async_elem(dst + (10 + k) * PL, pvk + k * a.vk.vs);
      }
    }
  };
  // planes beyond xe + 1 are never published; their group is committed empty
  auto issue_plane = [&](int p) {
    if (p <= xe) {
      bool fx;
      const int mx = map_index<ND>(p, g.n[0], g.bc[0][0], g.bc[0][1], fx);
      T* st = stage + ((p - xs + 1) & 3) * NRAW * PL;
      issue_cell(own, mx, st + wown);
      if (hpos >= 0) issue_cell(halo, mx, st + hpos);
    }
    cp_async_commit();
  };
  // argument (R_k + R_0)/2 or R_k and E = B x v of one staged cell -> ring
  auto publish_cell = [&](const T* src, bool fx, bool f2, T* wdst) {
    T B[3], v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const T b0 = src[k * PL];
      const T v0 = (k < ND) ? src[(3 + k) * PL] : T(0);
      if (have_k) {
        const T bk = src[(7 + k) * PL], vk = src[(10 + k) * PL];
        if (midpoint) { B[k] = (bk + b0) / T(2); v[k] = (vk + v0) / T(2); }
        else { B[k] = bk; v[k] = vk; }
      } else {
        B[k] = b0;
        v[k] = v0;
      }
    }
    if (fx) { B[1] = -B[1]; v[1] = -v[1]; }
    if (f2) { B[2] = -B[2]; v[2] = -v[2]; }
    wdst[0 * PL] = B[0];
    wdst[1 * PL] = B[1];
    wdst[2 * PL] = B[2];
    wdst[3 * PL] = B[1] * v[2] - B[2] * v[1];
    wdst[4 * PL] = B[2] * v[0] - B[0] * v[2];
    wdst[5 * PL] = B[0] * v[1] - B[1] * v[0];
  };
  auto publish_plane = [&](int p) {
    bool fx;
    (void)map_index<ND>(p, g.n[0], g.bc[0][0], g.bc[0][1], fx);
    const int slot = (p - xs + 1) & 3;
    const T* st = stage + slot * NRAW * PL;
    T* w = ring + slot * 6 * PL;
    publish_cell(st + wown, fx, own.f2, w + wown);
    if (hpos >= 0) publish_cell(st + hpos, fx, halo.f2, w + hpos);
  };

  T ch[2] = {T(0), T(0)};
  T* const pB1 = static_cast<T*>(a.Bk1.p) + (long long)gz * a.Bk1.s[LA] + (ND == 3 ? (long long)gy * a.Bk1.s[1] : 0);
  T* const pv1 = static_cast<T*>(a.vk1.p) + (long long)gz * a.vk1.s[LA] + (ND == 3 ? (long long)gy * a.vk1.s[1] : 0);
#if AMX_STRICT
  const T hdx = c.dx;
#else
  const T hdx = T(0.5) * c.inv_dx;
#endif
  // central difference 0.5*(fp - fm)/dx (_vector_maths.py:55-100)
  auto cd = [&](T fp, T fm) -> T {
#if AMX_STRICT
    return (T(0.5) * (fp - fm)) / hdx;
#else
    return hdx * (fp - fm);
#endif
  };
  // update of plane x from ring slots of planes x-1, x, x+1 and the staged raw centre values
  auto update = [&](int x) {
    const int sC = (x - xs + 1) & 3, sM = (sC + 3) & 3, sP = (sC + 1) & 3;
    const T* wm = ring + sM * 6 * PL + wown;
    const T* wc = ring + sC * 6 * PL + wown;
    const T* wp = ring + sP * 6 * PL + wown;
    const T* rw = stage + sC * NRAW * PL + wown;
    auto ddx = [&](int f) { return cd(wp[f * PL], wm[f * PL]); };
    auto ddl = [&](int f) { return cd(wc[f * PL + 1], wc[f * PL - 1]); };
    auto ddy = [&](int f) { return cd(wc[f * PL + WZ], wc[f * PL - WZ]); };
    T curlE[3], curlB[3];
    if constexpr (ND == 3) {  // curl3D: lane axis is z
      curlB[0] = ddy(2) - ddl(1); curlB[1] = ddl(0) - ddx(2); curlB[2] = ddx(1) - ddy(0);
      curlE[0] = ddy(5) - ddl(4); curlE[1] = ddl(3) - ddx(5); curlE[2] = ddx(4) - ddy(3);
    } else {  // curl2D: lane axis is y
      curlB[0] = ddl(2); curlB[1] = -ddx(2); curlB[2] = ddx(1) - ddl(0);
      curlE[0] = ddl(5); curlE[1] = -ddx(5); curlE[2] = ddx(4) - ddl(3);
    }
    const T Bc[3] = {wc[0], wc[PL], wc[2 * PL]};
    T cr[3];
    cr[0] = Bc[1] * curlB[2] - Bc[2] * curlB[1];
    cr[1] = Bc[2] * curlB[0] - Bc[0] * curlB[2];
    cr[2] = Bc[0] * curlB[1] - Bc[1] * curlB[0];
    const T rho = rw[6 * PL];
#if !AMX_STRICT
    const T ir = rdiv(T(1), rho);
#endif
    T dB = T(0), dv = T(0);
    T* qB = pB1 + (long long)x * a.Bk1.s[0];
    T* qv = pv1 + (long long)x * a.vk1.s[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#if AMX_STRICT
      const T phiB = cr[k] / rho;
#else
      const T phiB = cr[k] * ir;
#endif
      const T b0 = rw[k * PL];
      const T v0 = (k < ND) ? rw[(3 + k) * PL] : T(0);
      const T Bn = b0 - dt * curlE[k];
      const T vn = v0 - dt * phiB;
      const T bk = have_k ? rw[(7 + k) * PL] : b0;
      const T vk = have_k ? rw[(10 + k) * PL] : v0;
      dB = dB + sq(rabs(bk - Bn));
      dv = dv + sq(rabs(vk - vn));
      qB[k * a.Bk1.vs] = Bn;
      qv[k * a.vk1.vs] = vn;
    }
    ch[0] = rmax(ch[0], dB);
    ch[1] = rmax(ch[1], dv);
  };

  // ---- prologue: planes xs-1 .. xs+2 in flight, (B, E) of xs-1, xs, xs+1 published
  issue_plane(xs - 1);
  issue_plane(xs);
  issue_plane(xs + 1);
  issue_plane(xs + 2);
  cp_async_wait<2>();
  publish_plane(xs - 1);
  publish_plane(xs);
  // stage slot of plane xs-1 is free again (its raw values are never needed)
  issue_plane(xs + 3);
  cp_async_wait<2>();
  publish_plane(xs + 1);
  // ---- steady state: at step x the groups of planes x+2 and x+3 are outstanding
  for (int x = xs; x < xe; ++x) {
    cp_async_wait<1>();  // plane x+2 has landed (own entries)
    __syncthreads();     // ring(x+1) visible to all; every thread is done with update(x-1)
    if (x + 2 <= xe) publish_plane(x + 2);
    if (col_active) update(x);
    issue_plane(x + 4);  // into the stage slot of plane x, just consumed by this thread
  }
  cp_async_wait<0>();
  ch[0] = rsqrt_(ch[0]);  // sqrt is monotone: max of norms = norm of the max of squares
  ch[1] = rsqrt_(ch[1]);
  {
    __shared__ T red[2][8];
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ch[k] = rmax(ch[k], __shfl_xor_sync(0xffffffffu, ch[k], o));
      if (lane == 0) red[k][wid] = ch[k];
    }
    __syncthreads();
    if (wid == 0) {
      const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        T m = lane < 8 ? red[k][lane] : T(0);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = rmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(a.change_slots + (size_t)k * kAlphaSlots + (bid % kAlphaSlots), bits_of(m));
      }
    }
  }
}

template <typename T, int ND>
cudaError_t launch_mag(const MagIterArgs& a, cudaStream_t st) {
  constexpr int TY = Tile<ND>::TY, TZ = Tile<ND>::TZ;
  const size_t smem = mag_smem_bytes<T, ND>();
  auto kern = mag_iter_tiled<T, ND>;
  static SmemOptIn opt;
  if (opt.first_time()) {
    cudaError_t e = opt.ensure(kern, smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  dim3 grid;
  const int LA = ND - 1;
  grid.x = (a.g.n[LA] + TZ - 1) / TZ;
  const long long tiles = ND == 3 ? (long long)grid.x * ((a.g.n[1] + TY - 1) / TY) : grid.x;
  const int xchunk = pick_xchunk(a.g.n[0], tiles);
  const int nch = (a.g.n[0] + xchunk - 1) / xchunk;
  if (ND == 3) { grid.y = (a.g.n[1] + TY - 1) / TY; grid.z = nch; }
  else { grid.y = nch; grid.z = 1; }
  kern<<<grid, 256, smem, st>>>(a, xchunk);
  return cudaGetLastError();
}

}  // namespace tiled

// The tiled stage handles 2-D / 3-D grids whose lane axis is contiguous and at least one tile
// wide; everything else (1-D, tiny grids) stays on the generic kernel.
bool tiled_stage_supported(const StageArgs& a) {
  if (a.g.jet_hi > a.g.jet_lo) return false;  // MHD jet inflow cells: generic kernels only
  if (a.np.lim == LIM_VAN_ALBADA_PP) return false;  // multi-axis limiter: generic kernel (amx_kernels.cu)
  if (a.np.rs > RS_LF) return false;                 // HYBRID / AM-HLLC need the compression switch of the cell
  if (a.g.nd == 3)
    return a.in.s[2] == 1 && a.out.s[2] == 1 && a.g.n[2] >= tiled::Tile<3>::TZ && a.g.n[1] >= tiled::Tile<3>::TY + 2 &&
           a.g.n[0] >= 4 && (!a.in0.p || a.in0.s[2] == 1);
  if (a.g.nd == 2)
    return a.in.s[1] == 1 && a.out.s[1] == 1 && a.g.n[1] >= tiled::Tile<2>::TZ && a.g.n[0] >= 4 &&
           (!a.in0.p || a.in0.s[1] == 1);
  return false;
}

template <typename T>
static cudaError_t launch_tiled_t(const StageArgs& a, cudaStream_t st) {
#define AMX_TILED_CASE(ND_, RS_)                                                               \
  if (a.g.nd == ND_ && a.np.rs == RS_)                                                         \
    return a.np.lim == LIM_MINMOD ? tiled::launch_stage<T, ND_, RS_, LIM_MINMOD>(a, st)        \
                                  : tiled::launch_stage<T, ND_, RS_, -1>(a, st);
  AMX_TILED_CASE(3, RS_HLL) AMX_TILED_CASE(3, RS_HLLC) AMX_TILED_CASE(3, RS_HLLC_LM) AMX_TILED_CASE(3, RS_LF)
  AMX_TILED_CASE(2, RS_HLL) AMX_TILED_CASE(2, RS_HLLC) AMX_TILED_CASE(2, RS_HLLC_LM) AMX_TILED_CASE(2, RS_LF)
#undef AMX_TILED_CASE
  return cudaErrorInvalidValue;
}

bool tiled_mag_supported(const MagIterArgs& a) {
  if (a.g.jet_hi > a.g.jet_lo) return false;  // MHD jet inflow cells: generic kernels only
  const int la = a.g.nd - 1;
  if (a.g.nd != 2 && a.g.nd != 3) return false;
  const bool unit = a.B0.s[la] == 1 && a.gas0.s[la] == 1 && a.Bk1.s[la] == 1 && a.vk1.s[la] == 1 &&
                    (!a.Bk.p || (a.Bk.s[la] == 1 && a.vk.s[la] == 1));
  if (a.g.nd == 3) return unit && a.g.n[2] >= tiled::Tile<3>::TZ && a.g.n[1] >= tiled::Tile<3>::TY + 2 && a.g.n[0] >= 4;
  return unit && a.g.n[1] >= tiled::Tile<2>::TZ && a.g.n[0] >= 4;
}

cudaError_t launch_mag_iter_tiled(const MagIterArgs& a, cudaStream_t st) {
  if (a.g.nd == 3) return a.np.dtype == 0 ? tiled::launch_mag<float, 3>(a, st) : tiled::launch_mag<double, 3>(a, st);
  if (a.g.nd == 2) return a.np.dtype == 0 ? tiled::launch_mag<float, 2>(a, st) : tiled::launch_mag<double, 2>(a, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch_hydro_stage_tiled(const StageArgs& a, cudaStream_t st) {
  return a.np.dtype == 0 ? launch_tiled_t<float>(a, st) : launch_tiled_t<double>(a, st);
}

}  // namespace AMX_NS
