// amx_kernels.cu -- sm_100a kernels of the finite-volume (M)HD step, generic ("v1") set:
// one thread per active cell, boundary conditions applied by index mapping on the halo
// reads (so no ghost cells have to exist in memory), every stage fused into one launch:
//   hydro_stage_kernel   BC -> cons -> per axis {limited slopes, L/R states, Riemann flux,
//                        flux difference} -> prim [-> RK2 combine] -> wave-bound max
//                        (reference: evolve_state.py:339-414, :456-469)
//   split_sweep_kernel   one MUSCL-Hancock sweep (evolve_state.py:53-163)
//   mag_iter_kernel      one evaluation of the Pang-Wu fixed point + change norms
//                        (_magnetic_field_update.py:82-165)
//   mag_final_kernel     velocity / pressure write-back (:208-240)
//   alpha_kernel         max |u|+c per axis (_timestep_estimator.py:118-128)
// The plane-streaming shared-memory kernels for the 3-D/2-D hot configurations live in
// amx_tiled.cu and are selected by the API layer when the layout allows.
//
// Compiled twice: -DAMX_STRICT=0 (namespace amx_fast) and -DAMX_STRICT=1 -fmad=false
// (namespace amx_strict).
#include "amx_internal.h"
#include "amx_math.cuh"

#include <algorithm>

#if AMX_STRICT
#define AMX_NS amx_strict
#else
#define AMX_NS amx_fast
#endif

namespace AMX_NS {
using namespace amx;

// ---------------------------------------------------------------- indexing
struct Idx3 { int i[3]; };

// ghost index -> source index for boundary code bc; flip = the normal component changes sign
template <int ND>
AMX_DI int map_index(int i, int n, int bcl, int bcr, bool& flip) {
  flip = false;
  if (i < 0) {
    if (bcl == BC_PERIODIC) return i + n;
    if (bcl == BC_OPEN || bcl == BC_JET) return 0;  // boundaries.py:407-410: the jet side itself is an open copy
    if (bcl == BC_REFLECTIVE) { flip = true; return (ND == 1) ? 0 : (-1 - i); }
    return i;  // BC_HALO: ghost planes are resident
  }
  if (i >= n) {
    if (bcr == BC_PERIODIC) return i - n;
    if (bcr == BC_OPEN) return n - 1;
    if (bcr == BC_REFLECTIVE) { flip = true; return (ND == 1) ? (n - 1) : (2 * n - 1 - i); }
    return i;
  }
  return i;
}

template <typename T, int ND>
AMX_DI Gas<T, ND> load_gas(const View& v, long long off) {
  Gas<T, ND> s;
  const T* p = static_cast<const T*>(v.p) + off;
#pragma unroll
  for (int k = 0; k < ND + 2; ++k) s.q[k] = __ldg(p + k * v.vs);
  return s;
}

template <typename T, int ND>
AMX_DI void store_gas(const View& v, long long off, const Gas<T, ND>& s) {
  T* p = static_cast<T*>(v.p) + off;
#pragma unroll
  for (int k = 0; k < ND + 2; ++k) p[k * v.vs] = s.q[k];
}

template <int ND>
AMX_DI long long offset_of(const View& v, const Idx3& c) {
  long long o = (long long)c.i[0] * v.s[0];
  if constexpr (ND > 1) o += (long long)c.i[1] * v.s[1];
  if constexpr (ND > 2) o += (long long)c.i[2] * v.s[2];
  return o;
}

// gas state of an MHD-jet inflow ghost cell (boundaries.py:418: rho, v_x, v_y, p = 5/3, 800, 0, 1)
template <typename T, int ND>
AMX_DI Gas<T, ND> jet_gas() {
  Gas<T, ND> s;
#pragma unroll
  for (int k = 0; k < ND + 2; ++k) s.q[k] = jet_gas_value<T>(k);
  return s;
}

// gas state of the cell `d` steps from c along axis a (0-based), boundary rules applied
template <typename T, int ND>
AMX_DI Gas<T, ND> load_neighbor(const View& v, const Geom& g, const Idx3& c, int a, int d) {
  Idx3 q = c;
  bool flip;
  if constexpr (ND == 2) {
    if (a == 0 && jet_cell(g, c.i[0] + d, c.i[1])) return jet_gas<T, ND>();
  }
  q.i[a] = map_index<ND>(c.i[a] + d, g.n[a], g.bc[a][0], g.bc[a][1], flip);
  Gas<T, ND> s = load_gas<T, ND>(v, offset_of<ND>(v, q));
  if (flip) s.q[a + 1] = -s.q[a + 1];
  return s;
}

template <int ND>
AMX_DI bool decode_cell(const Geom& g, Idx3& c, long long lin = -1) {
  if (lin < 0) lin = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = g.n[0];
  if constexpr (ND > 1) total *= g.n[1];
  if constexpr (ND > 2) total *= g.n[2];
  if (lin >= total) return false;
  long long r = lin;
  if constexpr (ND == 3) { c.i[2] = (int)(r % g.n[2]); r /= g.n[2]; }
  if constexpr (ND >= 2) { c.i[1] = (int)(r % g.n[1]); r /= g.n[1]; }
  c.i[0] = (int)r;
  if constexpr (ND < 3) c.i[2] = 0;
  if constexpr (ND < 2) c.i[1] = 0;
  return true;
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

// ---------------------------------------------------------------- block max -> slots
AMX_DI unsigned long long bits_of(float v) { return (unsigned long long)__float_as_uint(v); }
AMX_DI unsigned long long bits_of(double v) { return (unsigned long long)__double_as_longlong(v); }

// max over the block of NV non-negative values per thread, one atomicMax per value into
// slot (blockIdx mod kAlphaSlots).  All threads of the block must call.
template <typename T, int NV>
AMX_DI void block_max_to_slots(T (&v)[NV], unsigned long long* slots) {
  __shared__ T red[NV][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] = rmax(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
    if (lane == 0) red[k][wid] = v[k];
  }
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      T m = lane < nw ? red[k][lane] : T(0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = rmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) atomicMax(slots + (size_t)k * kAlphaSlots + (blockIdx.x % kAlphaSlots), bits_of(m));
    }
  }
}

template <typename T, int ND>
AMX_DI Gas<T, ND> load_at(const View& v, const Geom& g, const Idx3& pos);

// hll.py:283-296: g = 1 where div v < -0.05 c / dx (central differences of the ghost-filled state)
template <typename T, int ND>
AMX_DI T compression_switch(const View& v, const Geom& g, const Idx3& pos, const Gas<T, ND>& s0, const Consts<T>& c) {
  T div = T(0);
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    Idx3 m = pos, p = pos;
    m.i[i] -= 1;
    p.i[i] += 1;
    const T d = load_at<T, ND>(v, g, p).q[i + 1] - load_at<T, ND>(v, g, m).q[i + 1];
    const T term = d / (T(2) * c.dx);
    div = i == 0 ? term : div + term;
  }
  const T cs = sound_speed(s0.q[0], s0.q[ND + 1], c);
  return div < (T(-0.05) * cs) / c.dx ? T(1) : T(0);
}

// ---------------------------------------------------------------- unsplit stage
template <typename T, int ND, int RS, int LIM>
__global__ void __launch_bounds__(256) hydro_stage_kernel(const StageArgs a) {
  if (a.ctl_ro && a.ctl_ro->done) return;
  const Consts<T> c = make_consts<T>(a.np);
  Idx3 cell;
  const bool active = decode_cell<ND>(a.g, cell);
  T wb[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) wb[k] = T(0);
  if (active) {
    const T dt = *static_cast<const T*>(a.dt) * (T)a.dt_scale;
    const Gas<T, ND> sc = load_gas<T, ND>(a.in, offset_of<ND>(a.in, cell));
    Gas<T, ND> U = cons_from_prim(sc, c);
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      const Gas<T, ND> sm2 = load_neighbor<T, ND>(a.in, a.g, cell, ax, -2);
      const Gas<T, ND> sm1 = load_neighbor<T, ND>(a.in, a.g, cell, ax, -1);
      const Gas<T, ND> sp1 = load_neighbor<T, ND>(a.in, a.g, cell, ax, +1);
      const Gas<T, ND> sp2 = load_neighbor<T, ND>(a.in, a.g, cell, ax, +2);
      T alpha = T(0), alpha_r = T(0);
      if constexpr (RS == RS_LF) alpha = alpha_r = static_cast<const T*>(a.alpha_lf)[ax];
      if constexpr (RS == RS_HYBRID_HLLC || RS == RS_AM_HLLC) {
        // flux array entry i (interface i-1/2) takes the switch of cell i
        Idx3 pp = cell;
        pp.i[ax] += 1;
        alpha = compression_switch<T, ND>(a.in, a.g, cell, sc, c);
        alpha_r = compression_switch<T, ND>(a.in, a.g, pp, sp1, c);
      }
      Gas<T, ND> lo_m, hi_m, lo_c, hi_c, lo_p, hi_p;
      faces_unsplit(sm1, slopes<T, ND, LIM>(sm2, sm1, sc, a.np.lim, c), c, lo_m, hi_m);
      faces_unsplit(sc, slopes<T, ND, LIM>(sm1, sc, sp1, a.np.lim, c), c, lo_c, hi_c);
      faces_unsplit(sp1, slopes<T, ND, LIM>(sc, sp1, sp2, a.np.lim, c), c, lo_p, hi_p);
      const Gas<T, ND> Fl = riemann<T, ND, RS>(hi_m, lo_c, ax + 1, alpha, c);    // interface i-1/2
      const Gas<T, ND> Fr = riemann<T, ND, RS>(hi_c, lo_p, ax + 1, alpha_r, c);  // interface i+1/2
#pragma unroll
      for (int k = 0; k < ND + 2; ++k) U.q[k] = U.q[k] + (c.inv_dx * (Fl.q[k] - Fr.q[k])) * dt;
    }
    Gas<T, ND> s = prim_from_cons(U, c);
    if (a.in0.p) {  // RK2-SSP combine: prim(0.5*(cons(s0) + cons(s2)))
      const Gas<T, ND> s0 = load_gas<T, ND>(a.in0, offset_of<ND>(a.in0, cell));
      const Gas<T, ND> U0 = cons_from_prim(s0, c), U2 = cons_from_prim(s, c);
      Gas<T, ND> Um;
#pragma unroll
      for (int k = 0; k < ND + 2; ++k) Um.q[k] = T(0.5) * (U0.q[k] + U2.q[k]);
      s = prim_from_cons(Um, c);
    }
    store_gas<T, ND>(a.out, offset_of<ND>(a.out, cell), s);
    if (a.alpha_slots) {
#pragma unroll
      for (int k = 0; k < ND; ++k) wb[k] = wave_bound(s, k + 1, c);
    }
  }
  if (a.alpha_slots) block_max_to_slots<T, ND>(wb, a.alpha_slots);
}

// ---------------------------------------------------------------- unsplit stage, VAN_ALBADA_PP
// reconstruction.py:134-394: the van-Albada half-differences of ALL axes of a cell are scaled
// together so that every interface state keeps rho > 0, p > 0 (alpha_density, kappa_pressure per
// axis) and positive internal energy (beta, one per cell; needs the global wave bounds C_a =
// alpha_a / sum(alpha)).  The flux at an interface needs the scaled difference of the cells on
// both sides, and a cell's factors need its differences along every axis: a thread evaluates the
// factors of its own cell and of its 2*ND face neighbours (ghost positions included: their
// differences are those of the ghost-filled array, like the reference's).  Functional path of
// one thread per cell, (1 + 2 ND)^2 state loads, all of them L1 / L2 hits.
template <typename T, int ND>
AMX_DI Gas<T, ND> load_at(const View& v, const Geom& g, const Idx3& pos) {
  Idx3 q = pos;
  bool flip[ND];
  if constexpr (ND == 2) {
    if (jet_cell(g, pos.i[0], pos.i[1])) return jet_gas<T, ND>();
  }
#pragma unroll
  for (int k = 0; k < ND; ++k) q.i[k] = map_index<ND>(pos.i[k], g.n[k], g.bc[k][0], g.bc[k][1], flip[k]);
  Gas<T, ND> s = load_gas<T, ND>(v, offset_of<ND>(v, q));
#pragma unroll
  for (int k = 0; k < ND; ++k)
    if (flip[k]) s.q[k + 1] = -s.q[k + 1];
  return s;
}

// scaled half-differences of the cell at `pos` along axis `want` (0-based); s0 = its state
template <typename T, int ND>
AMX_DI Gas<T, ND> pp_half_difference(const View& v, const Geom& g, const Idx3& pos, const Gas<T, ND>& s0,
                                     int want, const T (&Cax)[ND], T q, const Consts<T>& c) {
  constexpr int P = ND + 1;
  const T eps = (T)1e-14, one_eps = (T)(1.0 + 1e-14);
  Gas<T, ND> d[ND];
#pragma unroll
  for (int a = 0; a < ND; ++a) {
    Idx3 m = pos, p = pos;
    m.i[a] -= 1;
    p.i[a] += 1;
    const Gas<T, ND> gsl = slopes<T, ND, LIM_VAN_ALBADA>(load_at<T, ND>(v, g, m), s0, load_at<T, ND>(v, g, p),
                                                        LIM_VAN_ALBADA, c);
#pragma unroll
    for (int k = 0; k < ND + 2; ++k) d[a].q[k] = (gsl.q[k] * c.dx) / T(2);
  }
  const T rho = s0.q[0], pr = s0.q[P];
  T alpha_d[ND], kappa[ND];
#pragma unroll
  for (int a = 0; a < ND; ++a) {
    const T dr = d[a].q[0], dp = d[a].q[P];
    alpha_d[a] = rabs(dr) > eps ? rmin(rho / (rabs(dr) * one_eps), T(1)) : T(1);
    kappa[a] = rabs(dp) > eps ? rmin(pr / (rabs(dp) * one_eps), T(1)) : T(1);
  }
  // A1 = sum_j (sum_i C_j alpha_j drho_j d_i v_j)^2, A2 = sum_i C_i sum_j (d_i v_j)^2 (index pairing
  // of the reference's broadcast), vsum = sum_i sum_j (d_i v_j)^2
  T A1 = T(0), A2 = T(0), vsum = T(0);
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    const T f = (Cax[j] * alpha_d[j]) * d[j].q[0];
    T acc = T(0);
#pragma unroll
    for (int i = 0; i < ND; ++i) acc = i == 0 ? f * d[i].q[1 + j] : acc + f * d[i].q[1 + j];
    A1 = j == 0 ? acc * acc : A1 + acc * acc;
  }
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    T acc = T(0);
#pragma unroll
    for (int j = 0; j < ND; ++j) acc = j == 0 ? sq(d[i].q[1 + j]) : acc + sq(d[i].q[1 + j]);
    A2 = i == 0 ? Cax[i] * acc : A2 + Cax[i] * acc;
    vsum = i == 0 ? acc : vsum + acc;
  }
  T beta = T(1);
  if (vsum > eps) {
    const T qm2 = q - T(2);
    const T num = ((qm2 * qm2) * rho) * pr;
    const T den = c.gm1 * (T(2) * A1 + (qm2 * sq(rho)) * A2);
    beta = rmin(rsqrt_(num / den), T(1));
  }
  Gas<T, ND> r = d[0];
#pragma unroll
  for (int a = 1; a < ND; ++a)
    if (a == want) r = d[a];
  T al = alpha_d[0], ka = kappa[0];
#pragma unroll
  for (int a = 1; a < ND; ++a)
    if (a == want) { al = alpha_d[a]; ka = kappa[a]; }
  r.q[0] = r.q[0] * al;
  r.q[P] = r.q[P] * ka;
#pragma unroll
  for (int k = 1; k <= ND; ++k) r.q[k] = r.q[k] * beta;
  return r;
}

template <typename T, int ND, int RS>
__global__ void __launch_bounds__(128) hydro_stage_pp_kernel(const StageArgs a) {
  if (a.ctl_ro && a.ctl_ro->done) return;
  const Consts<T> c = make_consts<T>(a.np);
  Idx3 cell;
  const bool active = decode_cell<ND>(a.g, cell);
  T wb[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) wb[k] = T(0);
  if (active) {
    const T dt = *static_cast<const T*>(a.dt) * (T)a.dt_scale;
    const T q = (T)a.np.pp_q;
    T alpha[ND], Cax[ND], asum = T(0);
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      alpha[k] = static_cast<const T*>(a.alpha_lf)[k];
      asum = k == 0 ? alpha[k] : asum + alpha[k];
    }
#pragma unroll
    for (int k = 0; k < ND; ++k) Cax[k] = alpha[k] / asum;
    const Gas<T, ND> sc = load_gas<T, ND>(a.in, offset_of<ND>(a.in, cell));
    Gas<T, ND> U = cons_from_prim(sc, c);
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      Idx3 pm = cell, pp = cell;
      pm.i[ax] -= 1;
      pp.i[ax] += 1;
      const Gas<T, ND> sm1 = load_at<T, ND>(a.in, a.g, pm), sp1 = load_at<T, ND>(a.in, a.g, pp);
      const Gas<T, ND> dm = pp_half_difference<T, ND>(a.in, a.g, pm, sm1, ax, Cax, q, c);
      const Gas<T, ND> dc = pp_half_difference<T, ND>(a.in, a.g, cell, sc, ax, Cax, q, c);
      const Gas<T, ND> dp = pp_half_difference<T, ND>(a.in, a.g, pp, sp1, ax, Cax, q, c);
      Gas<T, ND> hi_m, lo_c, hi_c, lo_p;
#pragma unroll
      for (int k = 0; k < ND + 2; ++k) {
        hi_m.q[k] = sm1.q[k] + dm.q[k];
        lo_c.q[k] = sc.q[k] - dc.q[k];
        hi_c.q[k] = sc.q[k] + dc.q[k];
        lo_p.q[k] = sp1.q[k] - dp.q[k];
      }
      T al = RS == RS_LF ? alpha[ax] : T(0), ar = al;
      if constexpr (RS == RS_HYBRID_HLLC || RS == RS_AM_HLLC) {  // compression switch of the cell right of the interface
        al = compression_switch<T, ND>(a.in, a.g, cell, sc, c);
        ar = compression_switch<T, ND>(a.in, a.g, pp, sp1, c);
      }
      const Gas<T, ND> Fl = riemann<T, ND, RS>(hi_m, lo_c, ax + 1, al, c);
      const Gas<T, ND> Fr = riemann<T, ND, RS>(hi_c, lo_p, ax + 1, ar, c);
#pragma unroll
      for (int k = 0; k < ND + 2; ++k) U.q[k] = U.q[k] + (c.inv_dx * (Fl.q[k] - Fr.q[k])) * dt;
    }
    Gas<T, ND> s = prim_from_cons(U, c);
    if (a.in0.p) {
      const Gas<T, ND> s0 = load_gas<T, ND>(a.in0, offset_of<ND>(a.in0, cell));
      const Gas<T, ND> U0 = cons_from_prim(s0, c), U2 = cons_from_prim(s, c);
      Gas<T, ND> Um;
#pragma unroll
      for (int k = 0; k < ND + 2; ++k) Um.q[k] = T(0.5) * (U0.q[k] + U2.q[k]);
      s = prim_from_cons(Um, c);
    }
    store_gas<T, ND>(a.out, offset_of<ND>(a.out, cell), s);
    if (a.alpha_slots) {
#pragma unroll
      for (int k = 0; k < ND; ++k) wb[k] = wave_bound(s, k + 1, c);
    }
  }
  if (a.alpha_slots) block_max_to_slots<T, ND>(wb, a.alpha_slots);
}

// ---------------------------------------------------------------- split MUSCL-Hancock sweep
// AX = 0-based sweep axis (compile time, so every per-axis index stays in registers)
template <typename T, int ND, int RS, int AX>
__global__ void __launch_bounds__(256) split_sweep_kernel(const SweepArgs a) {
  if (a.ctl_ro && a.ctl_ro->done) return;
  const Consts<T> c = make_consts<T>(a.np);
  Idx3 cell;
  const bool active = decode_cell<ND>(a.g, cell);
  T wb[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) wb[k] = T(0);
  if (active) {
    const T dt = *static_cast<const T*>(a.dt) * (T)a.dt_scale;
    const Gas<T, ND> sc = load_gas<T, ND>(a.in, offset_of<ND>(a.in, cell));
    Gas<T, ND> U = cons_from_prim(sc, c);
    const Gas<T, ND> sm1 = load_neighbor<T, ND>(a.in, a.g, cell, AX, -1);
    const Gas<T, ND> sp1 = load_neighbor<T, ND>(a.in, a.g, cell, AX, +1);
    T alpha = T(0), alpha_r = T(0);
    if constexpr (RS == RS_LF) alpha = alpha_r = static_cast<const T*>(a.alpha_lf)[AX];
    if constexpr (RS == RS_HYBRID_HLLC || RS == RS_AM_HLLC) {
      // _riemann_solver.py:37-87 hands the sweep's (ghost-filled) input state to the hybrid solvers: flux
      // array entry i (interface i-1/2) takes the compression switch of cell i, as in the unsplit stage
      Idx3 pp = cell;
      pp.i[AX] += 1;
      alpha = compression_switch<T, ND>(a.in, a.g, cell, sc, c);
      alpha_r = compression_switch<T, ND>(a.in, a.g, pp, sp1, c);
    }
    Gas<T, ND> Fl, Fr;
    if (a.first_order) {  // evolve_state.py:72-74
      Fl = riemann<T, ND, RS>(sm1, sc, AX + 1, alpha, c);
      Fr = riemann<T, ND, RS>(sc, sp1, AX + 1, alpha_r, c);
    } else {
      const Gas<T, ND> sm2 = load_neighbor<T, ND>(a.in, a.g, cell, AX, -2);
      const Gas<T, ND> sp2 = load_neighbor<T, ND>(a.in, a.g, cell, AX, +2);
      const T hdt = dt / T(2);
      Gas<T, ND> lo_m, hi_m, lo_c, hi_c, lo_p, hi_p;
      faces_muscl(sm1, slopes<T, ND, -1>(sm2, sm1, sc, a.np.lim, c), AX + 1, hdt, c, lo_m, hi_m);
      faces_muscl(sc, slopes<T, ND, -1>(sm1, sc, sp1, a.np.lim, c), AX + 1, hdt, c, lo_c, hi_c);
      faces_muscl(sp1, slopes<T, ND, -1>(sc, sp1, sp2, a.np.lim, c), AX + 1, hdt, c, lo_p, hi_p);
      Fl = riemann<T, ND, RS>(hi_m, lo_c, AX + 1, alpha, c);
      Fr = riemann<T, ND, RS>(hi_c, lo_p, AX + 1, alpha_r, c);
    }
#pragma unroll
    for (int k = 0; k < ND + 2; ++k) U.q[k] = U.q[k] + (c.inv_dx * (Fl.q[k] - Fr.q[k])) * dt;
    const Gas<T, ND> s = prim_from_cons(U, c);
    store_gas<T, ND>(a.out, offset_of<ND>(a.out, cell), s);
    if (a.alpha_slots) {
#pragma unroll
      for (int k = 0; k < ND; ++k) wb[k] = wave_bound(s, k + 1, c);
    }
  }
  if (a.alpha_slots) block_max_to_slots<T, ND>(wb, a.alpha_slots);
}

// ---------------------------------------------------------------- wave bounds of a state
template <typename T, int ND>
__global__ void __launch_bounds__(256) alpha_kernel(const AlphaArgs a) {
  if (a.ctl_ro && a.ctl_ro->done) return;
  const Consts<T> c = make_consts<T>(a.np);
  Idx3 cell;
  const bool active = decode_cell<ND>(a.g, cell);
  T wb[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) wb[k] = T(0);
  if (active) {
    const Gas<T, ND> s = load_gas<T, ND>(a.gas, offset_of<ND>(a.gas, cell));
#pragma unroll
    for (int k = 0; k < ND; ++k) wb[k] = wave_bound(s, k + 1, c);
  }
  block_max_to_slots<T, ND>(wb, a.alpha_slots);
}

// ---------------------------------------------------------------- magnetic fixed point
template <typename T> struct V3 { T x[3]; };

// _vector_maths.py:6-12
template <typename T>
AMX_DI T cross_c(const V3<T>& a, const V3<T>& b, int comp) {
  if (comp == 0) return a.x[1] * b.x[2] - a.x[2] * b.x[1];
  if (comp == 1) return a.x[2] * b.x[0] - a.x[0] * b.x[2];
  return a.x[0] * b.x[1] - a.x[1] * b.x[0];
}

// B and v of the fixed-point argument at one (mapped) cell: (R_k + R_0)/2 or R_k.
// Reflective sides flip 3-vector component `axis` (1-based) -- i.e. the wrong one, and none
// for the z axis -- exactly like the reference (Q12).
template <typename T, int ND>
AMX_DI void load_arg(const MagIterArgs& a, const Idx3& c, int ax, int d, V3<T>& B, V3<T>& v) {
  Idx3 q = c;
  bool flip = false;
  if constexpr (ND == 2) {
    // MHD jet inflow cell: every iterate is ghost-filled with the same constants, so is their average
    if (d != 0 && ax == 0 && jet_cell(a.g, c.i[0] + d, c.i[1])) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { B.x[k] = jet_vector_value<T>(2, k); v.x[k] = jet_vector_value<T>(1, k); }
      return;
    }
  }
  if (d != 0) q.i[ax] = map_index<ND>(c.i[ax] + d, a.g.n[ax], a.g.bc[ax][0], a.g.bc[ax][1], flip);
  const int fc = ax + 1;  // flipped component index
  const long long o0 = offset_of<ND>(a.B0, q), og = offset_of<ND>(a.gas0, q);
  const T* pB0 = static_cast<const T*>(a.B0.p) + o0;
  const T* pg = static_cast<const T*>(a.gas0.p) + og;
  V3<T> B0, v0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    B0.x[k] = __ldg(pB0 + k * a.B0.vs);
    v0.x[k] = (k < ND) ? __ldg(pg + (k + 1) * a.gas0.vs) : T(0);
  }
  if (flip && fc < 3) { B0.x[fc] = -B0.x[fc]; v0.x[fc] = -v0.x[fc]; }
  if (a.Bk.p == nullptr) {  // first iterate: (R_0 + R_0)/2 == R_0 exactly
    B = B0; v = v0;
    return;
  }
  const long long ok = offset_of<ND>(a.Bk, q);
  const T* pBk = static_cast<const T*>(a.Bk.p) + ok;
  const T* pvk = static_cast<const T*>(a.vk.p) + offset_of<ND>(a.vk, q);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T bk = __ldg(pBk + k * a.Bk.vs), vk = __ldg(pvk + k * a.vk.vs);
    if (flip && k == fc) { bk = -bk; vk = -vk; }
    if (a.midpoint) { B.x[k] = (bk + B0.x[k]) / T(2); v.x[k] = (vk + v0.x[k]) / T(2); }
    else { B.x[k] = bk; v.x[k] = vk; }
  }
}

// central difference 0.5*(f[+1]-f[-1])/dx (_vector_maths.py:55-100)
template <typename T>
AMX_DI T cdiff(T fp, T fm, const Consts<T>& c) {
#if AMX_STRICT
  return (T(0.5) * (fp - fm)) / c.dx;
#else
  return (T(0.5) * (fp - fm)) * c.inv_dx;
#endif
}

template <typename T, int ND>
__global__ void __launch_bounds__(256) mag_iter_kernel(const MagIterArgs a) {
  if (a.ctl_ro && (a.ctl_ro->done || (!a.first && !a.ctl_ro->mag_active))) return;
  const Consts<T> c = make_consts<T>(a.np);
  Idx3 cell;
  const bool active = decode_cell<ND>(a.g, cell);
  T ch[2] = {T(0), T(0)};
  if (active) {
    const T dt = *static_cast<const T*>(a.dt);
    V3<T> Bc, vc;
    load_arg<T, ND>(a, cell, 0, 0, Bc, vc);
    // E = B x v and B at the two neighbours of every axis
    T dE[3][3], dB[3][3];  // [axis][component] central differences
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      V3<T> Bp, vp, Bm, vm;
      load_arg<T, ND>(a, cell, ax, +1, Bp, vp);
      load_arg<T, ND>(a, cell, ax, -1, Bm, vm);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dE[ax][k] = cdiff(cross_c(Bp, vp, k), cross_c(Bm, vm, k), c);
        dB[ax][k] = cdiff(Bp.x[k], Bm.x[k], c);
      }
    }
    V3<T> curlE, curlB;
    if constexpr (ND == 3) {  // curl3D
      curlE.x[0] = dE[1][2] - dE[2][1]; curlE.x[1] = dE[2][0] - dE[0][2]; curlE.x[2] = dE[0][1] - dE[1][0];
      curlB.x[0] = dB[1][2] - dB[2][1]; curlB.x[1] = dB[2][0] - dB[0][2]; curlB.x[2] = dB[0][1] - dB[1][0];
    } else {  // curl2D
      curlE.x[0] = dE[1][2]; curlE.x[1] = -dE[0][2]; curlE.x[2] = dE[0][1] - dE[1][0];
      curlB.x[0] = dB[1][2]; curlB.x[1] = -dB[0][2]; curlB.x[2] = dB[0][1] - dB[1][0];
    }
    // R_0 at the centre and rho
    const long long o0 = offset_of<ND>(a.B0, cell), og = offset_of<ND>(a.gas0, cell);
    const T* pB0 = static_cast<const T*>(a.B0.p) + o0;
    const T* pg = static_cast<const T*>(a.gas0.p) + og;
    const T rho = __ldg(pg);
    T Bn[3], vn[3], dBk = T(0), dvk = T(0);
    const long long ob = offset_of<ND>(a.Bk1, cell), ov = offset_of<ND>(a.vk1, cell);
    T* pB1 = static_cast<T*>(a.Bk1.p) + ob;
    T* pv1 = static_cast<T*>(a.vk1.p) + ov;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const T B0k = __ldg(pB0 + k * a.B0.vs);
      const T v0k = (k < ND) ? __ldg(pg + (k + 1) * a.gas0.vs) : T(0);
      const T phiB = rdiv(cross_c(Bc, curlB, k), rho);
      Bn[k] = B0k - dt * curlE.x[k];
      vn[k] = v0k - dt * phiB;
      T bk = B0k, vk = v0k;  // R_k at the centre (R_0 on the first iterate)
      if (a.Bk.p) {
        bk = __ldg(static_cast<const T*>(a.Bk.p) + offset_of<ND>(a.Bk, cell) + k * a.Bk.vs);
        vk = __ldg(static_cast<const T*>(a.vk.p) + offset_of<ND>(a.vk, cell) + k * a.vk.vs);
      }
      dBk = dBk + sq(rabs(bk - Bn[k]));
      dvk = dvk + sq(rabs(vk - vn[k]));
      pB1[k * a.Bk1.vs] = Bn[k];
      pv1[k * a.vk1.vs] = vn[k];
    }
    ch[0] = rsqrt_(dBk);
    ch[1] = rsqrt_(dvk);
  }
  block_max_to_slots<T, 2>(ch, a.change_slots);
}

// _magnetic_field_update.py:208-240 (gamma hard-coded 5/3, Q4; 2-D drops v_z, Q5)
template <typename T, int ND>
__global__ void __launch_bounds__(256) mag_final_kernel(const MagFinalArgs a) {
  // grid-stride: a launch that finds its work pre-empted (final_emitted) must cost a few blocks, not 65536
  if (a.ctl_ro && (a.ctl_ro->done || (a.skip_if_emitted && a.ctl_ro->final_emitted))) return;
  if (a.final_runs && blockIdx.x == 0 && threadIdx.x == 0) *a.final_runs += 1;
  const Consts<T> c = make_consts<T>(a.np);
  T wb[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) wb[k] = T(0);
  long long total = a.g.n[0];
  if constexpr (ND > 1) total *= a.g.n[1];
  if constexpr (ND > 2) total *= a.g.n[2];
  const int fin = a.fixed_final >= 0 ? a.fixed_final : a.ctl_ro->mag_final;
  const View& X = a.X[fin];
  for (long long lin = (long long)blockIdx.x * blockDim.x + threadIdx.x; lin < total;
       lin += (long long)gridDim.x * blockDim.x) {
    Idx3 cell;
    decode_cell<ND>(a.g, cell, lin);
    const long long ox = offset_of<ND>(X, cell), og = offset_of<ND>(a.gas, cell), ob = offset_of<ND>(a.B_out, cell);
    const long long oo = offset_of<ND>(a.gas_out, cell);
    const T* px = static_cast<const T*>(X.p) + ox;
    T* po = static_cast<T*>(a.gas_out.p) + oo;
    T* pb = static_cast<T*>(a.B_out.p) + ob;
    Gas<T, ND> s = load_gas<T, ND>(a.gas, og);
    T n0 = T(0), nn = T(0), vn[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pb[k * a.B_out.vs] = px[k * X.vs];
      vn[k] = px[(3 + k) * X.vs];
      const T v0k = (k < ND) ? s.q[k + 1] : T(0);
      n0 = n0 + sq(rabs(v0k));
      nn = nn + sq(rabs(vn[k]));
    }
    const T p2 = magnetic_writeback_pressure(s.q[0], s.q[ND + 1], n0, nn);
    if (a.gas_out.p != a.gas.p) po[0] = s.q[0];
#pragma unroll
    for (int k = 0; k < ND; ++k) { s.q[k + 1] = vn[k]; po[(k + 1) * a.gas_out.vs] = vn[k]; }
    s.q[ND + 1] = p2;
    po[(ND + 1) * a.gas_out.vs] = p2;
    if (a.peers.flags) {  // slab rank: edge planes of gas_out / B_out -> the neighbours' ghost planes
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        long long bd;
        int ps;
        if (peer_target(a.peers, cell.i[0], a.g.n[0], side, bd, ps)) {
          T* qo = reinterpret_cast<T*>(reinterpret_cast<char*>(po) + bd) + (long long)ps * a.gas_out.s[0];
          T* qb = reinterpret_cast<T*>(reinterpret_cast<char*>(pb) + bd) + (long long)ps * a.B_out.s[0];
          qo[0] = s.q[0];
#pragma unroll
          for (int k = 0; k < ND; ++k) qo[(k + 1) * a.gas_out.vs] = vn[k];
          qo[(ND + 1) * a.gas_out.vs] = p2;
#pragma unroll
          for (int k = 0; k < 3; ++k) qb[k * a.B_out.vs] = px[k * X.vs];
        }
      }
    }
    if (a.alpha_slots) {
#pragma unroll
      for (int k = 0; k < ND; ++k) wb[k] = rmax(wb[k], wave_bound(s, k + 1, c));
    }
  }
  if (a.alpha_slots) block_max_to_slots<T, ND>(wb, a.alpha_slots);
}

// ---------------------------------------------------------------- ghost shell of the gas state
// One thread per cell of the PADDED array; interior cells only take part in the block maximum.
// A ghost cell g has the interior source cell src(g) (index mapping per axis, the composition of
// the reference's axis-by-axis fills) and, per reflective axis it lies beyond, one sign flip.
//   mode 0  evolve_state.py:456-469 at a ghost cell: prim(0.5*(cons(G) + cons(fill(S2)))) with
//           cons(S2) = 2 cons(U1) - cons(U) at src(g) (U -> U1 is the interior RK2 combine) and the
//           gas rule: momentum `axis` changes sign
//   mode 1  _magnetic_field_update.py:208-240 at a ghost cell: rho, p of the shell; |v_0|, |v_n| of
//           the source cell; velocity rows <- v_n with the 3-vector rule (component index `axis`,
//           1-based, changes sign if it exists: Q12)
//   mode 2  nothing changes; wave bounds only
// then |v_axis| + c of the shell cell joins the wave-bound maximum (_timestep_estimator.py:118-128
// runs over the padded array).
template <typename T, int ND>
__global__ void __launch_bounds__(256) shell_kernel(const ShellArgs a) {
  if (a.ctl_ro && a.ctl_ro->done) return;
  const Consts<T> c = make_consts<T>(a.np);
  T wb[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) wb[k] = T(0);
  long long e[3], tot = 1;
#pragma unroll
  for (int k = 0; k < 3; ++k) { e[k] = k < ND ? a.g.n[k] + 2 * a.ng : 1; tot *= e[k]; }
  const long long lin = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool ghost = false, halo = false;
  if (lin < tot) {
    long long r = lin;
    Idx3 gi, si;
    gi.i[2] = (int)(r % e[2]); r /= e[2];
    gi.i[1] = (int)(r % e[1]); r /= e[1];
    gi.i[0] = (int)r;
    bool flip[3] = {false, false, false};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < ND) {
        gi.i[k] -= a.ng;
        si.i[k] = map_index<ND>(gi.i[k], a.g.n[k], a.g.bc[k][0], a.g.bc[k][1], flip[k]);
        ghost = ghost || gi.i[k] < 0 || gi.i[k] >= a.g.n[k];
        // slab rank: the planes beyond a BC_HALO side are the neighbour's cells (and its shell), not ghost cells
        if ((gi.i[k] < 0 && a.g.bc[k][0] == BC_HALO) || (gi.i[k] >= a.g.n[k] && a.g.bc[k][1] == BC_HALO)) halo = true;
      } else {
        gi.i[k] = 0;
        si.i[k] = 0;
      }
    }
    if (halo) ghost = false;
    if (ghost) {
      bool jet = false;  // MHD jet inflow cell: its fills are constants, not images
      if constexpr (ND == 2) jet = jet_cell(a.g, gi.i[0], gi.i[1]);
      Gas<T, ND> G = load_gas<T, ND>(a.shell_in, offset_of<ND>(a.shell_in, gi));
      if (a.mode == 0) {
        const Gas<T, ND> UG = cons_from_prim(G, c);
        Gas<T, ND> Um;
        if (jet) {
          const Gas<T, ND> UJ = cons_from_prim(jet_gas<T, ND>(), c);
#pragma unroll
          for (int k = 0; k < ND + 2; ++k) Um.q[k] = T(0.5) * (UG.q[k] + UJ.q[k]);
        } else {
          const Gas<T, ND> U0 = cons_from_prim(load_gas<T, ND>(a.u0, offset_of<ND>(a.u0, si)), c);
          const Gas<T, ND> U1 = cons_from_prim(load_gas<T, ND>(a.u1, offset_of<ND>(a.u1, si)), c);
#pragma unroll
          for (int k = 0; k < ND + 2; ++k) {
            T s2 = T(2) * U1.q[k] - U0.q[k];
            if (k >= 1 && k <= ND && flip[k - 1]) s2 = -s2;
            Um.q[k] = T(0.5) * (UG.q[k] + s2);
          }
        }
        G = prim_from_cons(Um, c);
      } else if (a.mode == 3) {
        if (jet) {
          G = jet_gas<T, ND>();
        } else {
          G = load_gas<T, ND>(a.u1, offset_of<ND>(a.u1, si));
#pragma unroll
          for (int k = 0; k < ND; ++k)
            if (flip[k]) G.q[k + 1] = -G.q[k + 1];
        }
      } else if (a.mode == 1) {
        const int fin = a.fixed_final >= 0 ? a.fixed_final : a.ctl_ro->mag_final;
        const View& X = a.X[fin];
        const T* px = static_cast<const T*>(X.p) + offset_of<ND>(X, si);
        const Gas<T, ND> s0 = load_gas<T, ND>(a.u0, offset_of<ND>(a.u0, si));
        T n0 = T(0), nn = T(0), vn[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          vn[k] = jet ? jet_vector_value<T>(1, k) : __ldg(px + (3 + k) * X.vs);
          const T v0k = jet ? jet_vector_value<T>(1, k) : ((k < ND) ? s0.q[k + 1] : T(0));
          n0 = n0 + sq(rabs(v0k));
          nn = nn + sq(rabs(vn[k]));
        }
        const T p2 = magnetic_writeback_pressure(G.q[0], G.q[ND + 1], n0, nn);
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          T v = vn[k];
          // the fill of axis k2 flips 3-vector component k2 + 1
          if (k >= 1 && flip[k - 1] && !jet) v = -v;
          G.q[k + 1] = v;
        }
        G.q[ND + 1] = p2;
      }
      if (a.mode != 2 || a.shell_out.p != a.shell_in.p) store_gas<T, ND>(a.shell_out, offset_of<ND>(a.shell_out, gi), G);
      if (a.alpha_slots) {
#pragma unroll
        for (int k = 0; k < ND; ++k) wb[k] = wave_bound(G, k + 1, c);
      }
    }
  }
  if (a.alpha_slots) block_max_to_slots<T, ND>(wb, a.alpha_slots);
}

// ---------------------------------------------------------------- launchers
template <int ND>
static inline unsigned grid_for(const Geom& g) {
  long long total = g.n[0];
  if (ND > 1) total *= g.n[1];
  if (ND > 2) total *= g.n[2];
  return (unsigned)((total + 255) / 256);
}

#define AMX_DISPATCH_ND(ND_, ...)            \
  switch (ND_) {                              \
    case 1: { constexpr int ND = 1; __VA_ARGS__ } break; \
    case 2: { constexpr int ND = 2; __VA_ARGS__ } break; \
    case 3: { constexpr int ND = 3; __VA_ARGS__ } break; \
    default: return cudaErrorInvalidValue;    \
  }

#define AMX_DISPATCH_RS(RS_, ...)                         \
  switch (RS_) {                                           \
    case RS_HLL: { constexpr int RS = RS_HLL; __VA_ARGS__ } break;        \
    case RS_HLLC: { constexpr int RS = RS_HLLC; __VA_ARGS__ } break;      \
    case RS_HLLC_LM: { constexpr int RS = RS_HLLC_LM; __VA_ARGS__ } break; \
    case RS_LF: { constexpr int RS = RS_LF; __VA_ARGS__ } break;          \
    default: return cudaErrorInvalidValue;                 \
  }

template <typename T>
static cudaError_t launch_hydro_stage_pp_t(const StageArgs& a, cudaStream_t st) {
  if (a.g.nd < 2 || !a.alpha_lf) return cudaErrorInvalidValue;
  long long total = (long long)a.g.n[0] * a.g.n[1] * (a.g.nd == 3 ? a.g.n[2] : 1);
  const unsigned grid = (unsigned)((total + 127) / 128);
  if (a.np.rs == RS_HYBRID_HLLC || a.np.rs == RS_AM_HLLC) {
    if (a.g.nd == 2) {
      if (a.np.rs == RS_HYBRID_HLLC) hydro_stage_pp_kernel<T, 2, RS_HYBRID_HLLC><<<grid, 128, 0, st>>>(a);
      else hydro_stage_pp_kernel<T, 2, RS_AM_HLLC><<<grid, 128, 0, st>>>(a);
    } else {
      if (a.np.rs == RS_HYBRID_HLLC) hydro_stage_pp_kernel<T, 3, RS_HYBRID_HLLC><<<grid, 128, 0, st>>>(a);
      else hydro_stage_pp_kernel<T, 3, RS_AM_HLLC><<<grid, 128, 0, st>>>(a);
    }
    return cudaGetLastError();
  }
  AMX_DISPATCH_RS(a.np.rs, {
    if (a.g.nd == 2) hydro_stage_pp_kernel<T, 2, RS><<<grid, 128, 0, st>>>(a);
    else hydro_stage_pp_kernel<T, 3, RS><<<grid, 128, 0, st>>>(a);
  })
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_hydro_stage_t(const StageArgs& a, cudaStream_t st) {
  if (a.np.rs == RS_HYBRID_HLLC || a.np.rs == RS_AM_HLLC) {  // generic limiter switch only
    AMX_DISPATCH_ND(a.g.nd, {
      const unsigned grid = grid_for<ND>(a.g);
      if (a.np.rs == RS_HYBRID_HLLC) hydro_stage_kernel<T, ND, RS_HYBRID_HLLC, -1><<<grid, 256, 0, st>>>(a);
      else hydro_stage_kernel<T, ND, RS_AM_HLLC, -1><<<grid, 256, 0, st>>>(a);
    })
    return cudaGetLastError();
  }
  AMX_DISPATCH_ND(a.g.nd, AMX_DISPATCH_RS(a.np.rs, {
    const unsigned grid = grid_for<ND>(a.g);
    if (a.np.lim == LIM_MINMOD) hydro_stage_kernel<T, ND, RS, LIM_MINMOD><<<grid, 256, 0, st>>>(a);
    else hydro_stage_kernel<T, ND, RS, -1><<<grid, 256, 0, st>>>(a);
  }))
  return cudaGetLastError();
}
cudaError_t launch_hydro_stage(const StageArgs& a, cudaStream_t st) {
  if (a.np.lim == LIM_VAN_ALBADA_PP)
    return a.np.dtype == 0 ? launch_hydro_stage_pp_t<float>(a, st) : launch_hydro_stage_pp_t<double>(a, st);
  return a.np.dtype == 0 ? launch_hydro_stage_t<float>(a, st) : launch_hydro_stage_t<double>(a, st);
}

template <typename T>
static cudaError_t launch_split_sweep_t(const SweepArgs& a, cudaStream_t st) {
  if (a.axis < 1 || a.axis > a.g.nd) return cudaErrorInvalidValue;
#define AMX_LAUNCH_SWEEP(RS_)                                                                            \
  AMX_DISPATCH_ND(a.g.nd, {                                                                              \
    const unsigned grid = grid_for<ND>(a.g);                                                             \
    if (a.axis == 1) split_sweep_kernel<T, ND, RS_, 0><<<grid, 256, 0, st>>>(a);                         \
    if constexpr (ND >= 2) { if (a.axis == 2) split_sweep_kernel<T, ND, RS_, 1><<<grid, 256, 0, st>>>(a); } \
    if constexpr (ND >= 3) { if (a.axis == 3) split_sweep_kernel<T, ND, RS_, 2><<<grid, 256, 0, st>>>(a); } \
  })
  if (a.np.rs == RS_HYBRID_HLLC) { AMX_LAUNCH_SWEEP(RS_HYBRID_HLLC) return cudaGetLastError(); }
  if (a.np.rs == RS_AM_HLLC) { AMX_LAUNCH_SWEEP(RS_AM_HLLC) return cudaGetLastError(); }
#undef AMX_LAUNCH_SWEEP
  AMX_DISPATCH_ND(a.g.nd, AMX_DISPATCH_RS(a.np.rs, {
    const unsigned grid = grid_for<ND>(a.g);
    if (a.axis == 1) split_sweep_kernel<T, ND, RS, 0><<<grid, 256, 0, st>>>(a);
    if constexpr (ND >= 2) { if (a.axis == 2) split_sweep_kernel<T, ND, RS, 1><<<grid, 256, 0, st>>>(a); }
    if constexpr (ND >= 3) { if (a.axis == 3) split_sweep_kernel<T, ND, RS, 2><<<grid, 256, 0, st>>>(a); }
  }))
  return cudaGetLastError();
}
cudaError_t launch_split_sweep(const SweepArgs& a, cudaStream_t st) {
  return a.np.dtype == 0 ? launch_split_sweep_t<float>(a, st) : launch_split_sweep_t<double>(a, st);
}

template <typename T>
static cudaError_t launch_mag_iter_t(const MagIterArgs& a, cudaStream_t st) {
  if (a.g.nd == 2) mag_iter_kernel<T, 2><<<grid_for<2>(a.g), 256, 0, st>>>(a);
  else if (a.g.nd == 3) mag_iter_kernel<T, 3><<<grid_for<3>(a.g), 256, 0, st>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
cudaError_t launch_mag_iter(const MagIterArgs& a, cudaStream_t st) {
  return a.np.dtype == 0 ? launch_mag_iter_t<float>(a, st) : launch_mag_iter_t<double>(a, st);
}

template <typename T>
static cudaError_t launch_mag_final_t(const MagFinalArgs& a, cudaStream_t st) {
  // at most 148 SMs x 8 resident blocks; the kernel strides over the cells
  const unsigned cap = 148u * 8u;
  if (a.g.nd == 2) mag_final_kernel<T, 2><<<std::min(grid_for<2>(a.g), cap), 256, 0, st>>>(a);
  else if (a.g.nd == 3) mag_final_kernel<T, 3><<<std::min(grid_for<3>(a.g), cap), 256, 0, st>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
cudaError_t launch_mag_final(const MagFinalArgs& a, cudaStream_t st) {
  return a.np.dtype == 0 ? launch_mag_final_t<float>(a, st) : launch_mag_final_t<double>(a, st);
}

template <typename T>
static cudaError_t launch_alpha_t(const AlphaArgs& a, cudaStream_t st) {
  AMX_DISPATCH_ND(a.g.nd, { alpha_kernel<T, ND><<<grid_for<ND>(a.g), 256, 0, st>>>(a); })
  return cudaGetLastError();
}
cudaError_t launch_alpha(const AlphaArgs& a, cudaStream_t st) {
  return a.np.dtype == 0 ? launch_alpha_t<float>(a, st) : launch_alpha_t<double>(a, st);
}

template <typename T>
static cudaError_t launch_shell_t(const ShellArgs& a, cudaStream_t st) {
  long long tot = 1;
  for (int k = 0; k < a.g.nd; ++k) tot *= a.g.n[k] + 2 * a.ng;
  const unsigned grid = (unsigned)((tot + 255) / 256);
  AMX_DISPATCH_ND(a.g.nd, { shell_kernel<T, ND><<<grid, 256, 0, st>>>(a); })
  return cudaGetLastError();
}
cudaError_t launch_shell(const ShellArgs& a, cudaStream_t st) {
  return a.np.dtype == 0 ? launch_shell_t<float>(a, st) : launch_shell_t<double>(a, st);
}

}  // namespace AMX_NS
