// amx_math.cuh -- cell-level arithmetic of the finite-volume (M)HD step (device functions).
//
// Everything here is elementwise algebra on a handful of registers: equation of state,
// Euler flux, slope limiters, HLL / HLLC / HLLC-LM / Lax-Friedrichs fluxes, the
// MUSCL-Hancock predictor.  The file is compiled twice (see amx_kernels.cu):
//   AMX_STRICT=1  -fmad=false, IEEE div/sqrt: the reference's operation order, so results
//                 agree with the NumPy oracle to the last bit wherever libm agrees;
//   AMX_STRICT=0  contracted FMAs, reciprocal multiplies, approximate fp32 sqrt/div.
// Reference lines (relative to /root/reference/astronomix/) are cited per function.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef AMX_STRICT
#define AMX_STRICT 0
#endif
#ifndef AMX_MINMOD_ALU
#define AMX_MINMOD_ALU 0
#endif

namespace amx {

enum : int { LIM_MINMOD = 0, LIM_OSHER = 1, LIM_DOUBLE_MINMOD = 2, LIM_SUPERBEE = 3, LIM_VAN_ALBADA = 4,
             LIM_VAN_ALBADA_PP = 5 };
enum : int { RS_HLL = 0, RS_HLLC = 1, RS_HLLC_LM = 2, RS_LF = 3, RS_HYBRID_HLLC = 4, RS_AM_HLLC = 5 };

#define AMX_DI __device__ __forceinline__

// ---------------------------------------------------------------- scalar helpers
AMX_DI float rdiv(float a, float b) {
#if AMX_STRICT
  return a / b;
#else
  // one MUFU.RCP + one FMUL (__fdividef wraps the same MUFU in a denormal-range fix-up of
  // 2 FSETP/FSEL + 2 FMUL per call: 55 instructions per cell of the hydro stage)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return a * r;
#endif
}
AMX_DI double rdiv(double a, double b) { return a / b; }

AMX_DI float rsqrt_(float a) {
#if AMX_STRICT
  return sqrtf(a);
#else
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#endif
}
AMX_DI double rsqrt_(double a) { return sqrt(a); }

AMX_DI float rsin(float a) {
#if AMX_STRICT
  return sinf(a);
#else
  return __sinf(a);
#endif
}
AMX_DI double rsin(double a) { return sin(a); }

AMX_DI float rabs(float a) { return fabsf(a); }
AMX_DI double rabs(double a) { return fabs(a); }
AMX_DI float rmin(float a, float b) { return fminf(a, b); }
AMX_DI double rmin(double a, double b) { return fmin(a, b); }
AMX_DI float rmax(float a, float b) { return fmaxf(a, b); }
AMX_DI double rmax(double a, double b) { return fmax(a, b); }
template <typename T> AMX_DI T sq(T a) { return a * a; }
template <typename T> AMX_DI T sgn(T a) { return a > T(0) ? T(1) : (a < T(0) ? T(-1) : T(0)); }

// ---------------------------------------------------------------- packed pair of floats
// Two cells adjacent along the contiguous axis handled by one thread (amx_vec.cu).  In the FAST
// flavour +, -, * are the packed sm_100a instructions FADD2 / FMUL2 (ptxas contracts them into
// FFMA2 and folds negations / scalar broadcasts into operand modifiers).  ptxas contracts
// mul.rn.f32x2 + add.rn.f32x2 even under -fmad=false, so the STRICT flavour uses the never-fused
// scalar intrinsics (same IEEE results, reference operation order).
struct F2 {
  float x, y;
  F2() = default;
  AMX_DI explicit F2(float a) : x(a), y(a) {}
  AMX_DI explicit F2(double a) : x((float)a), y((float)a) {}
  AMX_DI explicit F2(int a) : x((float)a), y((float)a) {}
  AMX_DI F2(float a, float b) : x(a), y(b) {}
};
AMX_DI F2 mk2(float a, float b) { return F2(a, b); }
AMX_DI F2 bc2(float a) { return F2(a, a); }
#if AMX_STRICT
AMX_DI F2 operator+(F2 a, F2 b) { return mk2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
AMX_DI F2 operator-(F2 a, F2 b) { return mk2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
AMX_DI F2 operator*(F2 a, F2 b) { return mk2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
#else
#define AMX_F2_BINOP(NAME, PTX)                                                                      \
  AMX_DI F2 NAME(F2 a, F2 b) {                                                                       \
    F2 d;                                                                                            \
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; " PTX                   \
        " rd, ra, rb; mov.b64 {%0, %1}, rd; }"                                                       \
        : "=f"(d.x), "=f"(d.y)                                                                       \
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                                                   \
    return d;                                                                                        \
  }
AMX_F2_BINOP(operator+, "add.rn.f32x2")
AMX_F2_BINOP(operator-, "sub.rn.f32x2")
AMX_F2_BINOP(operator*, "mul.rn.f32x2")
#undef AMX_F2_BINOP
// a * b + c as one packed FFMA2
AMX_DI F2 fma2(F2 a, F2 b, F2 c) {
  F2 d;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
#endif
AMX_DI F2 operator-(F2 a) { return mk2(-a.x, -a.y); }
AMX_DI F2 operator/(F2 a, F2 b) { return mk2(a.x / b.x, a.y / b.y); }  // IEEE division per half
AMX_DI F2 rdiv(F2 a, F2 b) { return mk2(rdiv(a.x, b.x), rdiv(a.y, b.y)); }
AMX_DI F2 rsqrt_(F2 a) { return mk2(rsqrt_(a.x), rsqrt_(a.y)); }
AMX_DI F2 rsin(F2 a) { return mk2(rsin(a.x), rsin(a.y)); }
AMX_DI F2 rabs(F2 a) { return mk2(fabsf(a.x), fabsf(a.y)); }
AMX_DI F2 rmin(F2 a, F2 b) { return mk2(fminf(a.x, b.x), fminf(a.y, b.y)); }
AMX_DI F2 rmax(F2 a, F2 b) { return mk2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }

// s >= 0 ? a : b and s <= 0 ? a : b (upwind selection of the HLLC family), per half for pairs
template <typename T> AMX_DI T sel_ge0(T s, T a, T b) { return (s >= T(0)) ? a : b; }
template <typename T> AMX_DI T sel_le0(T s, T a, T b) { return (s <= T(0)) ? a : b; }
AMX_DI F2 sel_ge0(F2 s, F2 a, F2 b) { return mk2(s.x >= 0.f ? a.x : b.x, s.y >= 0.f ? a.y : b.y); }
AMX_DI F2 sel_le0(F2 s, F2 a, F2 b) { return mk2(s.x <= 0.f ? a.x : b.x, s.y <= 0.f ? a.y : b.y); }

// Per-launch constants, rounded to the working type the way the reference's scalars are
// under jit: gamma is a traced scalar of the working dtype, so gamma-1 is
// evaluated in T; grid_spacing is a Python double folded at trace time and rounded on
// first contact with an array.
template <typename T>
struct Consts {
  T gamma, gm1;     // gamma, gamma - 1 (in T)
  T inv_gm1;        // 1/(gamma-1), used by the FAST flavour only
  T dx, inv_dx;     // T(dx), T(1/dx): "1 / grid_spacing * stencil" multiplies by T(1/dx)
  T half_dx;        // T(dx/2): split reconstruction distances
  T va_eps;         // T(3*dx): van Albada epsilon (Q10)
};

// Gas state of one cell: rho, v[0..ND), p  (conserved form: rho, m[0..ND), E)
template <typename T, int ND>
struct Gas {
  T q[ND + 2];
  static constexpr int P = ND + 1;
};

// ---------------------------------------------------------------- equation of state
// _fluid_equations/_equations.py:165-199 get_absolute_velocity, squared.  STRICT takes the
// square root and squares it again, as the reference does (1e-20 only in 3-D, Q2).
template <typename T, int ND>
AMX_DI T absvel_sq(const Gas<T, ND>& s) {
  if constexpr (ND == 1) {
#if AMX_STRICT
    return sq(rabs(s.q[1]));
#else
    return s.q[1] * s.q[1];
#endif
  } else {
    T a = sq(s.q[1]) + sq(s.q[2]);
    if constexpr (ND == 3) a = a + sq(s.q[3]) + T(1e-20);
#if AMX_STRICT
    return sq(rsqrt_(a));
#else
    return a;
#endif
  }
}

// get_absolute_velocity itself (|u| in 1-D, sqrt(sum v^2 [+ 1e-20 in 3-D]) otherwise)
template <typename T, int ND>
AMX_DI T absvel(const Gas<T, ND>& s) {
  if constexpr (ND == 1) return rabs(s.q[1]);
  T a = sq(s.q[1]) + sq(s.q[2]);
  if constexpr (ND == 3) a = a + sq(s.q[3]) + T(1e-20);
  return rsqrt_(a);
}

// _equations.py:253-266: p/(gamma-1) + 0.5*rho*u^2
template <typename T, int ND>
AMX_DI T total_energy(const Gas<T, ND>& s, const Consts<T>& c) {
#if AMX_STRICT
  return s.q[Gas<T, ND>::P] / c.gm1 + (T(0.5) * s.q[0]) * absvel_sq(s);
#else
  return s.q[Gas<T, ND>::P] * c.inv_gm1 + (T(0.5) * s.q[0]) * absvel_sq(s);
#endif
}

// _equations.py:104-156
template <typename T, int ND>
AMX_DI Gas<T, ND> cons_from_prim(const Gas<T, ND>& s, const Consts<T>& c) {
  Gas<T, ND> U;
  U.q[0] = s.q[0];
#pragma unroll
  for (int k = 1; k <= ND; ++k) U.q[k] = s.q[0] * s.q[k];
  U.q[ND + 1] = total_energy(s, c);
  return U;
}

// _equations.py:24-94: v = m/rho; u = sqrt(sum v^2 + 1e-20) (2-D, 3-D); p = (gamma-1)*rho*(E/rho - 0.5u^2)
template <typename T, int ND>
AMX_DI Gas<T, ND> prim_from_cons(const Gas<T, ND>& U, const Consts<T>& c) {
  Gas<T, ND> s;
  const T rho = U.q[0];
  s.q[0] = rho;
  T u2;
#if AMX_STRICT
#pragma unroll
  for (int k = 1; k <= ND; ++k) s.q[k] = U.q[k] / rho;
  const T Eor = U.q[ND + 1] / rho;
#else
  const T ir = rdiv(T(1), rho);
#pragma unroll
  for (int k = 1; k <= ND; ++k) s.q[k] = U.q[k] * ir;
  const T Eor = U.q[ND + 1] * ir;
#endif
  if constexpr (ND == 1) {
    u2 = sq(s.q[1]);
  } else {
    T a = sq(s.q[1]) + sq(s.q[2]);
    if constexpr (ND == 3) a = a + sq(s.q[3]);
    a = a + T(1e-20);
#if AMX_STRICT
    u2 = sq(rsqrt_(a));
#else
    u2 = a;
#endif
  }
  s.q[ND + 1] = (c.gm1 * rho) * (Eor - T(0.5) * u2);
  return s;
}

// _equations.py:269-280
template <typename T>
AMX_DI T sound_speed(T rho, T p, const Consts<T>& c) { return rsqrt_(rdiv(c.gamma * p, rho)); }

// _fluid_equations/_fluxes.py:25-93; ax = 1..ND
template <typename T, int ND>
AMX_DI Gas<T, ND> euler_flux(const Gas<T, ND>& s, int ax, const Consts<T>& c) {
  Gas<T, ND> F;
  const T rho = s.q[0], p = s.q[ND + 1], un = s.q[ax];
  const T E = total_energy(s, c);
  F.q[0] = un * rho;
#pragma unroll
  for (int k = 1; k <= ND; ++k) F.q[k] = un * (s.q[k] * rho);
  F.q[ND + 1] = un * (p + E);
  F.q[ax] = F.q[ax] + p;
  return F;
}

// _magnetic_field_update.py:208-240: pressure after the magnetic step from the energy bookkeeping
// E2 = E - rho |v_0|^2 / 2 + rho |v_n|^2 / 2 with gamma hard-coded to 5/3 (Q4).  n0 / nn are the
// sums of squares of the old / new velocity components (norm-then-square in the reference: the
// STRICT flavour takes the square root and squares again).
template <typename T>
AMX_DI T magnetic_writeback_pressure(T rho, T p, T n0, T nn) {
  const T g53m1 = (T)(5.0 / 3.0 - 1.0);
#if AMX_STRICT
  n0 = sq(rsqrt_(n0));
  nn = sq(rsqrt_(nn));
#endif
  const T E = rdiv(p, g53m1) + (T(0.5) * rho) * n0;
  const T E2 = E - (T(0.5) * rho) * n0 + (T(0.5) * rho) * nn;
  return (g53m1 * rho) * (rdiv(E2, rho) - T(0.5) * nn);
}

// ---------------------------------------------------------------- limiters
// _finite_volume/_state_evolution/limiters.py:14-15.  0.5*(sgn a + sgn b)*min(|a|,|b|) is
// min(a,b) for two positives, max(a,b) for two negatives and (signed) zero otherwise.
template <typename T>
AMX_DI T minmod(T a, T b) {
  // max(0, min(a,b)) + min(0, max(a,b)): exact, branch-free, 4 FMNMX + 1 FADD
  return rmax(T(0), rmin(a, b)) + rmin(T(0), rmax(a, b));
}
template <typename T>
AMX_DI T maxmod(T a, T b) {  // limiters.py:35-37
  return T(0.5) * (sgn(a) + sgn(b)) * rmax(rabs(a), rabs(b));
}
template <typename T>
AMX_DI T minmod3(T a, T b, T c) {  // limiters.py:18-25
  const bool same = (sgn(a) == sgn(b)) && (sgn(b) == sgn(c));
  return same ? sgn(a) * rmin(rabs(a), rmin(rabs(b), rabs(c))) : T(0);
}

// limited_gradients.py:50-87: a = backward difference / dx, b = forward difference / dx
template <typename T>
AMX_DI T limited_slope(T a, T b, int lim, const Consts<T>& c) {
  switch (lim) {
    case LIM_MINMOD:
      return minmod(a, b);
    case LIM_SUPERBEE:  // limiters.py:40-44
      return (a * b > T(0)) ? minmod(maxmod(a, b), minmod(T(2) * a, T(2) * b)) : T(0);
    case LIM_DOUBLE_MINMOD:  // limiters.py:28-32
      return (a * b > T(0)) ? minmod3((a + b) / T(2), T(2) * a, T(2) * b) : T(0);
    case LIM_VAN_ALBADA: {  // limiters.py:48-64
      const T e = c.va_eps;
      return rdiv((sq(b) + e) * a + (sq(a) + e) * b, sq(b) + sq(a) + T(2) * e);
    }
    default: {  // LIM_OSHER, limited_gradients.py:72-84
      const T eps = T(1e-11);
      const T g = (rabs(a) > eps) ? rdiv(b, a + eps) : T(0);
      return rmax(T(0), rmin(T(1.3), g)) * a;
    }
  }
}

#if !AMX_STRICT
// FAST pairs: minmod(a, b) = median(a, b, 0) = a + b - max3(a, b, 0) - min3(a, b, 0): two 3-input
// FMNMX3 per cell on the half-rate ALU pipe instead of four FMNMX, the sums as packed adds.  The sum
// rounds (error <= ulp(a + b)); the exact min / max form stays in the STRICT flavour.
AMX_DI F2 minmod(F2 a, F2 b) {
#if AMX_MINMOD_ALU
  // median(a, b, 0) = max(min(a, b), min(max(a, b), 0)): exact, four FMNMX per cell, nothing on the FMA pipe
  return mk2(fmaxf(fminf(a.x, b.x), fminf(fmaxf(a.x, b.x), 0.f)), fmaxf(fminf(a.y, b.y), fminf(fmaxf(a.y, b.y), 0.f)));
#else
  const F2 hi = mk2(fmaxf(fmaxf(a.x, b.x), 0.f), fmaxf(fmaxf(a.y, b.y), 0.f));
  const F2 lo = mk2(fminf(fminf(a.x, b.x), 0.f), fminf(fminf(a.y, b.y), 0.f));
  return ((a + b) - hi) - lo;
#endif
}
#endif

// pairs: the non-minmod limiters branch per cell, so they run per half on the scalar code
AMX_DI F2 limited_slope(F2 a, F2 b, int lim, const Consts<F2>& c) {
  if (lim == LIM_MINMOD) return minmod(a, b);
  Consts<float> cf;
  cf.gamma = c.gamma.x; cf.gm1 = c.gm1.x; cf.inv_gm1 = c.inv_gm1.x; cf.dx = c.dx.x; cf.inv_dx = c.inv_dx.x;
  cf.half_dx = c.half_dx.x; cf.va_eps = c.va_eps.x;
  return mk2(limited_slope(a.x, b.x, lim, cf), limited_slope(a.y, b.y, lim, cf));
}

// limited slope (per unit length) of every gas variable from the three cells m, c, p
template <typename T, int ND, int LIM>
AMX_DI Gas<T, ND> slopes(const Gas<T, ND>& m, const Gas<T, ND>& s, const Gas<T, ND>& p, int lim,
                         const Consts<T>& c) {
  Gas<T, ND> g;
#pragma unroll
  for (int k = 0; k < ND + 2; ++k) {
#if AMX_STRICT
    const T bd = (s.q[k] - m.q[k]) / c.dx;
    const T fd = (p.q[k] - s.q[k]) / c.dx;
#else
    const T bd = (s.q[k] - m.q[k]) * c.inv_dx;
    const T fd = (p.q[k] - s.q[k]) * c.inv_dx;
#endif
    g.q[k] = (LIM == LIM_MINMOD) ? minmod(bd, fd) : limited_slope(bd, fd, lim, c);
  }
  return g;
}

// ---------------------------------------------------------------- Riemann solvers
// _riemann_solver/hll.py:36-109
template <typename T, int ND>
AMX_DI Gas<T, ND> hll_flux(const Gas<T, ND>& L, const Gas<T, ND>& R, int ax, const Consts<T>& c) {
  constexpr int P = ND + 1;
#if AMX_STRICT
  const T uL = L.q[ax], uR = R.q[ax];
  const T cL = sound_speed(L.q[0], L.q[P], c), cR = sound_speed(R.q[0], R.q[P], c);
  const Gas<T, ND> FL = euler_flux(L, ax, c), FR = euler_flux(R, ax, c);
  const T SR = rmax(rmax(uL + cL, uR + cR), T(0));
  const T SL = rmin(rmin(uL - cL, uR - cR), T(0));
  const Gas<T, ND> UL = cons_from_prim(L, c), UR = cons_from_prim(R, c);
  Gas<T, ND> F;
  const T den = SR - SL;
  const T SLSR = SL * SR;
#pragma unroll
  for (int k = 0; k < ND + 2; ++k)
    F.q[k] = (SR * FL.q[k] - SL * FR.q[k] + SLSR * (UR.q[k] - UL.q[k])) / den;
  return F;
#else
  // Same flux, arranged for the FMA pipe.  With a = SR/(SR-SL), b = SL/(SR-SL) (so a*SL = b*SR) and
  // F_side = un*U + P (P = p in the normal-momentum row, p*un in the energy row, 0 elsewhere):
  //   F = a*FL - b*FR + a*SL*(UR - UL) = [a(unL - SL)]*UL - [b(unR - SR)]*UR + (a*PL - b*PR),
  // i.e. two multiply-adds per row on the conserved states instead of forming both fluxes and the
  // difference of the states; the momenta are never formed (AL*rho and BR*rho multiply the velocities).
  // Supersonic sides stay exact: SL = 0 gives b = 0, a = 1, F = FL.
  T un[2], cs[2], pr[2], rho[2], E[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const Gas<T, ND>& W = s == 0 ? L : R;
    rho[s] = W.q[0];
    pr[s] = W.q[P];
    un[s] = W.q[ax];
    cs[s] = rsqrt_((c.gamma * pr[s]) * rdiv(T(1), rho[s]));
    T ke = ND == 3 ? T(1e-20) : T(0);
#pragma unroll
    for (int k = 0; k < ND; ++k) ke = ke + W.q[k + 1] * W.q[k + 1];
    E[s] = pr[s] * c.inv_gm1 + (T(0.5) * rho[s]) * ke;
  }
  const T SR = rmax(rmax(un[0] + cs[0], un[1] + cs[1]), T(0));
  const T SL = rmin(rmin(un[0] - cs[0], un[1] - cs[1]), T(0));
  const T iden = rdiv(T(1), SR - SL);
  const T a = SR * iden, b = SL * iden;
  const T AL = a * (un[0] - SL), BR = b * (un[1] - SR);
  const T ALr = AL * rho[0], BRr = BR * rho[1];
  const T apL = a * pr[0], bpR = b * pr[1];
  const Gas<T, ND>& WL = L;
  const Gas<T, ND>& WR = R;
  Gas<T, ND> F;
  F.q[0] = ALr - BRr;
#pragma unroll
  for (int k = 0; k < ND; ++k) F.q[k + 1] = ALr * WL.q[k + 1] - BRr * WR.q[k + 1];
  F.q[ax] = F.q[ax] + (apL - bpR);
  F.q[P] = (AL * E[0] - BR * E[1]) + (apL * un[0] - bpR * un[1]);
  return F;
#endif
}

// _riemann_solver/hll.py:123-261 (HLLC; LM = HLLC-LM low-Mach rescaling of SL, SR in the
// dissipation term only; the upwind selection tests use the unscaled speeds)
// LMDC: low-Mach dissipation control of AM-HLLC (:230-245): the momentum rows of the dissipation
// term are scaled by min(1, max(|v_L|/c_L, |v_R|/c_R)).
template <typename T, int ND, bool LM, bool LMDC = false>
AMX_DI Gas<T, ND> hllc_flux(const Gas<T, ND>& L, const Gas<T, ND>& R, int ax, const Consts<T>& c) {
  constexpr int P = ND + 1;
  const T rhoL = L.q[0], rhoR = R.q[0], uL = L.q[ax], uR = R.q[ax], pL = L.q[P], pR = R.q[P];
  const T cL = sound_speed(rhoL, pL, c), cR = sound_speed(rhoR, pR, c);
  const Gas<T, ND> FL = euler_flux(L, ax, c), FR = euler_flux(R, ax, c);
  const T sL = rsqrt_(rhoL), sR = rsqrt_(rhoR);
  const T ssum = sL + sR;
  const T u_hat = rdiv(sL * uL + sR * uR, ssum);
  const T c_hat_sq = rdiv(sq(cL) * sL + sq(cR) * sR, ssum) + (T(0.5) * rdiv(sL * sR, sq(ssum))) * sq(uR - uL);
  const T c_hat = rsqrt_(c_hat_sq);
  const T SL = rmin(uL - cL, u_hat - c_hat);
  const T SR = rmax(uR + cR, u_hat + c_hat);
  const T dL = SL - uL, dR = SR - uR;
  const T Sstar = rdiv(pR - pL + rhoL * uL * dL - rhoR * uR * dR, rhoL * dL - rhoR * dR);
  const Gas<T, ND> UL = cons_from_prim(L, c), UR = cons_from_prim(R, c);
  Gas<T, ND> UsL = UL, UsR = UR;
  UsL.q[ax] = rhoL * Sstar;
  UsL.q[P] = UsL.q[P] + (Sstar - uL) * (rhoL * Sstar + rdiv(pL, dL));
  UsR.q[ax] = rhoR * Sstar;
  UsR.q[P] = UsR.q[P] + (Sstar - uR) * (rhoR * Sstar + rdiv(pR, dR));
  const T denL = SL - Sstar, denR = SR - Sstar;
#pragma unroll
  for (int k = 0; k < ND + 2; ++k) {
    UsL.q[k] = rdiv(UsL.q[k] * dL, denL);
    UsR.q[k] = rdiv(UsR.q[k] * dR, denR);
  }
  T SLs = SL, SRs = SR;
  if constexpr (LM) {
    const T Ma = rmax(rabs(rdiv(uL, cL)), rabs(rdiv(uR, cR)));
    const T phi = rsin(rmin(T(1), rdiv(Ma, T(0.1))) * T(3.141592653589793) / T(2));
    SLs = SL * phi;
    SRs = SR * phi;
  }
  const T aS = rabs(Sstar);
  T fm = T(1);
  if constexpr (LMDC) {
    const T aL = absvel(L), aR = absvel(R);
    fm = rmin(T(1), rmax(rabs(rdiv(aL, cL)), rabs(rdiv(aR, cR))));
  }
  Gas<T, ND> F;
#pragma unroll
  for (int k = 0; k < ND + 2; ++k) {
    const T bulk = T(0.5) * (FL.q[k] + FR.q[k]);
    T diss = T(0.5) * (SLs * (UsL.q[k] - UL.q[k]) + aS * (UsL.q[k] - UsR.q[k]) + SRs * (UsR.q[k] - UR.q[k]));
    if (LMDC && k >= 1 && k <= ND) diss = fm * diss;
    T f = bulk + diss;
    f = sel_ge0(SL, FL.q[k], f);
    f = sel_le0(SR, FR.q[k], f);
    F.q[k] = f;
  }
  return F;
}

// _riemann_solver/_lax_friedrichs.py:21-78 with the global alpha of the axis passed in
template <typename T, int ND>
AMX_DI Gas<T, ND> lf_flux(const Gas<T, ND>& L, const Gas<T, ND>& R, int ax, T alpha, const Consts<T>& c) {
  const Gas<T, ND> UL = cons_from_prim(L, c), UR = cons_from_prim(R, c);
  const Gas<T, ND> FL = euler_flux(L, ax, c), FR = euler_flux(R, ax, c);
  Gas<T, ND> F;
#pragma unroll
  for (int k = 0; k < ND + 2; ++k)
    F.q[k] = T(0.5) * (FL.q[k] + FR.q[k]) - (T(0.5) * alpha) * (UR.q[k] - UL.q[k]);
  return F;
}

// _riemann_solver/_riemann_solver.py:37-87
template <typename T, int ND, int RS>
AMX_DI Gas<T, ND> riemann(const Gas<T, ND>& L, const Gas<T, ND>& R, int ax, T alpha, const Consts<T>& c) {
  if constexpr (RS == RS_HLL) return hll_flux(L, R, ax, c);
  else if constexpr (RS == RS_HLLC) return hllc_flux<T, ND, false>(L, R, ax, c);
  else if constexpr (RS == RS_HLLC_LM) return hllc_flux<T, ND, true>(L, R, ax, c);
  else if constexpr (RS == RS_LF) return lf_flux(L, R, ax, alpha, c);
  else {
    // hll.py:268-326 HYBRID-HLLC / AM-HLLC: HLLC-LM where the cell on the right of the interface is
    // being compressed (g = 1, passed in `alpha`), plain HLLC elsewhere
    constexpr bool AM = RS == RS_AM_HLLC;
    const Gas<T, ND> Fc = hllc_flux<T, ND, false, AM>(L, R, ax, c);
    const Gas<T, ND> Fl = hllc_flux<T, ND, true, AM>(L, R, ax, c);
    Gas<T, ND> F;
#pragma unroll
    for (int k = 0; k < ND + 2; ++k) F.q[k] = alpha * Fl.q[k] + (T(1) - alpha) * Fc.q[k];
    return F;
  }
}

// ---------------------------------------------------------------- reconstruction
// reconstruction.py:399-431: face values of one cell from its limited slope g (per unit
// length): lo = s - (g*dx)/2 (right state of the interface below), hi = s + (g*dx)/2.
template <typename T, int ND>
AMX_DI void faces_unsplit(const Gas<T, ND>& s, const Gas<T, ND>& g, const Consts<T>& c, Gas<T, ND>& lo,
                          Gas<T, ND>& hi) {
#pragma unroll
  for (int k = 0; k < ND + 2; ++k) {
    const T d = (g.q[k] * c.dx) / T(2);
    lo.q[k] = s.q[k] - d;
    hi.q[k] = s.q[k] + d;
  }
}

// reconstruction.py:40-129: MUSCL-Hancock predictor + face values for a sweep along ax.
// proj = A_W . g with A_W = u*I, A_W[rho,ax] = rho, A_W[p,1] = rho c^2 (literal 1: Q1),
// A_W[ax,p] = 1/rho; pred = s - dt/2 * proj; faces = pred -/+ (dx/2) g.
template <typename T, int ND>
AMX_DI void faces_muscl(const Gas<T, ND>& s, const Gas<T, ND>& g, int ax, T half_dt, const Consts<T>& c,
                        Gas<T, ND>& lo, Gas<T, ND>& hi) {
  constexpr int P = ND + 1;
  const T rho = s.q[0], u = s.q[ax];
  const T cs = sound_speed(rho, s.q[P], c);
  Gas<T, ND> proj;
#pragma unroll
  for (int k = 0; k < ND + 2; ++k) proj.q[k] = u * g.q[k];
  proj.q[0] = proj.q[0] + rho * g.q[ax];
  proj.q[ax] = proj.q[ax] + rdiv(T(1), rho) * g.q[P];
  proj.q[P] = (rho * sq(cs)) * g.q[1] + proj.q[P];
#pragma unroll
  for (int k = 0; k < ND + 2; ++k) {
    const T pred = s.q[k] - half_dt * proj.q[k];
    const T d = c.half_dx * g.q[k];
    lo.q[k] = pred - d;
    hi.q[k] = pred + d;
  }
}

// |u_ax| + c of a cell: the CFL / Lax-Friedrichs wave-speed bound
// (_timestep_estimator.py:118-128, _lax_friedrichs.py:61-66)
template <typename T, int ND>
AMX_DI T wave_bound(const Gas<T, ND>& s, int ax, const Consts<T>& c) {
  return rabs(s.q[ax]) + sound_speed(s.q[0], s.q[ND + 1], c);
}

}  // namespace amx
