// amx_api.cu -- the C ABI of libamx (include/amx.h): descriptor validation, workspace
// layout, the control kernels that keep dt / t / the fixed-point flag on the device, and
// the step / integrate drivers.  Compiled with -fmad=false: the scalar dt arithmetic here
// follows the reference's operation order in both arithmetic flavours.
#include <cuda_runtime.h>

#include <algorithm>
#include <unordered_map>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cmath>
#include <vector>

#include "../../include/amx.h"
#include "amx_internal.h"

using namespace amx;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  return fail(AMX_ECUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CK(expr)                                        \
  do {                                                  \
    cudaError_t e__ = (expr);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #expr); \
  } while (0)

// ---- optional per-kernel-class timing (CUDA events on the launch stream; bench.py) -----------
enum { PC_HYDRO_STAGE = 0, PC_SPLIT_SWEEP, PC_MAG_ITER, PC_MAG_FINAL, PC_ALPHA, PC_CONTROL, PC_MAG_PAIR, PC_COUNT };
struct Profiler {
  bool on = false;
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[PC_COUNT] = {0};
  long long cnt[PC_COUNT] = {0};
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  void begin(int cls, cudaStream_t st) {
    if (!on) return;
    Rec r{cls, get(), get()};
    cudaEventRecord(r.a, st);
    recs.push_back(r);
  }
  void end(cudaStream_t st) {
    if (!on) return;
    cudaEventRecord(recs.back().b, st);
  }
  void resolve() {
    for (auto& r : recs) {
      float t = 0;
      if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
        ms[r.cls] += t;
        cnt[r.cls] += 1;
      }
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    recs.clear();
  }
};
Profiler g_prof;
struct ProfScope {
  cudaStream_t st;
  ProfScope(int cls, cudaStream_t s) : st(s) { g_prof.begin(cls, st); }
  ~ProfScope() { g_prof.end(st); }
};

struct Layout {
  int nd, ng, nv;          // dims, gas rows, total rows
  long long cells;         // active cells
  size_t esz;              // element size
  size_t row_bytes;        // cells * esz rounded up to 256
  // byte offsets into ws
  size_t off_ctl, off_S[2], off_A, off_G, off_Gf, off_X[2], total;
  size_t off_shell;        // ghost shell of the gas state (padded gas rows), shell mode only
};

// Written in front of the shell once it holds valid data: a resumed segment (t_start != 0) trusts the
// shell only if the stamp matches its own problem, otherwise it starts from images of the interior.
struct ShellStamp {
  unsigned long long magic;
  int nd, n[3], dtype, bc[3][2];
  bool operator==(const ShellStamp& o) const {
    bool same = magic == o.magic && nd == o.nd && dtype == o.dtype;
    for (int a = 0; a < 3; ++a) same = same && n[a] == o.n[a] && bc[a][0] == o.bc[a][0] && bc[a][1] == o.bc[a][1];
    return same;
  }
};
ShellStamp stamp_of(const amx_desc* d) {
  ShellStamp s{};
  s.magic = 0x414d585348454c4cull;  // "AMXSHELL"
  s.nd = d->ndim;
  s.dtype = d->dtype;
  for (int a = 0; a < 3; ++a) { s.n[a] = d->n[a]; s.bc[a][0] = d->bc[a][0]; s.bc[a][1] = d->bc[a][1]; }
  return s;
}

bool has_jet(const amx_desc* d) { return d->ndim == 2 && d->bc[1][0] == 3; }

// Unsplit MHD with ghost cells and a reflective side: the reference's ghost gas cells are not images
// of interior cells (Q3 + Q11 + Q12, see ShellArgs) and take part in the CFL reduction.
// Also with the MHD jet inflow (any scheme, hydro or MHD): its ghost cells hold constants that are not
// images of interior cells (|v_x| + c = 801 bounds dt from the first step on) and, in the unsplit
// scheme, drift from them by the rounding of prim(cons(.)).
bool shell_mode(const amx_desc* d) {
  if (d->boundary_handling != AMX_GHOST_CELLS) return false;
  if (d->bc[0][0] == AMX_BC_HALO || d->bc[0][1] == AMX_BC_HALO) return false;
  if (has_jet(d)) return true;
  if (!d->mhd || d->split != AMX_UNSPLIT) return false;
  for (int a = 0; a < d->ndim; ++a)
    if (d->bc[a][0] == AMX_BC_REFLECTIVE || d->bc[a][1] == AMX_BC_REFLECTIVE) return true;
  return false;
}

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

int validate(const amx_desc* d) {
  if (!d) return fail(AMX_EINVAL, "null descriptor");
  if (d->abi_version != AMX_ABI_VERSION) return fail(AMX_EINVAL, "amx_desc.abi_version mismatch");
  if (d->ndim < 1 || d->ndim > 3) return fail(AMX_EINVAL, "ndim must be 1, 2 or 3");
  for (int a = 0; a < 3; ++a) {
    if (a < d->ndim ? d->n[a] < 4 : d->n[a] != 1)
      return fail(AMX_EINVAL, "n[a] must be >= 4 on active axes and 1 on absent ones");
  }
  if (d->dtype != AMX_F32 && d->dtype != AMX_F64) return fail(AMX_EINVAL, "dtype must be AMX_F32 or AMX_F64");
  if (d->peer_flags & ~3) return fail(AMX_EINVAL, "peer_flags: only bits 0 (lower) and 1 (upper neighbour) exist");
  if (((d->peer_flags & 1) && d->bc[0][0] != AMX_BC_HALO) || ((d->peer_flags & 2) && d->bc[0][1] != AMX_BC_HALO))
    return fail(AMX_EINVAL, "peer_flags need AMX_BC_HALO on that side of axis 0");
  if (d->mhd && d->ndim == 1) return fail(AMX_EUNSUPPORTED, "MHD currently not supported in 1D.");
  if (d->riemann < AMX_HLL || d->riemann > 5) return fail(AMX_EINVAL, "Riemann solver not supported.");
  if (d->limiter < 0 || d->limiter > 5) return fail(AMX_EINVAL, "Unknown limiter.");
  if (d->limiter == 5 && d->split == AMX_UNSPLIT && d->ndim == 1)
    return fail(AMX_EUNSUPPORTED, "VAN_ALBADA_PP does not exist in 1-D (the reference raises there too: reconstruction.py reads velocity_index.x of an int)");
  if (d->split == AMX_UNSPLIT) {
    if (d->time_integrator != AMX_RK2_SSP)
      return fail(AMX_EINVAL, "Only the RK2 SSP time integrator is currently supported for the unsplit scheme.");
  } else if (d->split == AMX_SPLIT) {
    if (d->time_integrator != AMX_MUSCL && !d->first_order_fallback)
      return fail(AMX_EINVAL, "Only MUSCL is supported for split reconstruction.");
  } else {
    return fail(AMX_EINVAL, "split must be UNSPLIT or SPLIT");
  }
  if (d->boundary_handling != AMX_GHOST_CELLS && d->boundary_handling != AMX_PERIODIC_ROLL)
    return fail(AMX_EINVAL, "unknown boundary_handling");
  if (d->num_ghost != 2) return fail(AMX_EUNSUPPORTED, "num_ghost must be 2 (the reference pads by 2, Q8)");
  for (int a = 0; a < d->ndim; ++a) {
    for (int s = 0; s < 2; ++s) {
      const int b = d->bc[a][s];
      // _geometry/boundaries.py:180-188: the MHD jet inflow only exists as the LEFT y side of a 2-D run
      // (set anywhere else the reference's handler has no branch for it and the ghost cells go stale)
      const bool jet_ok = b == 3 && d->ndim == 2 && a == 1 && s == 0 && d->boundary_handling == AMX_GHOST_CELLS;
      if (b != AMX_BC_OPEN && b != AMX_BC_REFLECTIVE && b != AMX_BC_PERIODIC && b != AMX_BC_HALO && !jet_ok)
        return fail(b == 3 ? AMX_EUNSUPPORTED : AMX_EINVAL,
                    "boundary code not supported (the MHD jet boundary is the left y side of a 2-D ghost-cell run only)");
      if (b == AMX_BC_HALO && a != 0) return fail(AMX_EINVAL, "AMX_BC_HALO is only valid on axis 0");
    }
    if ((d->bc[a][0] == AMX_BC_PERIODIC) != (d->bc[a][1] == AMX_BC_PERIODIC))
      return fail(AMX_EUNSUPPORTED, "periodic boundaries must be set on both sides of an axis");
    if (d->boundary_handling == AMX_PERIODIC_ROLL && d->bc[a][0] != AMX_BC_PERIODIC && d->bc[a][0] != AMX_BC_HALO)
      return fail(AMX_EINVAL, "PERIODIC_ROLL implies periodic boundaries");
  }
  if (d->ndim == 2 && d->bc[1][0] == 3 && (d->riemann == AMX_LAX_FRIEDRICHS || (d->limiter == 5 && d->split == AMX_UNSPLIT)))
    return fail(AMX_EUNSUPPORTED, "MHD jet boundary with LAX_FRIEDRICHS / VAN_ALBADA_PP is not built (their global wave bounds "
                                  "would have to include the inflow cells)");
  if (!(d->dx > 0) || !(d->gamma > 1) || !(d->C_cfl > 0)) return fail(AMX_EINVAL, "dx, gamma-1, C_cfl must be positive");
  if (d->arithmetic != AMX_ARITH_FAST && d->arithmetic != AMX_ARITH_STRICT) return fail(AMX_EINVAL, "unknown arithmetic mode");
  return AMX_OK;
}

Layout make_layout(const amx_desc* d) {
  Layout L;
  L.nd = d->ndim;
  L.ng = d->ndim + 2;
  L.nv = L.ng + (d->mhd ? 3 : 0);
  L.cells = 1;
  // slab mode keeps the x ghost planes resident in every internal buffer
  const int gx = (d->bc[0][0] == AMX_BC_HALO || d->bc[0][1] == AMX_BC_HALO) ? 2 * d->num_ghost : 0;
  L.cells = (long long)(d->n[0] + gx) * d->n[1] * d->n[2];
  L.esz = d->dtype == AMX_F32 ? 4 : 8;
  L.row_bytes = align256((size_t)L.cells * L.esz);
  size_t off = 0;
  L.off_ctl = off; off += align256(sizeof(Ctl));
  for (int k = 0; k < 2; ++k) { L.off_S[k] = off; off += L.row_bytes * L.nv; }
  L.off_A = off; off += L.row_bytes * L.ng;
  L.off_G = off; off += L.row_bytes * L.ng;
  L.off_Gf = off; off += d->mhd ? L.row_bytes * L.ng : 0;  // gas after the magnetic write-back
  for (int k = 0; k < 2; ++k) { L.off_X[k] = off; off += d->mhd ? L.row_bytes * 6 : 0; }
  L.off_shell = off;
  if (shell_mode(d)) {
    size_t padded = 1;
    for (int a = 0; a < d->ndim; ++a) padded *= (size_t)(d->n[a] + 2 * d->num_ghost);
    off += 256 + align256(padded * L.esz * L.ng);  // 256-byte stamp (ShellStamp), then the padded gas rows
  }
  L.total = off;
  return L;
}

Geom geom_of(const amx_desc* d) {
  Geom g;
  g.nd = d->ndim;
  for (int a = 0; a < 3; ++a) {
    g.n[a] = d->n[a];
    g.bc[a][0] = d->bc[a][0];
    g.bc[a][1] = d->bc[a][1];
  }
  g.jet_lo = g.jet_hi = 0;
  g.jet_off = d->slab_flags & 1;
  if (d->ndim == 2 && d->bc[1][0] == 3) {
    // boundaries.py:412-413,424-428: half_inj_cell_num = int(0.025 / grid_spacing); padded y indices
    // [num_cells // 2 - h, num_cells // 2 + h) -> unpadded coordinates
    const int h = (int)(0.025 / d->dx);
    g.jet_lo = d->n[1] / 2 - h - d->num_ghost;
    g.jet_hi = d->n[1] / 2 + h - d->num_ghost;
  }
  return g;
}

NumPar numpar_of(const amx_desc* d) {
  NumPar p;
  p.dtype = d->dtype;
  p.rs = d->riemann;
  p.lim = d->limiter;
  p.dx = d->dx;
  // params.gamma enters the jitted loop as a scalar of the working dtype (Q20)
  p.gamma = d->dtype == AMX_F32 ? (double)(float)d->gamma : d->gamma;
  // split sweeps call _calculate_limited_gradients only: VAN_ALBADA_PP is plain van Albada there
  if (d->limiter == 5 && d->split != AMX_UNSPLIT) p.lim = 4;
  // reconstruction.py:190-197: q = 2 / C_cfl (MHD: the hydro part runs half steps) or 1 / C_cfl
  const double two = d->mhd ? 2.0 : 1.0;
  p.pp_q = d->dtype == AMX_F32 ? (double)((float)two / (float)d->C_cfl) : two / d->C_cfl;
  return p;
}

// View of `rows` rows over a buffer whose per-axis extents are n[a] + 2*pad[a]; p is moved to
// the first active cell.  row_stride_elems = 0 -> dense rows.
View make_view(void* base, const amx_desc* d, const int pad[3], size_t row_stride_elems, int first_row = 0) {
  View v;
  const long long e2 = d->n[2] + 2 * pad[2], e1 = d->n[1] + 2 * pad[1], e0 = d->n[0] + 2 * pad[0];
  long long s[3] = {e1 * e2, e2, 1};
  // strides follow the array rank: 1-D arrays are [rows][x], 2-D [rows][x][y]
  if (d->ndim == 1) { s[0] = 1; s[1] = 0; s[2] = 0; }
  else if (d->ndim == 2) { s[0] = e1; s[1] = 1; s[2] = 0; }
  for (int a = 0; a < 3; ++a) v.s[a] = s[a];
  long long vol = 1;
  if (d->ndim == 1) vol = e0; else if (d->ndim == 2) vol = e0 * e1; else vol = e0 * e1 * e2;
  v.vs = row_stride_elems ? (long long)row_stride_elems : vol;
  long long off = 0;
  for (int a = 0; a < d->ndim; ++a) off += (long long)pad[a] * v.s[a];
  const size_t esz = d->dtype == AMX_F32 ? 4 : 8;
  v.p = static_cast<char*>(base) + ((size_t)first_row * v.vs + off) * esz;
  return v;
}

View rows_of(const View& v, int first_row, size_t esz) {
  View r = v;
  r.p = static_cast<char*>(v.p) + (size_t)first_row * v.vs * esz;
  return r;
}

void seam_pad(const amx_desc* d, int pad[3]) {
  for (int a = 0; a < 3; ++a) pad[a] = (a < d->ndim && d->boundary_handling == AMX_GHOST_CELLS) ? d->num_ghost : 0;
}
void internal_pad(const amx_desc* d, int pad[3]) {
  for (int a = 0; a < 3; ++a) pad[a] = 0;
  if (d->bc[0][0] == AMX_BC_HALO || d->bc[0][1] == AMX_BC_HALO) pad[0] = d->num_ghost;
}

// ---------------------------------------------------------------- control kernels
template <typename T> __device__ inline T from_bits(unsigned long long b);
template <> __device__ inline float from_bits<float>(unsigned long long b) { return __uint_as_float((unsigned)b); }
template <> __device__ inline double from_bits<double>(unsigned long long b) { return __longlong_as_double((long long)b); }

// slots[NV][kAlphaSlots] -> NV maxima (widened to double in dst_d, and as T in dst_t if given); slots zeroed
template <typename T, int NV>
__global__ void finalize_slots_kernel(unsigned long long* slots, double* dst_d, float* dst_f, T* dst_t,
                                      const Ctl* ctl_ro) {
  if (ctl_ro && ctl_ro->done) return;
  __shared__ T red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < NV; ++k) {
    T m = T(0);
    for (int i = threadIdx.x; i < kAlphaSlots; i += blockDim.x) {
      const T v = from_bits<T>(slots[(size_t)k * kAlphaSlots + i]);
      m = v > m ? v : m;  // NaN never wins a comparison; a NaN state shows up in the state itself
      slots[(size_t)k * kAlphaSlots + i] = 0ull;
    }
    for (int o = 16; o > 0; o >>= 1) { const T w = __shfl_xor_sync(0xffffffffu, m, o); m = w > m ? w : m; }
    if (lane == 0) red[wid] = m;
    __syncthreads();
    if (wid == 0) {
      m = lane < (int)(blockDim.x >> 5) ? red[lane] : T(0);
      for (int o = 16; o > 0; o >>= 1) { const T w = __shfl_xor_sync(0xffffffffu, m, o); m = w > m ? w : m; }
      if (lane == 0) {
        if (dst_d) dst_d[k] = (double)m;
        if (dst_f) dst_f[k] = (float)m;
        if (dst_t) dst_t[k] = m;
      }
    }
    __syncthreads();
  }
}

struct DtPar {
  int nd, split, use_max, exact_end;
  double C_cfl, dx, dt_max, t_end;
  double t_stop;  // the loop also ends once time >= t_stop (no dt clipping against it); +inf: off
};

// _timestep_estimator.py:118-128 (unsplit) / :131-218 (split), time_integration.py:677-678
template <typename T>
__device__ T dt_from_alpha(const double* alpha, const DtPar& p, T time) {
  const T C = (T)p.C_cfl;
  T dt;
  if (p.split == AMX_UNSPLIT) {
    T sum = T(0);
    for (int k = 0; k < p.nd; ++k) sum = sum + (T)alpha[k] / (T)p.dx;
    dt = (C * T(1)) / sum;
  } else {
    T m = (T)alpha[0];
    for (int k = 1; k < p.nd; ++k) m = (T)alpha[k] > m ? (T)alpha[k] : m;
    dt = (C * (T)p.dx) / m;
    if (p.use_max) dt = dt < (T)p.dt_max ? dt : (T)p.dt_max;
  }
  if (p.exact_end) {
    const T rem = (T)p.t_end - time;
    dt = dt < rem ? dt : rem;
  }
  return dt;
}

template <typename T>
__global__ void step_begin_kernel(Ctl* ctl, DtPar p, long long max_steps) {
  if (ctl->done) return;
  const T time = (T)ctl->time_d;
  if (!(time < (T)p.t_end) || !(time < (T)p.t_stop) || (max_steps >= 0 && ctl->steps >= max_steps)) {
    ctl->done = 1;
    return;
  }
  const T dt = dt_from_alpha<T>(ctl->alpha, p, time);
  ctl->dt_d = (double)dt;
  ctl->dt_f = (float)dt;
  ctl->mag_active = 1;
  ctl->mag_iter = 0;
  ctl->mag_fix = 0;
  ctl->final_emitted = 0;
}

template <typename T>
__global__ void set_time_kernel(Ctl* ctl, double t0) {
  const T t = (T)t0;
  ctl->time_d = (double)t;
  ctl->time_f = (float)t;
}

template <typename T>
__global__ void step_end_kernel(Ctl* ctl) {
  if (ctl->done) return;
  const T t = (T)ctl->time_d + (T)ctl->dt_d;
  ctl->time_d = (double)t;
  ctl->time_f = (float)t;
  ctl->steps += 1;
  ctl->mag_iters_total += ctl->mag_iter;
}

template <typename T>
__global__ void dt_only_kernel(const T* alpha_t, const T* time_t, DtPar p, T* dt_out) {
  double a[3] = {0, 0, 0};
  for (int k = 0; k < p.nd; ++k) a[k] = (double)alpha_t[k];
  const T time = time_t ? *time_t : T(0);
  DtPar q = p;
  if (!time_t) q.exact_end = 0;
  *dt_out = dt_from_alpha<T>(a, q, time);
}

// _magnetic_field_update.py:127-141 while_condition.
// Slot finalisation and while-condition in one launch, after a fixed-point launch that made
// `nevals` evaluations (1: mag_iter_*, 2: mag_pair_vec) into iterate buffer `nb`, `c0` evaluations
// having been completed before it.  The condition is evaluated after EACH evaluation, as the
// reference does.  A pair whose first evaluation already ends the loop leaves R_{k+2} in the
// buffer: mag_fix asks for that one evaluation to be re-run alone into the same buffer.
template <typename T>
__global__ void mag_check_fused_kernel(Ctl* ctl, double tol, int c0, int nevals, int nb, int emit) {
  if (ctl->done || !ctl->mag_active) return;  // the launch before this one exited without touching the slots
  __shared__ T red[4][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < 2 * nevals; ++k) {
    unsigned long long* slots = k < 2 ? &ctl->change_slots[k][0] : &ctl->change2_slots[k - 2][0];
    T m = T(0);
    for (int i = threadIdx.x; i < kAlphaSlots; i += blockDim.x) {
      const T v = from_bits<T>(slots[i]);
      m = v > m ? v : m;
      slots[i] = 0ull;
    }
    for (int o = 16; o > 0; o >>= 1) { const T w = __shfl_xor_sync(0xffffffffu, m, o); m = w > m ? w : m; }
    if (lane == 0) red[k][wid] = m;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  T n[4] = {T(0), T(0), T(0), T(0)};
  for (int k = 0; k < 2 * nevals; ++k)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) n[k] = red[k][i] > n[k] ? red[k][i] : n[k];
  ctl->change[0] = (double)n[0];
  ctl->change[1] = (double)n[1];
  if (nevals == 2) ctl->pair_runs += 1; else ctl->single_runs += 1;
  const T m1 = n[0] > n[1] ? n[0] : n[1];
  if (!(m1 > (T)tol && ctl->mag_iter < 1000)) {
    ctl->mag_active = 0;
    ctl->mag_final = nb;
    if (m1 > (T)tol) ctl->not_converged = 1;
    if (nevals == 2) { ctl->mag_fix = 1; ctl->mag_fix_base = c0; ctl->single_runs += 1; }
    return;
  }
  ctl->mag_iter += 1;
  ctl->mag_last_buf = nb;
  if (nevals == 2) {
    const T m2 = n[2] > n[3] ? n[2] : n[3];
    ctl->change[0] = (double)n[2];
    ctl->change[1] = (double)n[3];
    if (m2 > (T)tol && ctl->mag_iter < 1000) {
      ctl->mag_iter += 1;
    } else {
      ctl->mag_active = 0;
      ctl->mag_final = nb;
      if (m2 > (T)tol) ctl->not_converged = 1;
      ctl->final_emitted = emit;  // the launch also wrote the write-back of exactly this iterate
    }
  }
}

// After the fixed-point launches a step was given.  mode 0 (amx_integrate's pipeline, which never waits for the
// device inside a step): still iterating -> raise `stalled` and `done`; every kernel enqueued behind exits at once
// and the host resumes the step when it looks.  mode 1 (amx_step without stats: a fixed launch budget and no host
// round trip, so that the call can be captured into a CUDA graph / XLA command buffer): still iterating -> the
// write-back takes the latest iterate and the sticky not_converged flag is raised (amx_magnetic_state reads it).
__global__ void mag_guard_kernel(Ctl* ctl, int mode) {
  if (ctl->done || !ctl->mag_active) return;
  if (mode == 0) {
    ctl->stalled = 1;
    ctl->done = 1;
  } else {
    ctl->mag_active = 0;
    ctl->mag_final = ctl->mag_last_buf;
    ctl->final_emitted = 0;
    ctl->not_converged = 1;
  }
}
__global__ void clear_stall_kernel(Ctl* ctl) {
  ctl->stalled = 0;
  ctl->done = 0;
}

// the same condition for slab ranks, on norms that were all-reduced across ranks (amx_magnetic_check)
template <typename T>
__global__ void mag_check_slab_kernel(Ctl* ctl, const T* change, double tol, int j) {
  if (j == 0) { ctl->mag_active = 1; ctl->mag_iter = 0; ctl->mag_fix = 0; ctl->final_emitted = 0; }
  if (!ctl->mag_active) return;
  const T m = change[0] > change[1] ? change[0] : change[1];
  if (m > (T)tol && ctl->mag_iter < 1000) {
    ctl->mag_iter += 1;
  } else {
    ctl->mag_active = 0;
    ctl->mag_final = j & 1;
    if (m > (T)tol) ctl->not_converged = 1;
  }
}

// pair launches of slab ranks: the condition after each of the two evaluations, on all-reduced norms
template <typename T>
__global__ void mag_check_pair_slab_kernel(Ctl* ctl, const T* n, double tol, int c0, int nb, int emit) {
  if (c0 == 0) { ctl->mag_active = 1; ctl->mag_iter = 0; ctl->mag_fix = 0; ctl->final_emitted = 0; }
  if (!ctl->mag_active) return;
  const T m1 = n[0] > n[1] ? n[0] : n[1];
  if (!(m1 > (T)tol && ctl->mag_iter < 1000)) {
    ctl->mag_active = 0;
    ctl->mag_final = nb;
    if (m1 > (T)tol) ctl->not_converged = 1;
    ctl->mag_fix = 1;
    ctl->mag_fix_base = c0;
    return;
  }
  ctl->mag_iter += 1;
  const T m2 = n[2] > n[3] ? n[2] : n[3];
  if (m2 > (T)tol && ctl->mag_iter < 1000) {
    ctl->mag_iter += 1;
  } else {
    ctl->mag_active = 0;
    ctl->mag_final = nb;
    if (m2 > (T)tol) ctl->not_converged = 1;
    ctl->final_emitted = emit;
  }
}

template <typename T>
__global__ void set_dt_kernel(Ctl* ctl, const T* dt_dev) {
  ctl->dt_d = (double)*dt_dev;
  ctl->dt_f = (float)*dt_dev;
  ctl->done = 0;
  ctl->mag_active = 1;
  ctl->mag_iter = 0;
  ctl->mag_fix = 0;
  ctl->final_emitted = 0;
}

// ---------------------------------------------------------------- layout kernels
// dst (padded by ng on active axes) <- src (unpadded).  Gas rows take the boundary rule of
// their side, the other rows are edge-padded (np.pad mode="edge", Q9).
// ---- peer-memory MAX all-reduce / barrier of a slab run (amx_peer_allreduce_max) ------------------------
// Communication buffer of a rank (peer-mapped, zeroed once): epoch counter, per parity a flag per source
// rank and 8 value slots per source rank.  Call number e (parity e & 1): every rank stores its values into
// slot [parity][own rank] of EVERY rank's buffer, fences, then raises flag [parity][own rank] = e on every
// rank (release, system scope); it waits until all `world` flags of its own buffer read e (acquire) and
// takes the maximum of the slots.  A rank cannot start call e + 2 (which reuses the parity) before every
// rank has raised its flag of call e + 1, i.e. has finished reading the slots of call e.
struct PeerBufs {
  unsigned long long p[8];
};
constexpr size_t kPeerFlagsOff = 64, kPeerValsOff = 128, kPeerCommBytes = 128 + 2 * 8 * 8 * sizeof(double);

template <typename T>
__global__ void __launch_bounds__(64) peer_allreduce_max_kernel(PeerBufs bufs, int rank, int world, T* vals, int count) {
  char* const mine = reinterpret_cast<char*>(bufs.p[rank]);
  __shared__ unsigned e_s;
  __shared__ int timed_out;
  const int tid = threadIdx.x;
  if (tid == 0) { e_s = *reinterpret_cast<volatile unsigned*>(mine) + 1u; timed_out = 0; }
  __syncthreads();
  const unsigned e = e_s, par = e & 1u;
  const int r = tid >> 3, i = tid & 7;
  if (r < world && i < count) {
    volatile double* dst = reinterpret_cast<volatile double*>(reinterpret_cast<char*>(bufs.p[r]) + kPeerValsOff) +
                           ((par * 8 + rank) * 8 + i);
    *dst = (double)vals[i];
    __threadfence_system();
  }
  __syncthreads();
  if (tid < world) {
    unsigned* flag = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(bufs.p[tid]) + kPeerFlagsOff) + (par * 8 + rank);
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(e) : "memory");
    const unsigned* own = reinterpret_cast<const unsigned*>(mine + kPeerFlagsOff) + (par * 8 + tid);
    // (bounded: ranks that do not make the same sequence of calls must produce NaNs, not a hung GPU)
    unsigned seen;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(own) : "memory");
      if (clock64() - t0 > 20000000000ll) { timed_out = 1; break; }
    } while (seen != e);
  }
  __syncthreads();
  if (tid < count) {
    const volatile double* src = reinterpret_cast<const volatile double*>(mine + kPeerValsOff) + par * 64 + tid;
    double m = src[0];
    for (int q = 1; q < world; ++q) m = fmax(m, src[q * 8]);
    vals[tid] = timed_out ? (T)NAN : (T)m;
  }
  if (tid == 0) *reinterpret_cast<volatile unsigned*>(mine) = e;
}

template <typename T>
__global__ void pad_kernel(Geom g, int ng, int rows, int gas_rows, View src, View dst) {
  long long e[3], tot = 1;
  for (int a = 0; a < 3; ++a) { e[a] = a < g.nd ? g.n[a] + 2 * ng : 1; tot *= e[a]; }
  const long long lin = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lin >= tot) return;
  long long r = lin;
  int i[3];
  i[2] = (int)(r % e[2]); r /= e[2];
  i[1] = (int)(r % e[1]); r /= e[1];
  i[0] = (int)r;
  long long od = 0, os_gas = 0, os_edge = 0;
  unsigned flipmask = 0;  // bit k: gas row k changes sign (one reflective axis each)
  for (int a = 0; a < g.nd; ++a) {
    const int ia = i[a] - ng;
    od += (long long)ia * dst.s[a];
    int ie = ia < 0 ? 0 : (ia >= g.n[a] ? g.n[a] - 1 : ia);
    os_edge += (long long)ie * src.s[a];
    int ig = ia;
    if (ia < 0) {
      const int b = g.bc[a][0];
      ig = b == BC_PERIODIC ? ia + g.n[a] : (b == BC_REFLECTIVE ? (g.nd == 1 ? 0 : -1 - ia) : 0);
      if (b == BC_REFLECTIVE) flipmask |= 1u << (a + 1);
    } else if (ia >= g.n[a]) {
      const int b = g.bc[a][1];
      ig = b == BC_PERIODIC ? ia - g.n[a] : (b == BC_REFLECTIVE ? (g.nd == 1 ? g.n[a] - 1 : 2 * g.n[a] - 1 - ia) : g.n[a] - 1);
      if (b == BC_REFLECTIVE) flipmask |= 1u << (a + 1);
    }
    os_gas += (long long)ig * src.s[a];
  }
  const T* ps = static_cast<const T*>(src.p);
  T* pd = static_cast<T*>(dst.p);
  for (int k = 0; k < rows; ++k) {
    T v;
    if (k < gas_rows) {
      v = ps[k * src.vs + os_gas];
      if (flipmask & (1u << k)) v = -v;
      if (g.nd == 2 && jet_cell(g, i[0] - ng, i[1] - ng)) v = jet_gas_value<T>(k);
    } else {
      v = ps[k * src.vs + os_edge];
    }
    pd[k * dst.vs + od] = v;
  }
}

template <typename T>
__global__ void copy_rows_kernel(Geom g, int rows, View src, View dst) {
  const long long tot = (long long)g.n[0] * g.n[1] * g.n[2];
  const long long lin = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lin >= tot) return;
  long long r = lin;
  int i[3];
  i[2] = (int)(r % g.n[2]); r /= g.n[2];
  i[1] = (int)(r % g.n[1]); r /= g.n[1];
  i[0] = (int)r;
  long long os = 0, od = 0;
  for (int a = 0; a < g.nd; ++a) { os += (long long)i[a] * src.s[a]; od += (long long)i[a] * dst.s[a]; }
  const T* ps = static_cast<const T*>(src.p);
  T* pd = static_cast<T*>(dst.p);
  for (int k = 0; k < rows; ++k) pd[k * dst.vs + od] = ps[k * src.vs + os];
}

// _boundary_handler on a padded array, in place: every ghost cell <- its source cell under the
// per-axis rules (open, reflective, periodic compose axis by axis exactly like the reference's
// sequential fills).  type_handled 0: gas rows (flip row `axis`); 1/2: 3-vectors (flip row
// `axis` when < rows, i.e. the reference's Q12 behaviour).
template <typename T>
__global__ void ghost_fill_kernel(Geom g, int ng, int rows, int type_handled, View a) {
  long long e[3], tot = 1;
  for (int ax = 0; ax < 3; ++ax) { e[ax] = ax < g.nd ? g.n[ax] + 2 * ng : 1; tot *= e[ax]; }
  const long long lin = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lin >= tot) return;
  long long r = lin;
  int i[3];
  i[2] = (int)(r % e[2]); r /= e[2];
  i[1] = (int)(r % e[1]); r /= e[1];
  i[0] = (int)r;
  long long od = 0, os = 0;
  unsigned flipmask = 0;
  bool ghost = false;
  for (int ax = 0; ax < g.nd; ++ax) {
    const int ia = i[ax] - ng;
    od += (long long)ia * a.s[ax];
    int ig = ia;
    if (ia < 0) {
      ghost = true;
      const int b = g.bc[ax][0];
      // (the MHD jet side itself is an open copy, boundaries.py:407-410)
      ig = b == BC_PERIODIC ? ia + g.n[ax]
                            : (b == BC_REFLECTIVE ? (g.nd == 1 ? 0 : -1 - ia) : ((b == BC_OPEN || b == BC_JET) ? 0 : ia));
      if (b == BC_REFLECTIVE) flipmask |= 1u << (ax + 1);
    } else if (ia >= g.n[ax]) {
      ghost = true;
      const int b = g.bc[ax][1];
      ig = b == BC_PERIODIC ? ia - g.n[ax]
                            : (b == BC_REFLECTIVE ? (g.nd == 1 ? g.n[ax] - 1 : 2 * g.n[ax] - 1 - ia) : (b == BC_OPEN ? g.n[ax] - 1 : ia));
      if (b == BC_REFLECTIVE) flipmask |= 1u << (ax + 1);
    }
    os += (long long)ig * a.s[ax];
  }
  const bool jet = g.nd == 2 && jet_cell(g, i[0] - ng, i[1] - ng);
  if (!ghost || (os == od && !jet)) return;
  T* p = static_cast<T*>(a.p);
  for (int k = 0; k < rows; ++k) {
    T v = p[k * a.vs + os];
    if (flipmask & (1u << k)) v = -v;
    if (jet) v = type_handled == 0 ? jet_gas_value<T>(k) : jet_vector_value<T>(type_handled, k);
    p[k * a.vs + od] = v;
  }
}

unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

// ---------------------------------------------------------------- engine
#define CKE(expr)                                  \
  do {                                             \
    cudaError_t e__ = (expr);                      \
    if (e__ != cudaSuccess) return e__;            \
  } while (0)

// AMX_FORCE_GENERIC=1 keeps every launch on the one-thread-per-cell kernels (A/B testing)
bool force_generic() {
  static const bool v = [] { const char* e = getenv("AMX_FORCE_GENERIC"); return e && e[0] == '1'; }();
  return v;
}

// AMX_NO_PAIR=1 keeps the fixed point on one evaluation per launch (amx_pair.cu off)
bool no_pair() {
  static const bool v = [] { const char* e = getenv("AMX_NO_PAIR"); return e && e[0] == '1'; }();
  return v;
}

// AMX_PAIR_GREEDY=1 enqueues pair launches regardless of the expected evaluation count, so an odd
// count ends on the first evaluation of a pair and exercises the Ctl::mag_fix re-run (tests)
bool pair_greedy() {
  static const bool v = [] { const char* e = getenv("AMX_PAIR_GREEDY"); return e && e[0] == '1'; }();
  return v;
}

// AMX_NO_VEC=1 skips the z-vectorised kernels (amx_vec.cu) in favour of the scalar tiled ones
bool no_vec() {
  static const bool v = [] { const char* e = getenv("AMX_NO_VEC"); return e && e[0] == '1'; }();
  return v;
}

struct Engine {
  const amx_desc* d;
  Layout L;
  Geom g;
  NumPar np;
  char* ws;
  Ctl* ctl;
  cudaStream_t st;
  long long launches = 0;
  bool strict, is32, lf;  // lf: every stage needs the wave bounds of its input (Lax-Friedrichs alpha, PP limiter C)
  // ghost shell (ShellArgs): where the shell is read / written by the next update; off unless set
  bool shell_on = false;
  View shell_src, shell_dst;
  Peers peers;  // in-kernel halo stores (slab ranks; amx_desc::peer_*)

  Engine(const amx_desc* d_, void* ws_, void* stream)
      : d(d_), L(make_layout(d_)), g(geom_of(d_)), np(numpar_of(d_)), ws(static_cast<char*>(ws_)),
        ctl(reinterpret_cast<Ctl*>(static_cast<char*>(ws_) + L.off_ctl)), st(static_cast<cudaStream_t>(stream)),
        strict(d_->arithmetic == AMX_ARITH_STRICT), is32(d_->dtype == AMX_F32),
        lf(d_->riemann == AMX_LAX_FRIEDRICHS || (d_->limiter == 5 && d_->split == AMX_UNSPLIT)) {
    peers.lo = d_->peer_delta_lo; peers.hi = d_->peer_delta_hi; peers.flags = d_->peer_flags & 3;
  }

  // view over an internal workspace buffer (x ghost planes resident in slab mode)
  View internal(size_t off, int first_row = 0) const {
    int pad[3];
    internal_pad(d, pad);
    return make_view(ws + off, d, pad, L.row_bytes / L.esz, first_row);
  }
  // view over a caller array in the internal layout (dense rows)
  View user_internal(const void* p, int first_row = 0) const {
    int pad[3];
    internal_pad(d, pad);
    return make_view(const_cast<void*>(p), d, pad, 0, first_row);
  }
  View user_seam(const void* p, int first_row = 0) const {
    int pad[3];
    seam_pad(d, pad);
    if (d->bc[0][0] == AMX_BC_HALO || d->bc[0][1] == AMX_BC_HALO) internal_pad(d, pad);
    return make_view(const_cast<void*>(p), d, pad, 0, first_row);
  }
  View user_unpadded(const void* p, int first_row = 0) const {
    int pad[3] = {0, 0, 0};
    return make_view(const_cast<void*>(p), d, pad, 0, first_row);
  }
  const void* dt_ptr() const { return is32 ? (const void*)&ctl->dt_f : (const void*)&ctl->dt_d; }
  const void* alpha_ptr() const { return is32 ? (const void*)ctl->alpha_f : (const void*)ctl->alpha; }
  unsigned long long* aslots() const { return &ctl->alpha_slots[0][0]; }
  unsigned long long* cslots() const { return &ctl->change_slots[0][0]; }
  DtPar dtpar(double t_end) const {
    DtPar p;
    p.nd = d->ndim; p.split = d->split; p.use_max = d->use_max_adaptive_timestep; p.exact_end = d->exact_end_time;
    p.C_cfl = is32 ? (double)(float)d->C_cfl : d->C_cfl;
    p.dx = d->dx;
    p.dt_max = is32 ? (double)(float)d->dt_max : d->dt_max;
    p.t_end = is32 ? (double)(float)t_end : t_end;
    p.t_stop = HUGE_VAL;
    return p;
  }

  cudaError_t finalize_alpha(void* alpha_t_out, bool guarded) {
    ++launches;
    const Ctl* c = guarded ? ctl : nullptr;
    if (is32) finalize_slots_kernel<float, 3><<<1, 256, 0, st>>>(aslots(), ctl->alpha, ctl->alpha_f, (float*)alpha_t_out, c);
    else finalize_slots_kernel<double, 3><<<1, 256, 0, st>>>(aslots(), ctl->alpha, ctl->alpha_f, (double*)alpha_t_out, c);
    return cudaGetLastError();
  }
  cudaError_t finalize_change(void* change_t_out, bool guarded) {
    ++launches;
    const Ctl* c = guarded ? ctl : nullptr;
    if (is32) finalize_slots_kernel<float, 2><<<1, 256, 0, st>>>(cslots(), ctl->change, nullptr, (float*)change_t_out, c);
    else finalize_slots_kernel<double, 2><<<1, 256, 0, st>>>(cslots(), ctl->change, nullptr, (double*)change_t_out, c);
    return cudaGetLastError();
  }

  cudaError_t finalize_change2(void* change_t_out) {
    ++launches;
    unsigned long long* s2 = &ctl->change2_slots[0][0];
    if (is32) finalize_slots_kernel<float, 2><<<1, 256, 0, st>>>(s2, ctl->change, nullptr, (float*)change_t_out, nullptr);
    else finalize_slots_kernel<double, 2><<<1, 256, 0, st>>>(s2, ctl->change, nullptr, (double*)change_t_out, nullptr);
    return cudaGetLastError();
  }

  // padded_gas: also reduce over the ghost cells of this padded array (view at its first active cell)
  cudaError_t alpha_of(const View& gas, void* alpha_t_out, bool guarded, const View* padded_gas = nullptr) {
    AlphaArgs a{g, np, gas, guarded ? ctl : nullptr, aslots()};
    ++launches;
    {
      ProfScope ps(PC_ALPHA, st);
      CKE(strict ? amx_strict::launch_alpha(a, st) : amx_fast::launch_alpha(a, st));
    }
    if (padded_gas) CKE(shell(2, *padded_gas, *padded_gas, nullptr, nullptr, nullptr, true, guarded));
    return finalize_alpha(alpha_t_out, guarded);
  }

  // one update of the ghost shell (ShellArgs); with_alpha: its wave bounds join the slots
  cudaError_t shell(int mode, const View& src, const View& dst, const View* u0, const View* u1, const View* X,
                    bool with_alpha, bool guarded) {
    ShellArgs a;
    a.g = g; a.np = np; a.mode = mode; a.ng = d->num_ghost;
    a.shell_in = src; a.shell_out = dst;
    a.u0 = u0 ? *u0 : src; a.u1 = u1 ? *u1 : src;
    a.X[0] = X ? X[0] : src; a.X[1] = X ? X[1] : src;
    a.ctl_ro = (guarded || mode == 1) ? ctl : nullptr;
    a.fixed_final = -1;
    a.alpha_slots = with_alpha ? aslots() : nullptr;
    ++launches;
    return strict ? amx_strict::launch_shell(a, st) : amx_fast::launch_shell(a, st);
  }

  cudaError_t stage(const View& in, const View* in0, const View& out, const void* dt, double dt_scale,
                    const void* alpha_lf, bool want_alpha, void* alpha_t_out, bool guarded,
                    bool defer_finalize = false) {
    StageArgs a;
    a.g = g; a.np = np; a.in = in; a.out = out;
    a.in0 = in0 ? *in0 : in;
    if (!in0) a.in0.p = nullptr;
    a.ctl_ro = guarded ? ctl : nullptr;
    a.dt = dt; a.dt_scale = dt_scale; a.alpha_lf = alpha_lf;
    a.alpha_slots = want_alpha ? aslots() : nullptr;
    a.peers = peers;
    ++launches;
    {
      ProfScope ps(PC_HYDRO_STAGE, st);
      const bool tiled = !force_generic() && (strict ? amx_strict::tiled_stage_supported(a) : amx_fast::tiled_stage_supported(a));
      const bool vec = !force_generic() && !no_vec() && (strict ? amx_strict::vec_stage_supported(a) : amx_fast::vec_stage_supported(a));
      if (vec) CKE(strict ? amx_strict::launch_hydro_stage_vec(a, st) : amx_fast::launch_hydro_stage_vec(a, st));
      else if (tiled) CKE(strict ? amx_strict::launch_hydro_stage_tiled(a, st) : amx_fast::launch_hydro_stage_tiled(a, st));
      else CKE(strict ? amx_strict::launch_hydro_stage(a, st) : amx_fast::launch_hydro_stage(a, st));
    }
    return (want_alpha && !defer_finalize) ? finalize_alpha(alpha_t_out, guarded) : cudaSuccess;
  }

  cudaError_t sweep(const View& in, const View& out, int axis, const void* dt, double dt_scale,
                    const void* alpha_lf, bool want_alpha, void* alpha_t_out, bool guarded,
                    bool defer_finalize = false) {
    SweepArgs a;
    a.g = g; a.np = np; a.in = in; a.out = out; a.ctl_ro = guarded ? ctl : nullptr;
    a.dt = dt; a.dt_scale = dt_scale; a.axis = axis; a.first_order = d->first_order_fallback;
    a.alpha_lf = alpha_lf; a.alpha_slots = want_alpha ? aslots() : nullptr;
    ++launches;
    {
      ProfScope ps(PC_SPLIT_SWEEP, st);
      CKE(strict ? amx_strict::launch_split_sweep(a, st) : amx_fast::launch_split_sweep(a, st));
    }
    return (want_alpha && !defer_finalize) ? finalize_alpha(alpha_t_out, guarded) : cudaSuccess;
  }

  // evolve_state.py:419-488 / :168-325: gas_in -> gas_out with time step dt*scale.
  // tmpA / tmpB are scratch gas buffers (internal layout).  final_alpha: also produce the wave
  // bounds of gas_out (next CFL; always needed by Lax-Friedrichs).
  cudaError_t hydro(const View& gas_in, const View& gas_out, const View& tmpA, const View& tmpB,
                    double scale, bool final_alpha) {
    if (d->split == AMX_UNSPLIT) {
      CKE(stage(gas_in, nullptr, tmpA, dt_ptr(), scale, alpha_ptr(), lf, nullptr, true));
      if (!shell_on) return stage(tmpA, &gas_in, gas_out, dt_ptr(), scale, alpha_ptr(), final_alpha || lf, nullptr, true);
      // shell mode: the ghost shell follows the half step, and the wave bounds the NEXT CFL reads
      // (final_alpha) are those of the padded array
      CKE(stage(tmpA, &gas_in, gas_out, dt_ptr(), scale, alpha_ptr(), final_alpha || lf, nullptr, true, true));
      CKE(shell(0, shell_src, shell_dst, &gas_in, &gas_out, nullptr, final_alpha, true));
      shell_src = shell_dst;
      return (final_alpha || lf) ? finalize_alpha(nullptr, true) : cudaSuccess;
    }
    // Strang sequences of evolve_state.py:168-325
    int axes[5];
    double frac[5];
    int ns;
    if (d->ndim == 1) { ns = 1; axes[0] = 1; frac[0] = 1; }
    else if (d->ndim == 2) { ns = 3; axes[0] = 1; axes[1] = 2; axes[2] = 1; frac[0] = 0.5; frac[1] = 1; frac[2] = 0.5; }
    else { ns = 5; const int ax[5] = {1, 2, 3, 2, 1}; const double fr[5] = {0.5, 0.5, 1, 0.5, 0.5};
           for (int k = 0; k < 5; ++k) { axes[k] = ax[k]; frac[k] = fr[k]; } }
    const View* src = &gas_in;
    for (int k = 0; k < ns; ++k) {
      const bool last = k == ns - 1;
      const View* dst = last ? &gas_out : ((k & 1) ? &tmpB : &tmpA);
      const bool want = (last && final_alpha) || lf;
      const bool defer = last && shell_on;
      CKE(sweep(*src, *dst, axes[k], dt_ptr(), scale * frac[k], alpha_ptr(), want, nullptr, true, defer));
      src = dst;
    }
    if (shell_on) {  // a sweep ends with a ghost fill: the shell is the fill of the result (images / jet constants)
      CKE(shell(3, shell_src, shell_dst, nullptr, &gas_out, nullptr, final_alpha, true));
      shell_src = shell_dst;
      if (final_alpha || lf) return finalize_alpha(nullptr, true);
    }
    return cudaSuccess;
  }

  MagIterArgs mag_args(const View& B0, const View& gas0, const View* Xk, const View& Xk1, const void* dt,
                       int j, bool guarded) const {
    MagIterArgs a;
    a.g = g; a.np = np; a.B0 = B0; a.gas0 = gas0;
    if (Xk) { a.Bk = *Xk; a.vk = rows_of(*Xk, 3, L.esz); }
    else { a.Bk = B0; a.Bk.p = nullptr; a.vk = B0; a.vk.p = nullptr; }
    a.Bk1 = Xk1; a.vk1 = rows_of(Xk1, 3, L.esz);
    a.ctl_ro = guarded ? ctl : nullptr;
    a.first = (j == 0);
    a.midpoint = d->magnetic_integrator == AMX_IMPLICIT_MIDPOINT;
    a.dt = dt; a.change_slots = cslots(); a.change_slots2 = &ctl->change2_slots[0][0];
    a.fix_base = -1;
    a.emit = 0; a.emit_gas = B0; a.emit_B = B0;
    a.peers = peers;
    return a;
  }
  bool pair_supported(const MagIterArgs& a) const {
    if (force_generic() || no_vec() || no_pair()) return false;
    return strict ? amx_strict::pair_mag_supported(a) : amx_fast::pair_mag_supported(a);
  }
  // two evaluations in one launch (amx_pair.cu); c0 = evaluations completed before it
  cudaError_t mag_pair(const View& B0, const View& gas0, const View* Xk, const View& Xk2, const void* dt, int c0,
                       const View* emit_gas = nullptr, const View* emit_B = nullptr) {
    MagIterArgs a = mag_args(B0, gas0, Xk, Xk2, dt, c0, true);
    if (emit_gas) { a.emit = 1; a.emit_gas = *emit_gas; a.emit_B = *emit_B; }
    ++launches;
    ProfScope ps(PC_MAG_PAIR, st);
    return strict ? amx_strict::launch_mag_pair_vec(a, st) : amx_fast::launch_mag_pair_vec(a, st);
  }
  // the single evaluation c0 + 1 re-run into the pair's buffer when the loop ended on it (Ctl::mag_fix)
  cudaError_t mag_fix(const View& B0, const View& gas0, const View* Xk, const View& Xk1, const void* dt, int c0) {
    MagIterArgs a = mag_args(B0, gas0, Xk, Xk1, dt, c0, true);
    a.fix_base = c0;
    a.change_slots = nullptr;
    ++launches;
    ProfScope ps(PC_MAG_ITER, st);
    return strict ? amx_strict::launch_mag_iter_vec(a, st) : amx_fast::launch_mag_iter_vec(a, st);
  }
  cudaError_t mag_check_fused(int c0, int nevals, int nb, int emit = 0) {
    ++launches;
    const double tol = d->fp_tol;
    if (is32) mag_check_fused_kernel<float><<<1, 256, 0, st>>>(ctl, tol, c0, nevals, nb, emit);
    else mag_check_fused_kernel<double><<<1, 256, 0, st>>>(ctl, tol, c0, nevals, nb, emit);
    return cudaGetLastError();
  }

  cudaError_t mag_iter(const View& B0, const View& gas0, const View* Xk, const View& Xk1, const void* dt,
                       int j, bool guarded) {
    const MagIterArgs a = mag_args(B0, gas0, Xk, Xk1, dt, j, guarded);
    ++launches;
    ProfScope ps(PC_MAG_ITER, st);
    const bool vec = !force_generic() && !no_vec() && (strict ? amx_strict::vec_mag_supported(a) : amx_fast::vec_mag_supported(a));
    if (vec) return strict ? amx_strict::launch_mag_iter_vec(a, st) : amx_fast::launch_mag_iter_vec(a, st);
    const bool tiled = !force_generic() && (strict ? amx_strict::tiled_mag_supported(a) : amx_fast::tiled_mag_supported(a));
    if (tiled) return strict ? amx_strict::launch_mag_iter_tiled(a, st) : amx_fast::launch_mag_iter_tiled(a, st);
    return strict ? amx_strict::launch_mag_iter(a, st) : amx_fast::launch_mag_iter(a, st);
  }

  // gas (rho, v_0, p) + converged iterate -> gas_out (rho, v_n, p_2), B_out; skipped on the device when
  // an emitting pair launch already did it (skip_if_emitted)
  cudaError_t mag_final(const View& gas, const View& gas_out, const View& B_out, const View X[2], int fixed_final,
                        bool want_alpha, void* alpha_t_out, bool skip_if_emitted) {
    MagFinalArgs a;
    a.g = g; a.np = np; a.gas = gas; a.gas_out = gas_out; a.B_out = B_out; a.X[0] = X[0]; a.X[1] = X[1];
    a.ctl_ro = ctl; a.fixed_final = fixed_final; a.alpha_slots = want_alpha ? aslots() : nullptr;
    a.skip_if_emitted = skip_if_emitted ? 1 : 0;
    a.final_runs = &ctl->final_runs;
    a.peers = peers;
    ++launches;
    {
      ProfScope ps(PC_MAG_FINAL, st);
      CKE(strict ? amx_strict::launch_mag_final(a, st) : amx_fast::launch_mag_final(a, st));
    }
    return want_alpha ? finalize_alpha(alpha_t_out, true) : cudaSuccess;
  }
};

struct HostMirror {
  Ctl* pinned = nullptr;  // only the head of Ctl (before the slots) is mirrored
  ~HostMirror() { if (pinned) cudaFreeHost(pinned); }
};
thread_local HostMirror g_mirror;
constexpr size_t kCtlHead = offsetof(Ctl, alpha_slots);

int read_ctl(Engine& E, Ctl** out) {
  if (!g_mirror.pinned) CK(cudaMallocHost(&g_mirror.pinned, sizeof(Ctl)));
  CK(cudaMemcpyAsync(g_mirror.pinned, E.ctl, kCtlHead, cudaMemcpyDeviceToHost, E.st));
  CK(cudaStreamSynchronize(E.st));
  *out = g_mirror.pinned;
  return AMX_OK;
}

// One full step state_in -> state_out (views of nv rows) with dt already in ctl: evolve_state.py:497-600.
// MHD is Strang: hydro(dt/2) on the gas rows . magnetic(dt) . hydro(dt/2); the step is enqueued in two parts
// around the fixed point so that amx_integrate can keep the device fed without looking at it.
struct StepPlan {
  View A, G, Gf, Bin, Bout, X[2];
  bool pairs = false, emit_ok = false;
  int expect = 1;
  int c = 0, nl = 0;  // evaluations / launches enqueued so far in this step
};

// one fixed-point launch (_magnetic_field_update.py:82-205) + the device-side while-condition.  Launch number L
// writes iterate buffer X[L & 1]; launches past convergence exit at once.
int fp_enqueue(Engine& E, StepPlan& P, bool pair) {
  const View* src = P.nl == 0 ? nullptr : &P.X[(P.nl - 1) & 1];
  const View& dst = P.X[P.nl & 1];
  if (pair) {
    // The pair launch expected to end the loop also writes the write-back of its result (gas -> Gf, B -> the
    // output state), which saves the separate mag_final pass (18 words per cell) whenever the expectation holds
    const bool emit = P.emit_ok && P.c + 2 == P.expect;
    CK(E.mag_pair(P.Bin, P.G, src, dst, E.dt_ptr(), P.c, emit ? &P.Gf : nullptr, emit ? &P.Bout : nullptr));
    CK(E.mag_check_fused(P.c, 2, P.nl & 1, emit ? 1 : 0));
    CK(E.mag_fix(P.Bin, P.G, src, dst, E.dt_ptr(), P.c));
    P.c += 2;
  } else {
    CK(E.mag_iter(P.Bin, P.G, src, dst, E.dt_ptr(), P.nl, true));
    CK(E.mag_check_fused(P.c, 1, P.nl & 1));
    P.c += 1;
  }
  ++P.nl;
  return AMX_OK;
}

// first part: hydro half step and the fixed-point launches the previous step needed (*kspec evaluations, as pair
// launches where the kernel supports the problem, amx_pair.cu) plus `slack` more launches
int step_front(Engine& E, const View& sin, const View& sout, int kspec, int slack, StepPlan& P) {
  const amx_desc* d = E.d;
  const Layout& L = E.L;
  P.A = E.internal(L.off_A);
  P.G = E.internal(L.off_G);
  if (!d->mhd) {
    CK(E.hydro(sin, sout, P.A, P.G, 1.0, true));
    return AMX_OK;
  }
  P.X[0] = E.internal(L.off_X[0]);
  P.X[1] = E.internal(L.off_X[1]);
  P.Gf = E.internal(L.off_Gf);
  P.Bin = rows_of(sin, L.ng, L.esz);
  P.Bout = rows_of(sout, L.ng, L.esz);
  // the second half step needs two scratch buffers besides G: A and the gas rows of ... X is free by then
  CK(E.hydro(sin, P.G, P.A, P.X[0], 0.5, false));
  P.pairs = E.pair_supported(E.mag_args(P.Bin, P.G, nullptr, P.X[0], E.dt_ptr(), 0, true));
  // Lax-Friedrichs / PP need the wave bounds of the written-back gas and keep mag_final.  AMX_NO_EMIT=1: off.
  static const bool no_emit = [] { const char* e = getenv("AMX_NO_EMIT"); return e && e[0] == '1'; }();
  P.emit_ok = P.pairs && !E.lf && !no_emit && !E.shell_on;
  if (P.emit_ok) {
    MagIterArgs probe = E.mag_args(P.Bin, P.G, nullptr, P.X[0], E.dt_ptr(), 0, true);
    probe.emit = 1; probe.emit_gas = P.Gf; probe.emit_B = P.Bout;
    P.emit_ok = E.pair_supported(probe);
  }
  P.expect = std::max(1, kspec);
  P.c = P.nl = 0;
  int rc;
  while (P.c < P.expect)
    if ((rc = fp_enqueue(E, P, P.pairs && (P.expect - P.c >= 2 || pair_greedy())))) return rc;
  for (int k = 0; k < slack; ++k)
    if ((rc = fp_enqueue(E, P, P.pairs))) return rc;
  return AMX_OK;
}

// second part: write-back, ghost shell, hydro half step
int step_back(Engine& E, const View& sout, StepPlan& P) {
  const amx_desc* d = E.d;
  if (!d->mhd) return AMX_OK;
  CK(E.mag_final(P.G, P.Gf, P.Bout, P.X, -1, E.lf, nullptr, true));
  if (E.shell_on && d->split == AMX_UNSPLIT) CK(E.shell(1, E.shell_src, E.shell_dst, &P.G, nullptr, P.X, false, true));
  // X buffers are free again: reuse X[0] rows as sweep scratch
  CK(E.hydro(P.Gf, sout, P.A, P.X[0], 0.5, true));
  return AMX_OK;
}

// the host looks until the fixed point has ended (one round trip per extra launch): the slow path of a step whose
// enqueued launches did not suffice, and amx_step with stats
int fp_finish_blocking(Engine& E, StepPlan& P, int* kspec) {
  Ctl* m;
  int rc = read_ctl(E, &m);
  if (rc) return rc;
  while (m->mag_active && !m->done) {
    if ((rc = fp_enqueue(E, P, P.pairs))) return rc;
    if (!P.pairs && (rc = fp_enqueue(E, P, false))) return rc;
    if ((rc = read_ctl(E, &m))) return rc;
  }
  if (!m->done) *kspec = m->mag_iter + 1;
  return AMX_OK;
}

int run_step(Engine& E, const View& sin, const View& sout, int* kspec) {
  StepPlan P;
  int rc = step_front(E, sin, sout, *kspec, 1, P);
  if (rc) return rc;
  if (E.d->mhd && (rc = fp_finish_blocking(E, P, kspec))) return rc;
  return step_back(E, sout, P);
}

int check_common(const amx_desc* d, const void* a, const void* b) {
  int rc = validate(d);
  if (rc) return rc;
  if (!a || !b) return fail(AMX_EINVAL, "null array pointer");
  return AMX_OK;
}

template <typename F>
int launch_by_dtype(const amx_desc* d, F&& f) {
  return d->dtype == AMX_F32 ? f(float{}) : f(double{});
}

}  // namespace

// ================================================================ exported C ABI
extern "C" {

int amx_abi_version(void) { return AMX_ABI_VERSION; }
const char* amx_last_error(void) { return g_err.c_str(); }

int amx_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (n == 0) return fail(AMX_ECUDA, "no CUDA device visible; libamx has no CPU path");
  return AMX_OK;
}

int amx_profile_enable(int on) {
  g_prof.resolve();
  g_prof.on = on != 0;
  for (int k = 0; k < PC_COUNT; ++k) { g_prof.ms[k] = 0; g_prof.cnt[k] = 0; }
  return AMX_OK;
}

int amx_profile_read(int kernel_class, double* ms_total, int64_t* launches) {
  if (kernel_class < 0 || kernel_class >= PC_COUNT || !ms_total || !launches) return fail(AMX_EINVAL, "bad kernel_class / null output");
  g_prof.resolve();
  *ms_total = g_prof.ms[kernel_class];
  *launches = g_prof.cnt[kernel_class];
  return AMX_OK;
}

int amx_workspace_bytes(const amx_desc* d, size_t* bytes_out) {
  int rc = validate(d);
  if (rc) return rc;
  if (!bytes_out) return fail(AMX_EINVAL, "bytes_out is null");
  *bytes_out = make_layout(d).total;
  return AMX_OK;
}

int amx_pad(const amx_desc* d, const void* unpadded_in, void* seam_out, void* stream) {
  int rc = check_common(d, unpadded_in, seam_out);
  if (rc) return rc;
  if ((rc = amx_device_check())) return rc;
  Engine E(d, nullptr, stream);
  const View src = E.user_unpadded(unpadded_in), dst = E.user_seam(seam_out);
  if (d->boundary_handling != AMX_GHOST_CELLS) {
    const long long tot = (long long)d->n[0] * d->n[1] * d->n[2];
    if (E.is32) copy_rows_kernel<float><<<blocks_for(tot), 256, 0, E.st>>>(E.g, E.L.nv, src, dst);
    else copy_rows_kernel<double><<<blocks_for(tot), 256, 0, E.st>>>(E.g, E.L.nv, src, dst);
  } else {
    long long tot = 1;
    for (int a = 0; a < d->ndim; ++a) tot *= d->n[a] + 2 * d->num_ghost;
    if (E.is32) pad_kernel<float><<<blocks_for(tot), 256, 0, E.st>>>(E.g, d->num_ghost, E.L.nv, E.L.ng, src, dst);
    else pad_kernel<double><<<blocks_for(tot), 256, 0, E.st>>>(E.g, d->num_ghost, E.L.nv, E.L.ng, src, dst);
  }
  CK(cudaGetLastError());
  return AMX_OK;
}

int amx_unpad(const amx_desc* d, const void* seam_in, void* unpadded_out, void* stream) {
  int rc = check_common(d, seam_in, unpadded_out);
  if (rc) return rc;
  if ((rc = amx_device_check())) return rc;
  Engine E(d, nullptr, stream);
  const View src = E.user_seam(seam_in), dst = E.user_unpadded(unpadded_out);
  const long long tot = (long long)d->n[0] * d->n[1] * d->n[2];
  if (E.is32) copy_rows_kernel<float><<<blocks_for(tot), 256, 0, E.st>>>(E.g, E.L.nv, src, dst);
  else copy_rows_kernel<double><<<blocks_for(tot), 256, 0, E.st>>>(E.g, E.L.nv, src, dst);
  CK(cudaGetLastError());
  return AMX_OK;
}

int amx_boundary_fill(const amx_desc* d, void* seam_inout, int nrows, int type_handled, void* stream) {
  int rc = check_common(d, seam_inout, seam_inout);
  if (rc) return rc;
  if (nrows < 1 || nrows > 8 || type_handled < 0 || type_handled > 2) return fail(AMX_EINVAL, "bad nrows / type_handled");
  if (d->boundary_handling != AMX_GHOST_CELLS) return AMX_OK;  // nothing to fill (PERIODIC_ROLL)
  if ((rc = amx_device_check())) return rc;
  Engine E(d, nullptr, stream);
  int pad[3];
  seam_pad(d, pad);
  const View a = make_view(seam_inout, d, pad, 0, 0);
  long long tot = 1;
  for (int ax = 0; ax < d->ndim; ++ax) tot *= d->n[ax] + 2 * d->num_ghost;
  // 3-vectors flip row `axis` (1-based) like the gas state does: the kernel's rule is the same
  if (E.is32) ghost_fill_kernel<float><<<blocks_for(tot), 256, 0, E.st>>>(E.g, d->num_ghost, nrows, type_handled, a);
  else ghost_fill_kernel<double><<<blocks_for(tot), 256, 0, E.st>>>(E.g, d->num_ghost, nrows, type_handled, a);
  CK(cudaGetLastError());
  return AMX_OK;
}

int amx_cfl(const amx_desc* d, const void* seam_state, void* dt_out, void* ws, void* stream) {
  int rc = check_common(d, seam_state, dt_out);
  if (rc) return rc;
  if (!ws) return fail(AMX_EINVAL, "ws is null");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  CK(cudaMemsetAsync(E.ctl, 0, sizeof(Ctl), E.st));
  // _timestep_estimator.py:118-128 reduces over the array it is given: in GHOST_CELLS mode that is
  // the padded one, ghost cells included (Q3)
  const View seam = E.user_seam(seam_state);
  const bool padded = d->boundary_handling == AMX_GHOST_CELLS && d->bc[0][0] != AMX_BC_HALO && d->bc[0][1] != AMX_BC_HALO;
  CK(E.alpha_of(seam, nullptr, false, padded ? &seam : nullptr));
  DtPar p = E.dtpar(0.0);
  if (E.is32) dt_only_kernel<float><<<1, 1, 0, E.st>>>(E.ctl->alpha_f, nullptr, p, (float*)dt_out);
  else dt_only_kernel<double><<<1, 1, 0, E.st>>>(E.ctl->alpha, nullptr, p, (double*)dt_out);
  CK(cudaGetLastError());
  return AMX_OK;
}

int amx_step(const amx_desc* d, const void* seam_in, void* seam_out, const void* dt_dev, void* ws,
             void* stream, amx_stats* stats) {
  int rc = check_common(d, seam_in, seam_out);
  if (rc) return rc;
  if (!ws || !dt_dev) return fail(AMX_EINVAL, "ws / dt_dev is null");
  if (seam_in == seam_out) return fail(AMX_EINVAL, "in and out must not alias");
  if (d->bc[0][0] == AMX_BC_HALO || d->bc[0][1] == AMX_BC_HALO)
    return fail(AMX_EUNSUPPORTED, "amx_step needs whole-domain boundaries; slab ranks use the stage-level entry points");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  CK(cudaMemsetAsync(E.ctl, 0, sizeof(Ctl), E.st));
  if (E.is32) set_dt_kernel<float><<<1, 1, 0, E.st>>>(E.ctl, (const float*)dt_dev);
  else set_dt_kernel<double><<<1, 1, 0, E.st>>>(E.ctl, (const double*)dt_dev);
  CK(cudaGetLastError());
  const View sin = E.user_seam(seam_in), sout = E.user_seam(seam_out);
  if (E.lf) CK(E.alpha_of(sin, nullptr, false));
  if (shell_mode(d)) {  // ghost gas cells: read from seam_in, carried through the step in seam_out's ghost cells
    E.shell_on = true;
    E.shell_src = sin;
    E.shell_dst = sout;
  }
  int kspec = d->dtype == AMX_F32 ? 4 : 8;  // first guess of the evaluations per step
  if (stats || !d->mhd) {
    if ((rc = run_step(E, sin, sout, &kspec))) return rc;  // (MHD: the host looks at the fixed point)
  } else {
    // No stats wanted: a fixed budget of fixed-point launches (launches past convergence exit at once) and no
    // host round trip, so the call is a pure stream of launches -- it can be captured into a CUDA graph / XLA
    // command buffer.  A fixed point that is still iterating when the budget is spent (the
    // reference would go on, up to 1000 iterations) raises the sticky not_converged flag, which
    // amx_magnetic_state reports; AMX_STEP_EVALS sets the budget (default 12 evaluations in fp32, 24 in fp64).
    static const int budget_env = [] { const char* e = getenv("AMX_STEP_EVALS"); return e ? atoi(e) : 0; }();
    const int budget = budget_env > 0 ? budget_env : (d->dtype == AMX_F32 ? 12 : 24);
    StepPlan P;
    if ((rc = step_front(E, sin, sout, budget, 0, P))) return rc;
    ++E.launches;
    mag_guard_kernel<<<1, 1, 0, E.st>>>(E.ctl, 1);
    CK(cudaGetLastError());
    if ((rc = step_back(E, sout, P))) return rc;
  }
  if (d->boundary_handling == AMX_GHOST_CELLS) {
    // ghosts of the result: what the reference's trailing _boundary_handler calls leave behind (shell
    // mode: the gas ghosts are the shell and are already in place)
    if (!E.shell_on && (rc = amx_boundary_fill(d, seam_out, E.L.ng, 0, stream))) return rc;
    if (d->mhd) {
      int pad[3];
      seam_pad(d, pad);
      View b = make_view(seam_out, d, pad, 0, E.L.ng);
      long long tot = 1;
      for (int ax = 0; ax < d->ndim; ++ax) tot *= d->n[ax] + 2 * d->num_ghost;
      if (E.is32) ghost_fill_kernel<float><<<blocks_for(tot), 256, 0, E.st>>>(E.g, d->num_ghost, 3, 2, b);
      else ghost_fill_kernel<double><<<blocks_for(tot), 256, 0, E.st>>>(E.g, d->num_ghost, 3, 2, b);
      CK(cudaGetLastError());
    }
  }
  if (stats) {
    Ctl* m;
    if ((rc = read_ctl(E, &m))) return rc;
    stats->steps = 1;
    stats->magnetic_iterations = m->mag_iter;
    stats->time = 0;
    stats->last_dt = m->dt_d;
    stats->kernel_launches = E.launches;
    stats->pair_launches = m->pair_runs;
    stats->single_launches = m->single_runs;
    stats->final_launches = m->final_runs;
    if (m->not_converged) return fail(AMX_ENOTCONVERGED, "magnetic fixed point hit the 1000-iteration cap");
  }
  return AMX_OK;
}

int amx_integrate(const amx_desc* d, const void* unpadded_in, void* unpadded_out, double t_end,
                  int64_t max_steps, void* ws, void* stream, amx_stats* stats) {
  return amx_integrate_from(d, unpadded_in, unpadded_out, 0.0, HUGE_VAL, t_end, max_steps, ws, stream, stats);
}

int amx_integrate_from(const amx_desc* d, const void* unpadded_in, void* unpadded_out, double t_start,
                       double t_stop, double t_end, int64_t max_steps, void* ws, void* stream,
                       amx_stats* stats) {
  int rc = check_common(d, unpadded_in, unpadded_out);
  if (rc) return rc;
  if (!ws) return fail(AMX_EINVAL, "ws is null");
  if (d->bc[0][0] == AMX_BC_HALO || d->bc[0][1] == AMX_BC_HALO)
    return fail(AMX_EUNSUPPORTED, "amx_integrate needs whole-domain boundaries; slab ranks use the stage-level entry points");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  const Layout& L = E.L;
  CK(cudaMemsetAsync(E.ctl, 0, sizeof(Ctl), E.st));
  const View vin = E.user_unpadded(unpadded_in);
  const View S[2] = {E.internal(L.off_S[0]), E.internal(L.off_S[1])};
  if (shell_mode(d)) {
    int pad[3];
    seam_pad(d, pad);
    const View sh = make_view(E.ws + L.off_shell + 256, d, pad, 0, 0);
    const ShellStamp mine = stamp_of(d);
    bool keep = false;
    if (t_start != 0.0) {  // resumed segment on the same ws: keep the shell the previous segment left
      ShellStamp found{};
      CK(cudaMemcpyAsync(&found, E.ws + L.off_shell, sizeof(found), cudaMemcpyDeviceToHost, E.st));
      CK(cudaStreamSynchronize(E.st));
      keep = found == mine;
    }
    if (!keep) {
      // time_integration.py:274-286: pad + ghost fill of the gas rows = images of the interior cells
      CK(cudaMemcpyAsync(E.ws + L.off_shell, &mine, sizeof(mine), cudaMemcpyHostToDevice, E.st));
      long long tot = 1;
      for (int a = 0; a < d->ndim; ++a) tot *= d->n[a] + 2 * d->num_ghost;
      if (E.is32) pad_kernel<float><<<blocks_for(tot), 256, 0, E.st>>>(E.g, d->num_ghost, L.ng, L.ng, vin, sh);
      else pad_kernel<double><<<blocks_for(tot), 256, 0, E.st>>>(E.g, d->num_ghost, L.ng, L.ng, vin, sh);
      CK(cudaGetLastError());
    }
    E.shell_on = true;
    E.shell_src = sh;
    E.shell_dst = sh;
  }
  // wave bounds of the initial state (time_integration.py:643-653 on the first pass)
  CK(E.alpha_of(vin, nullptr, false, E.shell_on ? &E.shell_src : nullptr));
  DtPar p = E.dtpar(t_end);
  p.t_stop = E.is32 ? (double)(float)t_stop : t_stop;
  if (t_start != 0.0) {
    if (E.is32) set_time_kernel<float><<<1, 1, 0, E.st>>>(E.ctl, t_start);
    else set_time_kernel<double><<<1, 1, 0, E.st>>>(E.ctl, t_start);
    CK(cudaGetLastError());
  }
  // Speculative pipeline: the host never waits for the device inside a step.  Every step is enqueued whole --
  // begin, hydro half step, as many fixed-point launches as the previous step needed (+ slack), a guard, the
  // write-back, the second half step, end -- and a copy of the control block taken after the guard is looked at
  // one step (hydro: one batch of 8 steps) LATER, when the device already has the next step queued.  The look
  // tells (a) how many evaluations the step needed (next steps enqueue that many), (b) whether the time loop has
  // ended (everything enqueued since was a no-op), (c) whether the guard found the fixed point still iterating
  // (`stalled`: it raised `done`, so everything behind it was a no-op too): then the host clears the flags, adds
  // launches with a look after each (the slow path) and re-enqueues from that step on.
  static thread_local std::unordered_map<const void*, int> kspec_hint;  // evaluations per step of the last run on ws
  int kspec = d->dtype == AMX_F32 ? 4 : 8;  // first guess of the evaluations per step
  {
    auto it = kspec_hint.find(ws);
    if (d->mhd && it != kspec_hint.end()) kspec = it->second;
  }
  long long issued = 0;
  const int batch = d->mhd ? 1 : 8;  // steps between two copies of the control block
  constexpr int kRing = 4;
  static thread_local Ctl* ring = nullptr;
  static thread_local cudaEvent_t ring_ev[kRing];
  if (!ring) {
    CK(cudaMallocHost(&ring, kRing * sizeof(Ctl)));
    for (int k = 0; k < kRing; ++k) CK(cudaEventCreateWithFlags(&ring_ev[k], cudaEventDisableTiming));
  }
  auto head = [&](int slot) { return reinterpret_cast<Ctl*>(reinterpret_cast<char*>(ring) + (size_t)slot * sizeof(Ctl)); };
  long long snaps = 0, looked = 0;  // copies enqueued / looked at
  Ctl* m = nullptr;
  bool ended = false;
  while (!ended) {
    StepPlan P;
    for (int k = 0; k < batch; ++k, ++issued) {
      ++E.launches;
      if (E.is32) step_begin_kernel<float><<<1, 1, 0, E.st>>>(E.ctl, p, (long long)max_steps);
      else step_begin_kernel<double><<<1, 1, 0, E.st>>>(E.ctl, p, (long long)max_steps);
      CK(cudaGetLastError());
      const View& sin = issued == 0 ? vin : S[(issued - 1) & 1];
      // shell mode: ctl->alpha now holds the bounds of the padded array (what dt needed); the
      // Lax-Friedrichs / PP stages want those of the ghost-FILLED state, i.e. of the interior
      // (also on the first step of a resumed segment, whose initial bounds include the carried shell)
      if (E.shell_on && E.lf) CK(E.alpha_of(sin, nullptr, true));
      if ((rc = step_front(E, sin, S[issued & 1], kspec, 1, P))) return rc;
      if (d->mhd) {
        ++E.launches;
        mag_guard_kernel<<<1, 1, 0, E.st>>>(E.ctl, 0);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(head((int)(snaps % kRing)), E.ctl, kCtlHead, cudaMemcpyDeviceToHost, E.st));
        CK(cudaEventRecord(ring_ev[snaps % kRing], E.st));
        ++snaps;
      }
      if ((rc = step_back(E, S[issued & 1], P))) return rc;
      ++E.launches;
      if (E.is32) step_end_kernel<float><<<1, 1, 0, E.st>>>(E.ctl);
      else step_end_kernel<double><<<1, 1, 0, E.st>>>(E.ctl);
      CK(cudaGetLastError());
    }
    if (!d->mhd) {
      CK(cudaMemcpyAsync(head((int)(snaps % kRing)), E.ctl, kCtlHead, cudaMemcpyDeviceToHost, E.st));
      CK(cudaEventRecord(ring_ev[snaps % kRing], E.st));
      ++snaps;
    }
    // look at every copy but the newest (whose step the device is still working on)
    while (looked + 1 < snaps && !ended) {
      CK(cudaEventSynchronize(ring_ev[looked % kRing]));
      const Ctl* c = head((int)(looked % kRing));
      ++looked;
      if (c->stalled) {
        // the fixed point of step c->steps needs more launches than it was given
        CK(cudaStreamSynchronize(E.st));
        if ((rc = read_ctl(E, &m))) return rc;
        const long long s_idx = m->steps;  // the stalled step (0-based) = steps completed so far
        ++E.launches;
        clear_stall_kernel<<<1, 1, 0, E.st>>>(E.ctl);
        CK(cudaGetLastError());
        const View& sin = s_idx == 0 ? vin : S[(s_idx - 1) & 1];
        StepPlan R;
        // the plan of the stalled step, without enqueuing anything: its buffers and where the fixed point stands
        R.A = E.internal(L.off_A); R.G = E.internal(L.off_G); R.Gf = E.internal(L.off_Gf);
        R.X[0] = E.internal(L.off_X[0]); R.X[1] = E.internal(L.off_X[1]);
        R.Bin = rows_of(sin, L.ng, L.esz); R.Bout = rows_of(S[s_idx & 1], L.ng, L.esz);
        R.pairs = E.pair_supported(E.mag_args(R.Bin, R.G, nullptr, R.X[0], E.dt_ptr(), 0, true));
        R.emit_ok = false;
        R.c = m->mag_iter;                // evaluations completed: every one that kept the loop going counted itself
        R.nl = (m->mag_last_buf ^ 1) | 2;  // next launch writes the other buffer; never "launch 0"
        R.expect = 1 << 30;
        if ((rc = fp_finish_blocking(E, R, &kspec))) return rc;
        if ((rc = step_back(E, S[s_idx & 1], R))) return rc;
        ++E.launches;
        if (E.is32) step_end_kernel<float><<<1, 1, 0, E.st>>>(E.ctl);
        else step_end_kernel<double><<<1, 1, 0, E.st>>>(E.ctl);
        CK(cudaGetLastError());
        issued = s_idx + 1;
        looked = snaps;  // the later copies belong to launches that did nothing
      } else if (c->done) {
        ended = true;
      } else if (d->mhd) {
        kspec = c->mag_iter + 1;
      }
    }
  }
  if ((rc = read_ctl(E, &m))) return rc;
  if (d->mhd) kspec_hint[ws] = kspec;
  // a step that found the loop finished wrote nothing: the state lives where step `steps` put it
  const long long steps = m->steps;
  const size_t bytes = (size_t)L.cells * L.esz;
  if (steps == 0) {
    for (int r = 0; r < L.nv; ++r)
      CK(cudaMemcpyAsync((char*)unpadded_out + r * bytes, (const char*)unpadded_in + r * bytes, bytes,
                         cudaMemcpyDeviceToDevice, E.st));
  } else {
    const char* src = E.ws + L.off_S[(steps - 1) & 1];
    for (int r = 0; r < L.nv; ++r)
      CK(cudaMemcpyAsync((char*)unpadded_out + r * bytes, src + r * L.row_bytes, bytes, cudaMemcpyDeviceToDevice, E.st));
  }
  CK(cudaStreamSynchronize(E.st));
  g_prof.resolve();
  if (stats) {
    stats->steps = steps;
    stats->magnetic_iterations = m->mag_iters_total;
    stats->time = m->time_d;
    stats->last_dt = m->dt_d;
    stats->kernel_launches = E.launches;
    stats->pair_launches = m->pair_runs;
    stats->single_launches = m->single_runs;
    stats->final_launches = m->final_runs;
  }
  if (m->not_converged) return fail(AMX_ENOTCONVERGED, "magnetic fixed point hit the 1000-iteration cap");
  return AMX_OK;
}

// ---- slab-decomposed building blocks ---------------------------------------------------------
int amx_alpha(const amx_desc* d, const void* gas, void* alpha_out, void* ws, void* stream) {
  int rc = check_common(d, gas, alpha_out);
  if (rc) return rc;
  if (!ws) return fail(AMX_EINVAL, "ws is null");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  CK(E.alpha_of(E.user_internal(gas), alpha_out, false));
  return AMX_OK;
}

int amx_dt_from_alpha(const amx_desc* d, const void* alpha_dev, const void* time_dev, double t_end,
                      void* dt_out, void* stream) {
  int rc = check_common(d, alpha_dev, dt_out);
  if (rc) return rc;
  if ((rc = amx_device_check())) return rc;
  Engine E(d, nullptr, stream);
  const DtPar p = E.dtpar(t_end);
  if (E.is32) dt_only_kernel<float><<<1, 1, 0, E.st>>>((const float*)alpha_dev, (const float*)time_dev, p, (float*)dt_out);
  else dt_only_kernel<double><<<1, 1, 0, E.st>>>((const double*)alpha_dev, (const double*)time_dev, p, (double*)dt_out);
  CK(cudaGetLastError());
  return AMX_OK;
}

int amx_hydro_stage(const amx_desc* d, const void* gas_in, const void* gas0_or_null, void* gas_out,
                    const void* dt_dev, double dt_scale, const void* alpha_lf, void* alpha_out, void* ws,
                    void* stream) {
  int rc = check_common(d, gas_in, gas_out);
  if (rc) return rc;
  if (!ws || !dt_dev) return fail(AMX_EINVAL, "ws / dt_dev is null");
  if (d->riemann == AMX_LAX_FRIEDRICHS && !alpha_lf) return fail(AMX_EINVAL, "Lax-Friedrichs needs alpha_lf");
  // reconstruction.py:186-193: the PP limiter weights C = alpha / sum(alpha) come from the stage's input
  if (d->limiter == 5 && d->split == AMX_UNSPLIT && !alpha_lf)
    return fail(AMX_EINVAL, "VAN_ALBADA_PP needs alpha_lf (wave bounds of the stage's input state)");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  const View in = E.user_internal(gas_in), out = E.user_internal(gas_out);
  View in0;
  if (gas0_or_null) in0 = E.user_internal(gas0_or_null);
  CK(E.stage(in, gas0_or_null ? &in0 : nullptr, out, dt_dev, dt_scale, alpha_lf, alpha_out != nullptr, alpha_out, false));
  return AMX_OK;
}

int amx_split_sweep(const amx_desc* d, const void* gas_in, void* gas_out, int axis, const void* dt_dev,
                    double dt_scale, const void* alpha_lf, void* alpha_out, void* ws, void* stream) {
  int rc = check_common(d, gas_in, gas_out);
  if (rc) return rc;
  if (!ws || !dt_dev) return fail(AMX_EINVAL, "ws / dt_dev is null");
  if (axis < 1 || axis > d->ndim) return fail(AMX_EINVAL, "axis out of range");
  if (d->riemann == AMX_LAX_FRIEDRICHS && !alpha_lf) return fail(AMX_EINVAL, "Lax-Friedrichs needs alpha_lf");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  CK(E.sweep(E.user_internal(gas_in), E.user_internal(gas_out), axis, dt_dev, dt_scale, alpha_lf,
             alpha_out != nullptr, alpha_out, false));
  return AMX_OK;
}

int amx_magnetic_iter(const amx_desc* d, const void* B0, const void* gas0, const void* bvk_or_null, void* bvk1,
                      const void* dt_dev, void* change_out, void* ws, void* stream) {
  int rc = check_common(d, B0, bvk1);
  if (rc) return rc;
  if (!ws || !dt_dev || !gas0 || !change_out) return fail(AMX_EINVAL, "null pointer");
  if (!d->mhd) return fail(AMX_EINVAL, "descriptor is not MHD");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  const View vB0 = E.user_internal(B0), vg = E.user_internal(gas0), vk1 = E.user_internal(bvk1);
  View vk;
  if (bvk_or_null) vk = E.user_internal(bvk_or_null);
  CK(E.mag_iter(vB0, vg, bvk_or_null ? &vk : nullptr, vk1, dt_dev, bvk_or_null ? 1 : 0, false));
  CK(E.finalize_change(change_out, false));
  return AMX_OK;
}

int amx_magnetic_final(const amx_desc* d, void* gas_inout, const void* bv_final, void* B_out, void* alpha_out,
                       void* ws, void* stream) {
  int rc = check_common(d, gas_inout, bv_final);
  if (rc) return rc;
  if (!ws || !B_out) return fail(AMX_EINVAL, "null pointer");
  if (!d->mhd) return fail(AMX_EINVAL, "descriptor is not MHD");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  const View X[2] = {E.user_internal(bv_final), E.user_internal(bv_final)};
  MagFinalArgs a;
  a.g = E.g; a.np = E.np; a.gas = E.user_internal(gas_inout); a.B_out = E.user_internal(B_out);
  a.gas_out = a.gas; a.skip_if_emitted = 0; a.final_runs = nullptr;
  a.X[0] = X[0]; a.X[1] = X[1]; a.ctl_ro = nullptr; a.fixed_final = 0;
  a.alpha_slots = alpha_out ? E.aslots() : nullptr;
  a.peers = E.peers;
  CK(E.strict ? amx_strict::launch_mag_final(a, E.st) : amx_fast::launch_mag_final(a, E.st));
  if (alpha_out) CK(E.finalize_alpha(alpha_out, false));
  return AMX_OK;
}

int amx_magnetic_iter_guarded(const amx_desc* d, const void* B0, const void* gas0, const void* bvk_or_null,
                              void* bvk1, const void* dt_dev, void* change_out, int launch_index, void* ws,
                              void* stream) {
  int rc = check_common(d, B0, bvk1);
  if (rc) return rc;
  if (!ws || !dt_dev || !gas0 || !change_out) return fail(AMX_EINVAL, "null pointer");
  if (!d->mhd) return fail(AMX_EINVAL, "descriptor is not MHD");
  if (launch_index < 0 || (launch_index == 0) != (bvk_or_null == nullptr))
    return fail(AMX_EINVAL, "launch 0 (and only launch 0) starts from R_0: bvk must be NULL exactly then");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  const View vB0 = E.user_internal(B0), vg = E.user_internal(gas0), vk1 = E.user_internal(bvk1);
  View vk;
  if (bvk_or_null) vk = E.user_internal(bvk_or_null);
  CK(E.mag_iter(vB0, vg, bvk_or_null ? &vk : nullptr, vk1, dt_dev, launch_index, true));
  CK(E.finalize_change(change_out, false));
  return AMX_OK;
}

int amx_magnetic_check(const amx_desc* d, const void* change_dev, int launch_index, void* ws, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (!ws || !change_dev || launch_index < 0) return fail(AMX_EINVAL, "null pointer / negative launch index");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  if (E.is32) mag_check_slab_kernel<float><<<1, 1, 0, E.st>>>(E.ctl, (const float*)change_dev, d->fp_tol, launch_index);
  else mag_check_slab_kernel<double><<<1, 1, 0, E.st>>>(E.ctl, (const double*)change_dev, d->fp_tol, launch_index);
  CK(cudaGetLastError());
  return AMX_OK;
}

int amx_magnetic_final_guarded(const amx_desc* d, const void* gas_in, void* gas_out, const void* bv0,
                               const void* bv1, void* B_out, void* alpha_out, void* ws, void* stream) {
  int rc = check_common(d, gas_in, bv0);
  if (rc) return rc;
  if (!ws || !B_out || !bv1 || !gas_out) return fail(AMX_EINVAL, "null pointer");
  if (!d->mhd) return fail(AMX_EINVAL, "descriptor is not MHD");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  MagFinalArgs a;
  a.g = E.g; a.np = E.np; a.gas = E.user_internal(gas_in); a.B_out = E.user_internal(B_out);
  a.gas_out = E.user_internal(gas_out); a.skip_if_emitted = 1; a.final_runs = nullptr;
  a.X[0] = E.user_internal(bv0); a.X[1] = E.user_internal(bv1); a.ctl_ro = E.ctl; a.fixed_final = -1;
  a.alpha_slots = alpha_out ? E.aslots() : nullptr;
  a.peers = E.peers;
  CK(E.strict ? amx_strict::launch_mag_final(a, E.st) : amx_fast::launch_mag_final(a, E.st));
  if (alpha_out) CK(E.finalize_alpha(alpha_out, false));
  return AMX_OK;
}

int amx_magnetic_pair_supported(const amx_desc* d) {
  if (validate(d) || !d->mhd || d->dtype != AMX_F32 || d->ndim != 3) return 0;
  Engine E(d, nullptr, nullptr);
  MagIterArgs a{};
  a.g = E.g; a.np = E.np;
  // layout rules are those of the internal slab layout: dense, even extents; probe with aligned dummies
  int pad[3];
  internal_pad(d, pad);
  const View v = make_view(reinterpret_cast<void*>(uintptr_t(256)), d, pad, 0, 0);
  a.B0 = a.gas0 = a.Bk = a.vk = a.Bk1 = a.vk1 = v;
  a.ctl_ro = nullptr; a.first = 0; a.midpoint = 1; a.dt = nullptr;
  a.change_slots = a.change_slots2 = reinterpret_cast<unsigned long long*>(uintptr_t(256));
  a.fix_base = -1;
  return E.pair_supported(a) ? 1 : 0;
}

int amx_magnetic_pair_guarded(const amx_desc* d, const void* B0, const void* gas0, const void* bvk_or_null,
                              void* bvk2, const void* dt_dev, void* change4_out, int evals_before,
                              void* emit_gas_or_null, void* emit_B_or_null, void* ws, void* stream) {
  int rc = check_common(d, B0, bvk2);
  if (rc) return rc;
  if (!ws || !dt_dev || !gas0 || !change4_out) return fail(AMX_EINVAL, "null pointer");
  if (!d->mhd) return fail(AMX_EINVAL, "descriptor is not MHD");
  if (evals_before < 0 || (evals_before == 0) != (bvk_or_null == nullptr))
    return fail(AMX_EINVAL, "the first pair (and only the first) starts from R_0: bvk must be NULL exactly then");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  const View vB0 = E.user_internal(B0), vg = E.user_internal(gas0), vk2 = E.user_internal(bvk2);
  View vk;
  if (bvk_or_null) vk = E.user_internal(bvk_or_null);
  if ((emit_gas_or_null == nullptr) != (emit_B_or_null == nullptr))
    return fail(AMX_EINVAL, "emit_gas and emit_B go together");
  View eg, eb;
  MagIterArgs probe = E.mag_args(vB0, vg, bvk_or_null ? &vk : nullptr, vk2, dt_dev, evals_before, true);
  if (emit_gas_or_null) {
    eg = E.user_internal(emit_gas_or_null); eb = E.user_internal(emit_B_or_null);
    probe.emit = 1; probe.emit_gas = eg; probe.emit_B = eb;
  }
  if (!E.pair_supported(probe))
    return fail(AMX_EUNSUPPORTED, "pair launches need 3-D fp32 with periodic y / z and halo or periodic x");
  CK(E.mag_pair(vB0, vg, bvk_or_null ? &vk : nullptr, vk2, dt_dev, evals_before, emit_gas_or_null ? &eg : nullptr,
                emit_gas_or_null ? &eb : nullptr));
  const size_t esz = E.L.esz;
  CK(E.finalize_change(change4_out, false));
  CK(E.finalize_change2(static_cast<char*>(change4_out) + 2 * esz));
  return AMX_OK;
}

int amx_magnetic_check_pair(const amx_desc* d, const void* change4_dev, int evals_before, int buffer_index,
                            int emit, void* ws, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (!ws || !change4_dev || evals_before < 0 || buffer_index < 0 || buffer_index > 1)
    return fail(AMX_EINVAL, "null pointer / bad index");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  if (E.is32) mag_check_pair_slab_kernel<float><<<1, 1, 0, E.st>>>(E.ctl, (const float*)change4_dev, d->fp_tol, evals_before, buffer_index, emit ? 1 : 0);
  else mag_check_pair_slab_kernel<double><<<1, 1, 0, E.st>>>(E.ctl, (const double*)change4_dev, d->fp_tol, evals_before, buffer_index, emit ? 1 : 0);
  CK(cudaGetLastError());
  return AMX_OK;
}

int amx_magnetic_fix_guarded(const amx_desc* d, const void* B0, const void* gas0, const void* bvk_or_null,
                             void* bvk1, const void* dt_dev, int evals_before, void* ws, void* stream) {
  int rc = check_common(d, B0, bvk1);
  if (rc) return rc;
  if (!ws || !dt_dev || !gas0) return fail(AMX_EINVAL, "null pointer");
  if (evals_before < 0 || (evals_before == 0) != (bvk_or_null == nullptr))
    return fail(AMX_EINVAL, "bvk must be NULL exactly when evals_before == 0");
  if (!d->mhd) return fail(AMX_EINVAL, "descriptor is not MHD");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  const View vB0 = E.user_internal(B0), vg = E.user_internal(gas0), vk1 = E.user_internal(bvk1);
  View vk;
  if (bvk_or_null) vk = E.user_internal(bvk_or_null);
  // the re-run goes to the vectorised single-evaluation kernel: same layout rules as the pair launch it repairs
  if (!E.pair_supported(E.mag_args(vB0, vg, bvk_or_null ? &vk : nullptr, vk1, dt_dev, evals_before, true)))
    return fail(AMX_EUNSUPPORTED, "pair launches (and their re-run) need 3-D fp32 with periodic y / z and halo or periodic x");
  CK(E.mag_fix(vB0, vg, bvk_or_null ? &vk : nullptr, vk1, dt_dev, evals_before));
  return AMX_OK;
}

int amx_magnetic_state(const amx_desc* d, void* ws, void* stream, int32_t* active_out, int32_t* iterations_out) {
  int rc = validate(d);
  if (rc) return rc;
  if (!ws || !active_out || !iterations_out) return fail(AMX_EINVAL, "null pointer");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  Ctl* m;
  if ((rc = read_ctl(E, &m))) return rc;
  *active_out = m->mag_active;
  *iterations_out = m->mag_iter;
  if (m->not_converged) return fail(AMX_ENOTCONVERGED, "magnetic fixed point hit the 1000-iteration cap");
  return AMX_OK;
}

int amx_shell_update(const amx_desc* d, int mode, void* shell_inout, const void* u0, const void* u1, const void* bv0,
                     const void* bv1, void* alpha_out, void* ws, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (!shell_inout || !ws) return fail(AMX_EINVAL, "null pointer");
  if (mode < 0 || mode > 3) return fail(AMX_EINVAL, "shell mode must be 0 .. 3");
  if ((mode == 0 && (!u0 || !u1)) || (mode == 1 && (!u0 || !bv0 || !bv1)) || (mode == 3 && !u1))
    return fail(AMX_EINVAL, "shell update: missing input array for this mode");
  if ((rc = amx_device_check())) return rc;
  Engine E(d, ws, stream);
  int pad[3] = {0, 0, 0};
  for (int a = 0; a < d->ndim; ++a) pad[a] = d->num_ghost;
  const View sh = make_view(shell_inout, d, pad, 0, 0);
  View v0 = sh, v1 = sh, X[2] = {sh, sh};
  if (u0) v0 = E.user_internal(u0);
  if (u1) v1 = E.user_internal(u1);
  if (bv0) X[0] = E.user_internal(bv0);
  if (bv1) X[1] = E.user_internal(bv1);
  CK(E.shell(mode, sh, sh, &v0, &v1, X, alpha_out != nullptr, false));
  if (alpha_out) CK(E.finalize_alpha(alpha_out, false));
  return AMX_OK;
}

int amx_peer_push_supported(const amx_desc* d) {
  if (validate(d)) return 0;
  if (d->bc[0][0] != AMX_BC_HALO && d->bc[0][1] != AMX_BC_HALO) return 0;
  Engine E(d, nullptr, nullptr);
  int pad[3];
  internal_pad(d, pad);
  const View v = make_view(reinterpret_cast<void*>(uintptr_t(256)), d, pad, 0, 0);
  int bits = 0;
  if (d->split == AMX_UNSPLIT && !force_generic() && !no_vec()) {
    StageArgs a{};
    a.g = E.g; a.np = E.np; a.in = a.in0 = a.out = v;
    a.alpha_lf = reinterpret_cast<const void*>(uintptr_t(256));
    if (E.strict ? amx_strict::vec_stage_supported(a) : amx_fast::vec_stage_supported(a)) bits |= 1;
  }
  if (d->mhd) {
    MagIterArgs a{};
    a.g = E.g; a.np = E.np;
    a.B0 = a.gas0 = a.Bk = a.vk = a.Bk1 = a.vk1 = v;
    a.change_slots = a.change_slots2 = reinterpret_cast<unsigned long long*>(uintptr_t(256));
    a.fix_base = -1;
    const bool vec = !force_generic() && !no_vec() &&
                     (E.strict ? amx_strict::vec_mag_supported(a) : amx_fast::vec_mag_supported(a));
    if (vec) bits |= 2;
    bits |= 4;
  }
  return bits;
}

size_t amx_peer_comm_bytes(void) { return kPeerCommBytes; }

int amx_peer_allreduce_max(void* const* comm_bufs, int rank, int world, void* vals_inout, int count, int dtype,
                           void* stream) {
  if (!comm_bufs || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(AMX_EINVAL, "peer all-reduce: 1 <= world <= 8, 0 <= rank < world");
  if (count < 0 || count > 8 || (count > 0 && !vals_inout)) return fail(AMX_EINVAL, "peer all-reduce: 0 <= count <= 8");
  if (dtype != AMX_F32 && dtype != AMX_F64) return fail(AMX_EINVAL, "dtype must be AMX_F32 or AMX_F64");
  int rc = amx_device_check();
  if (rc) return rc;
  PeerBufs b{};
  for (int r = 0; r < world; ++r) {
    if (!comm_bufs[r]) return fail(AMX_EINVAL, "peer all-reduce: null communication buffer");
    b.p[r] = reinterpret_cast<unsigned long long>(comm_bufs[r]);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == AMX_F32) peer_allreduce_max_kernel<float><<<1, 64, 0, st>>>(b, rank, world, static_cast<float*>(vals_inout), count);
  else peer_allreduce_max_kernel<double><<<1, 64, 0, st>>>(b, rank, world, static_cast<double*>(vals_inout), count);
  CK(cudaGetLastError());
  return AMX_OK;
}

}  // extern "C"
