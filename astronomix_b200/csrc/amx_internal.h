// amx_internal.h -- plain structs shared by the C-ABI layer (amx_api.cu) and the two
// arithmetic flavours of the kernel translation unit (amx_kernels.cu compiled with
// AMX_STRICT=0 -> namespace amx_fast, AMX_STRICT=1 -> namespace amx_strict).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amx {

constexpr int kAlphaSlots = 1024;  // spread of the per-block wave-bound atomics

// boundary codes (include/amx.h)
enum : int { BC_OPEN = 0, BC_REFLECTIVE = 1, BC_PERIODIC = 2, BC_JET = 3, BC_HALO = 4 };

// Strided view of a [rows][x][y][z] array; p points at row 0, FIRST ACTIVE cell.
struct View {
  void* p;
  long long s[3];  // element strides of x, y, z (absent axes: 0)
  long long vs;    // element stride between rows
};

struct Geom {
  int nd;
  int n[3];      // active cells per axis (absent axes: 1)
  int bc[3][2];  // boundary code per axis and side
  // MHD jet inflow (_geometry/boundaries.py:399-441; 2-D, y-left code BC_JET): the ghost cells of the
  // x-LEFT side whose y index lies in [jet_lo, jet_hi) (unpadded coordinates) hold constants; the
  // y-left side itself is an open side.  jet_hi <= jet_lo: no jet.
  int jet_lo, jet_hi;
  int jet_off;  // slab rank that does not own the x-left boundary: the window exists (same kernel family on every
                // rank) but no cell of this rank is an inflow cell
};

// the reference writes state[:, 0:ng, N/2-h : N/2+h] on the padded array, i.e. after every other fill
__host__ __device__ inline bool jet_cell(const Geom& g, int ix, int iy) {
  return g.jet_hi > g.jet_lo && !g.jet_off && ix < 0 && iy >= g.jet_lo && iy < g.jet_hi;
}
// boundaries.py:418-421: gas rows (rho, v_x, v_y, p) = (5/3, 800, 0, 1); velocity (800, 0, 0); B (sqrt(200), 0, 0)
template <typename T>
__host__ __device__ inline T jet_gas_value(int row) {
  return row == 0 ? (T)(5.0 / 3.0) : (row == 1 ? (T)800.0 : (row == 3 ? (T)1.0 : (T)0.0));
}
template <typename T>
__host__ __device__ inline T jet_vector_value(int type_handled, int comp) {  // 1 velocity-only, 2 magnetic-only
  if (comp != 0) return (T)0.0;
  return type_handled == 1 ? (T)800.0 : (T)14.142135623730951;  // 200 ** 0.5
}

// In-kernel halo stores of a slab rank (include/amx.h amx_desc::peer_*): a kernel that writes active plane
// x of an output array repeats the store into the neighbour's copy of the array, plane x + n of the lower
// neighbour (its right ghost planes) for x < 2 and plane x - n of the upper neighbour for x >= n - 2.
struct Peers {
  long long lo = 0, hi = 0;  // byte deltas from an own pointer to the same element of the neighbour's array
  int flags = 0;             // bit 0: lower neighbour, bit 1: upper neighbour
};
// element offset (in units of the x stride) and byte delta of the extra store for plane x, or false
__host__ __device__ inline bool peer_target(const Peers& P, int x, int n, int k, long long& byte_delta, int& plane_shift) {
  if (k == 0) {
    if (!(P.flags & 1) || x >= 2) return false;
    byte_delta = P.lo; plane_shift = n;
    return true;
  }
  if (!(P.flags & 2) || x < n - 2) return false;
  byte_delta = P.hi; plane_shift = -n;
  return true;
}

struct NumPar {
  int dtype;  // 0 f32, 1 f64
  int rs, lim;
  double dx, gamma;
  double pp_q;  // VAN_ALBADA_PP: 2 / C_cfl (MHD half steps) or 1 / C_cfl, evaluated in the working type
};

// Device-resident loop control.  Scalars are stored as double AND in the working type where
// the arithmetic must be done in the working type (time, dt); see amx_kernels.cu.
struct Ctl {
  double time_d, dt_d;       // mirrors for the host (exact conversions of the T values)
  float time_f, dt_f;        // working-type copies when dtype == f32
  double alpha[3];           // finalised wave bounds of the latest gas state (T values, widened)
  float alpha_f[3];          // the same as float (what fp32 Lax-Friedrichs stages read)
  float pad1;
  double change[2];          // finalised fixed-point change norms
  int done;                  // time loop ended (t >= t_end or step cap reached)
  int mag_active;            // fixed point still iterating
  int mag_iter;              // while-body iterations executed this step
  int mag_final;             // buffer index holding the converged iterate
  int not_converged;         // sticky: 1000-iteration cap was hit
  int mag_fix;               // the loop ended on the FIRST evaluation of pair launch `mag_fix_base`:
  int mag_fix_base;          //   that evaluation is re-run alone (evaluations completed before the pair)
  int final_emitted;         // the pair launch that ended the loop already wrote the write-back (amx_pair.cu EMIT)
  int mag_last_buf;          // iterate buffer the latest fixed-point launch wrote (while the loop is still active)
  int stalled;               // the launches enqueued for this step's fixed point did not suffice: `done` was raised
                             // with it, so everything enqueued behind exits at once; the host clears both, adds
                             // launches and resumes the step (amx_integrate's speculative pipeline)
  long long steps;
  long long mag_iters_total;
  long long pair_runs, single_runs;  // fixed-point launches that did work (bench.py byte accounting)
  long long final_runs;              // write-back launches that did work (not pre-empted by an emitting pair)
  unsigned long long alpha_slots[3][kAlphaSlots];  // bit patterns of non-negative T values
  unsigned long long change_slots[2][kAlphaSlots];
  unsigned long long change2_slots[2][kAlphaSlots];  // second evaluation of a pair launch (amx_pair.cu)
};

struct StageArgs {
  Geom g;
  NumPar np;
  View in;          // primitive gas state entering the stage
  View in0;         // p == nullptr: plain stage; else RK2 combine with this state (Q11)
  View out;
  const Ctl* ctl_ro;  // nullable: skip the launch's work when ctl->done
  const void* dt;     // device scalar of the working type
  double dt_scale;    // 1 or 0.5 (exact)
  const void* alpha_lf;  // device, 3 values of the working type (Lax-Friedrichs only)
  unsigned long long* alpha_slots;  // nullable: [3][kAlphaSlots] wave bounds of the result
  Peers peers;      // in-kernel halo stores of `out` (vec family only; see amx_peer_push_supported)
};

struct SweepArgs {
  Geom g;
  NumPar np;
  View in, out;
  const Ctl* ctl_ro;
  const void* dt;
  double dt_scale;
  int axis;  // 1-based sweep axis
  int first_order;
  const void* alpha_lf;
  unsigned long long* alpha_slots;
};

struct MagIterArgs {
  Geom g;
  NumPar np;
  View B0;    // 3 rows
  View gas0;  // ND+2 rows: rho (row 0) and v0 (rows 1..ND; v_z = 0 in 2-D)
  View Bk, vk;    // 3 rows each; Bk.p == nullptr: first iterate (R_k = R_0)
  View Bk1, vk1;  // 3 rows each
  const Ctl* ctl_ro;  // nullable: skip when done or !mag_active
  int first;          // launch 0 of the fixed point never skips on mag_active
  int midpoint;       // IMPLICIT_MIDPOINT (1) or IMPLICIT_EULER (0)
  const void* dt;
  unsigned long long* change_slots;   // [2][kAlphaSlots]; nullable (no norms wanted)
  unsigned long long* change_slots2;  // pair launch: norms of the second evaluation
  int fix_base;  // >= 0: run only if ctl->mag_fix && ctl->mag_fix_base == fix_base (see Ctl)
  // pair launch expected to end the loop: its second evaluation also writes what mag_final would
  // (_magnetic_field_update.py:208-240): gas rows (rho, v_n, p_2) to emit_gas, B_n to emit_B
  int emit;
  View emit_gas, emit_B;
  Peers peers;  // in-kernel halo stores of Bk1 / vk1 (and emit_gas / emit_B); vec / pair kernels only
};

struct MagFinalArgs {
  Geom g;
  NumPar np;
  View gas;       // in: rho, v_0, p
  View gas_out;   // out: rho, v_n, p_2 (may be the same array as gas)
  View B_out;     // 3 rows
  View X[2];      // the two iterate buffers, 6 rows each (B then v)
  const Ctl* ctl_ro;
  int fixed_final;  // >= 0: use this buffer index instead of ctl->mag_final
  int skip_if_emitted;  // do nothing when ctl->final_emitted (an emitting pair launch did the work)
  long long* final_runs;  // nullable: counts launches that did work
  unsigned long long* alpha_slots;
  Peers peers;  // in-kernel halo stores of gas_out / B_out
};

struct AlphaArgs {
  Geom g;
  NumPar np;
  View gas;
  const Ctl* ctl_ro;
  unsigned long long* alpha_slots;
};

// Ghost shell of the gas state (amx_kernels.cu shell_kernel): the reference reduces the CFL bound
// over the PADDED array (Q3), leaves the ghost cells of the RK2 combine un-refilled (Q11) and fills
// the ghost velocities in magnetic_update with the 3-vector flip rule (Q12).  With a reflective
// side in an unsplit MHD run the ghost gas cells follow their own recurrence and can set dt; the
// kernels work on unpadded arrays, so the shell is carried beside them.
struct ShellArgs {
  Geom g;
  NumPar np;
  int mode;        // 0 hydro half step, 1 magnetic write-back, 2 wave bounds of the shell as it is,
                   // 3 refill: shell <- ghost fill of u1 (what a split sweep's trailing _boundary_handler leaves)
  int ng;          // ghost cells per side (2)
  View shell_in;   // ND+2 gas rows of a padded array, p at the first ACTIVE cell
  View shell_out;  // updated shell (may be the same array)
  View u0;         // mode 0: interior gas entering the half step; mode 1: interior gas before the write-back
  View u1;         // mode 0: interior gas leaving the half step
  View X[2];       // mode 1: fixed-point iterate buffers (rows 3..5 = v_n)
  const Ctl* ctl_ro;
  int fixed_final;  // >= 0: iterate buffer index; < 0: ctl_ro->mag_final
  unsigned long long* alpha_slots;  // nullable: add the wave bounds of the (updated) shell cells
};

struct FillArgs {  // ghost fill of a padded array (GHOST_CELLS seam arrays)
  Geom g;
  int dtype;
  int ng;
  int rows;
  int type_handled;  // 0 gas, 1 velocity-only, 2 magnetic-only (Q12 flip rule)
  View a;            // p at first ACTIVE cell of the padded array
};

struct CopyArgs {  // strided row copy (pad / unpad / concatenate)
  Geom g;
  int dtype;
  int rows;
  View src, dst;
  int edge_pad;  // >0: dst is padded by this many cells, fill them edge-mode (np.pad 'edge')
};

// Opt a kernel into more than 48 KB of dynamic shared memory once per DEVICE: the attribute belongs
// to the device's context, and one process may drive several devices (JAX runs all local GPUs from
// one process and calls the FFI handlers per device).  `done` is a per-launcher static.
struct SmemOptIn {
  bool done[64] = {};
  template <typename K>
  cudaError_t ensure(K kern, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) done[dev] = true;
    return cudaSuccess;
  }
  bool first_time() const {  // true until the current device has been configured
    int dev = 0;
    return cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !done[dev];
  }
};

#define AMX_DECLARE_LAUNCHERS                                                             \
  cudaError_t launch_hydro_stage(const amx::StageArgs&, cudaStream_t);                    \
  cudaError_t launch_split_sweep(const amx::SweepArgs&, cudaStream_t);                    \
  cudaError_t launch_mag_iter(const amx::MagIterArgs&, cudaStream_t);                     \
  cudaError_t launch_mag_final(const amx::MagFinalArgs&, cudaStream_t);                   \
  cudaError_t launch_alpha(const amx::AlphaArgs&, cudaStream_t);                          \
  cudaError_t launch_shell(const amx::ShellArgs&, cudaStream_t);                          \
  bool tiled_stage_supported(const amx::StageArgs&);                                      \
  cudaError_t launch_hydro_stage_tiled(const amx::StageArgs&, cudaStream_t);             \
  bool tiled_mag_supported(const amx::MagIterArgs&);                                      \
  cudaError_t launch_mag_iter_tiled(const amx::MagIterArgs&, cudaStream_t);               \
  bool vec_mag_supported(const amx::MagIterArgs&);                                        \
  cudaError_t launch_mag_iter_vec(const amx::MagIterArgs&, cudaStream_t);                 \
  bool vec_stage_supported(const amx::StageArgs&);                                        \
  cudaError_t launch_hydro_stage_vec(const amx::StageArgs&, cudaStream_t);                \
  bool pair_mag_supported(const amx::MagIterArgs&);                                       \
  cudaError_t launch_mag_pair_vec(const amx::MagIterArgs&, cudaStream_t);

}  // namespace amx

namespace amx_fast { AMX_DECLARE_LAUNCHERS }
namespace amx_strict { AMX_DECLARE_LAUNCHERS }
