/* amx.h -- C ABI of libamx: the B200-native finite-volume (M)HD time step.
 *
 * This is the drop-in boundary.  The reference (leo1200/astronomix) has
 * no FFI seam of its own: its step is a set of pure jitted Python functions.  Each entry
 * point below names the reference function it replaces (paths relative to
 * /root/reference/astronomix/).  A jax.ffi shim over exactly these symbols is shown in
 * INTEGRATION.md; here the library is driven through ctypes (astronomix_b200/_lib.py).
 *
 * Conventions
 *  - All state pointers are DEVICE pointers to C-contiguous arrays [var][x][y][z]
 *    (z fastest; unused trailing axes absent) of desc->dtype.
 *  - "unpadded" arrays have desc->n[a] cells per axis.  "seam" arrays are what the
 *    reference's per-step functions see: padded by desc->num_ghost cells per side and
 *    axis when desc->boundary_handling == AMX_GHOST_CELLS, unpadded for AMX_PERIODIC_ROLL.
 *  - Nothing allocates, frees or synchronises unless stated; work is enqueued on `stream`
 *    (a cudaStream_t passed as void*).  `ws` is caller-owned scratch of
 *    amx_workspace_bytes() bytes, 256-byte aligned.  `in` is never written; `in` and
 *    `out` must not alias.
 *  - Return value 0 = ok; otherwise an AMX_E* code and amx_last_error() holds a message
 *    (thread-local).  No exceptions cross the ABI.  There is no CPU path: every compute
 *    entry point fails with AMX_ECUDA when no CUDA device is usable.
 */
#ifndef AMX_H_
#define AMX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMX_ABI_VERSION 3

#if defined(__GNUC__)
#define AMX_API __attribute__((visibility("default")))
#else
#define AMX_API
#endif

/* option codes = option_classes/simulation_config.py:20-93 */
enum { AMX_F32 = 0, AMX_F64 = 1 };
enum { AMX_MINMOD = 0, AMX_OSHER = 1, AMX_DOUBLE_MINMOD = 2, AMX_SUPERBEE = 3, AMX_VAN_ALBADA = 4 };
enum { AMX_UNSPLIT = 0, AMX_SPLIT = 1 };
enum { AMX_HLL = 0, AMX_HLLC = 1, AMX_HLLC_LM = 2, AMX_LAX_FRIEDRICHS = 3, AMX_HYBRID_HLLC = 4, AMX_AM_HLLC = 5 };
enum { AMX_RK2_SSP = 0, AMX_MUSCL = 1 };
/* AMX_BC_HALO: the ghost planes on that side exist in memory and were filled by the caller
 * (neighbour rank's cells in a slab decomposition); only valid on axis 0.
 * AMX_BC_MHD_JET (_geometry/boundaries.py:180-188, :399-441): only valid as bc[1][0] of a 2-D
 * AMX_GHOST_CELLS run -- the y-left side is an open copy and the x-left ghost rows hold the inflow
 * constants in a window of y; not with LAX_FRIEDRICHS / VAN_ALBADA_PP. */
enum { AMX_BC_OPEN = 0, AMX_BC_REFLECTIVE = 1, AMX_BC_PERIODIC = 2, AMX_BC_MHD_JET = 3, AMX_BC_HALO = 4 };
enum { AMX_GHOST_CELLS = 0, AMX_PERIODIC_ROLL = 1 };
enum { AMX_IMPLICIT_MIDPOINT = 0, AMX_IMPLICIT_EULER = 1 };
/* arithmetic: FAST = contracted FMAs, reciprocal multiplies, approximate sqrt/div (fp32);
 * STRICT = the reference's operation order with IEEE div/sqrt and no contraction. */
enum { AMX_ARITH_FAST = 0, AMX_ARITH_STRICT = 1 };

enum {
  AMX_OK = 0,
  AMX_EINVAL = 1,      /* malformed descriptor / null pointer */
  AMX_EUNSUPPORTED = 2,/* valid reference option outside this library's scope */
  AMX_ECUDA = 3,       /* CUDA runtime error (message has the cudaError string) */
  AMX_ENOTCONVERGED = 4/* magnetic fixed point hit the 1000-iteration cap */
};

/* POD mirror of the static options (SimulationConfig + SimulationParams + registry). */
typedef struct amx_desc {
  int32_t abi_version;        /* AMX_ABI_VERSION */
  int32_t ndim;               /* config.dimensionality: 1, 2, 3 */
  int32_t n[3];               /* active cells per axis (x, y, z); 1 for absent axes */
  int32_t dtype;              /* AMX_F32 / AMX_F64 (array dtype == JAX default dtype, Q13) */
  int32_t mhd;                /* config.mhd; state rows = ndim+2 (+3 magnetic rows last) */
  int32_t riemann;            /* config.riemann_solver */
  int32_t limiter;            /* config.limiter */
  int32_t split;              /* config.split */
  int32_t time_integrator;    /* config.time_integrator */
  int32_t boundary_handling;  /* config.boundary_handling */
  int32_t bc[3][2];           /* config.boundary_settings: [axis][left,right] */
  int32_t num_ghost;          /* config.num_ghost_cells (must be 2, Q8) */
  int32_t magnetic_integrator;/* config.fv_magnetic_integrator */
  int32_t first_order_fallback;/* config.first_order_fallback (split path only, Q7) */
  int32_t exact_end_time;     /* config.exact_end_time */
  int32_t use_max_adaptive_timestep; /* config.use_max_adaptive_timestep (split CFL only, Q3) */
  int32_t arithmetic;         /* AMX_ARITH_* */
  /* Slab ranks (AMX_BC_HALO on axis 0) whose arrays live in peer-mapped memory with the same layout on
   * every rank: the kernels that write an array also store its two first / last active x planes into the
   * right / left ghost planes of the lower / upper neighbour's copy of that array, at
   * (char*)own_pointer + peer_delta_{lo,hi} (P2P stores over NVLink, overlapped with the rest of the
   * launch), so that a halo exchange is only a barrier.  peer_flags: bit 0 = lower neighbour, bit 1 =
   * upper neighbour; 0 = off.  amx_peer_push_supported() tells which stage entry points honour it. */
  int32_t peer_flags;
  int32_t slab_flags;         /* bit 0: this slab rank does not own the x-left boundary (no MHD-jet inflow cells here) */
  int64_t peer_delta_lo, peer_delta_hi;
  int32_t reserved[2];
  double dx;                  /* config.grid_spacing */
  double gamma, C_cfl, dt_max;/* params.* (rounded to dtype on entry, Q20) */
  double fp_tol;              /* 1e-5 / 1e-10 from config.numerical_precision (Q6) */
} amx_desc;

/* step statistics, written by amx_integrate / amx_step into a host struct */
typedef struct amx_stats {
  int64_t steps;              /* time steps completed */
  int64_t magnetic_iterations;/* while-body iterations executed, summed over steps */
  double time;                /* simulation time reached */
  double last_dt;
  int64_t kernel_launches;    /* libamx kernels enqueued by the call */
  int64_t pair_launches;      /* fixed-point launches that made two evaluations (temporal blocking) */
  int64_t single_launches;    /* fixed-point launches that made one evaluation */
  int64_t final_launches;     /* write-back launches that did work (the others were pre-empted by a pair
                                 launch that wrote the write-back of its own result) */
} amx_stats;

AMX_API int amx_abi_version(void);
AMX_API const char* amx_last_error(void);
/* 0 when a CUDA device is usable from this process, else AMX_ECUDA (+ message) */
AMX_API int amx_device_check(void);

/* Per-kernel-class device timing for bench.py: when enabled, amx_integrate / amx_step bracket
 * every launch of a class with CUDA events on the launch stream and accumulate the elapsed
 * time.  Classes: 0 hydro stage, 1 split sweep, 2 magnetic iteration, 3 magnetic write-back,
 * 4 wave-bound reduction, 5 control, 6 magnetic pair (two evaluations per launch).  amx_profile_enable also resets the counters. */
AMX_API int amx_profile_enable(int on);
AMX_API int amx_profile_read(int kernel_class, double* ms_total, int64_t* launches);

/* scratch needed by every entry point below for this descriptor */
AMX_API int amx_workspace_bytes(const amx_desc* d, size_t* bytes_out);

/* time_stepping/_utils.py:65-75 + time_integration.py:274-286: edge-pad by 2 and fill the
 * gas rows' ghost cells (magnetic rows stay edge-padded, Q9).  unpadded -> seam array. */
AMX_API int amx_pad(const amx_desc* d, const void* unpadded_in, void* seam_out, void* stream);
/* time_stepping/_utils.py:19-61: seam array -> unpadded. */
AMX_API int amx_unpad(const amx_desc* d, const void* seam_in, void* unpadded_out, void* stream);
/* _geometry/boundaries.py:74-319 _boundary_handler(state, config, type_handled) on a seam
 * array with `nrows` rows, in place.  type_handled: 0 gas, 1 velocity-only, 2 B-only. */
AMX_API int amx_boundary_fill(const amx_desc* d, void* seam_inout, int nrows, int type_handled, void* stream);

/* _finite_volume/_timestep_estimation/_timestep_estimator.py:95-218 _cfl_time_step:
 * seam state -> dt (device scalar of desc->dtype at dt_out).  Like the reference, the reduction runs
 * over the array it is given: with AMX_GHOST_CELLS the ghost cells of the seam array take part
 * (SURVEY.md Q3), whatever they hold. */
AMX_API int amx_cfl(const amx_desc* d, const void* seam_state, void* dt_out, void* ws, void* stream);

/* _finite_volume/_state_evolution/evolve_state.py:497-600 _evolve_state_fv: one hydro step
 * (RK2-SSP unsplit :419-488 over :339-414, or Strang-split MUSCL sweeps :168-325 over
 * :53-163) or, for desc->mhd, hydro(dt/2) . magnetic_update(dt) . hydro(dt/2)
 * (_magnetic_update/_magnetic_field_update.py:37-242).  dt is read from device memory.
 * Ghost cells of seam_out: the fills the reference's trailing _boundary_handler calls leave.
 * One configuration differs, as it does in the reference: unsplit MHD with a reflective side.  There
 * the RK2 combine (evolve_state.py:456-469, no refill) mixes the INPUT's ghost gas cells with those of
 * the second stage, and magnetic_update fills ghost velocities with the 3-vector flip rule (Q12), so the
 * ghost gas cells of seam_out are a function of the ghost gas cells of seam_in ("ghost shell",
 * DESIGN.md 3.6) and the next amx_cfl on seam_out sees them. */
AMX_API int amx_step(const amx_desc* d, const void* seam_in, void* seam_out, const void* dt_dev,
             void* ws, void* stream, amx_stats* stats_or_null);

/* time_stepping/time_integration.py:61,239 time_integration (FORWARDS, no snapshots):
 * pad -> { dt = cfl; dt = min(dt, t_end - t); step; t += dt } while t < t_end -> unpad.
 * unpadded in -> unpadded out.  The loop runs on the device (dt, t and the magnetic
 * convergence flag live in `ws`); max_steps < 0 means no cap.  Blocks the calling host
 * thread until the loop has ended (it polls a pinned mirror of t), then fills *stats. */
AMX_API int amx_integrate(const amx_desc* d, const void* unpadded_in, void* unpadded_out, double t_end,
                  int64_t max_steps, void* ws, void* stream, amx_stats* stats_or_null);

/* The same loop resumed at simulation time t_start and also ending (without clipping dt) as soon as
 * time >= t_stop: the segments between two snapshots of time_integration.py:397-570, whose
 * snapshot condition `time >= checkpoint * t_end / num_snapshots` is tested at the start of a step.
 * stats->time is the absolute time reached.
 * Unsplit MHD with a reflective side keeps its ghost shell (see amx_step) in `ws`: a segment with
 * t_start != 0 continues from the shell the previous segment on the SAME ws left (checked by a stamp
 * of the problem shape; without one the shell restarts from images of the interior cells). */
AMX_API int amx_integrate_from(const amx_desc* d, const void* unpadded_in, void* unpadded_out, double t_start,
                  double t_stop, double t_end, int64_t max_steps, void* ws, void* stream,
                  amx_stats* stats_or_null);

/* ---- slab-decomposed building blocks (multi-GPU driver, astronomix_b200/slab.py) --------
 * The local slab is [rows][n[0] + 2*num_ghost][n[1]][n[2]] with desc->bc[0] = {HALO, HALO};
 * the caller exchanges the x ghost planes (NCCL send/recv) before each call that reads them. */

/* max over active cells of |u_axis| + c per axis -> alpha_out[3] (device, desc->dtype):
 * the per-rank part of _cfl_time_step (all-reduce MAX across ranks, then amx_dt_from_alpha). */
AMX_API int amx_alpha(const amx_desc* d, const void* gas, void* alpha_out, void* ws, void* stream);
/* dt = C/sum(alpha/dx) (unsplit) or min(C dx/max alpha, dt_max) (split); then, if
 * exact_end_time, dt = min(dt, t_end - *time_dev).  All scalars device-resident. */
AMX_API int amx_dt_from_alpha(const amx_desc* d, const void* alpha_dev, const void* time_dev, double t_end,
                      void* dt_out, void* stream);
/* evolve_state.py:339-414 one unsplit forward-Euler stage; with gas0 != NULL it also applies
 * the RK2 combine prim(0.5*(cons(gas0)+cons(stage))) (:456-469).  alpha_lf (device, 3 values,
 * global max |u|+c of gas_in per axis) is read by the Lax-Friedrichs flux only.  alpha_out
 * (nullable) receives the per-axis max |u|+c of the result. dt_scale is 1 or 0.5. */
AMX_API int amx_hydro_stage(const amx_desc* d, const void* gas_in, const void* gas0_or_null, void* gas_out,
                    const void* dt_dev, double dt_scale, const void* alpha_lf, void* alpha_out,
                    void* ws, void* stream);
/* evolve_state.py:53-163 one split MUSCL-Hancock sweep along `axis` (1-based). */
AMX_API int amx_split_sweep(const amx_desc* d, const void* gas_in, void* gas_out, int axis,
                    const void* dt_dev, double dt_scale, const void* alpha_lf, void* alpha_out,
                    void* ws, void* stream);
/* _magnetic_field_update.py:82-165 one evaluation R_{k+1} = R_0 - dt*Psi(arg(R_k, R_0)) of the
 * fixed point.  B0 is [3][...], gas0 the gas rows (rho and v_0 are read from it; v_z = 0 in
 * 2-D), bvk / bvk1 are [6][...] iterates (B then v); bvk == NULL is the first iterate
 * (R_k = R_0).  change_out (device, 2 values) receives max_cells |B_{k+1}-B_k|_2 and
 * max_cells |v_{k+1}-v_k|_2 of this rank's active cells. */
AMX_API int amx_magnetic_iter(const amx_desc* d, const void* B0, const void* gas0, const void* bvk_or_null,
                      void* bvk1, const void* dt_dev, void* change_out, void* ws, void* stream);
/* _magnetic_field_update.py:208-240: write the converged iterate back (v rows, pressure row of
 * gas_inout; B_out <- B_n); alpha_out (nullable) = wave bounds of the updated gas. */
AMX_API int amx_magnetic_final(const amx_desc* d, void* gas_inout, const void* bv_final, void* B_out,
                       void* alpha_out, void* ws, void* stream);

/* ---- device-resident fixed-point loop for slab ranks ------------------------------------
 * The same loop as amx_step's, cut at the points where ranks exchange data, so that the caller
 * can enqueue a speculative sequence without looking at the norms on the host:
 *   iter_guarded(0) . all-reduce MAX . check(0) | halo exchange . iter_guarded(1) . all-reduce . check(1) | ...
 * Launch `launch_index` >= 1 does nothing once the loop has ended; launch L writes iterate
 * buffer L & 1.  The loop state lives in the control block of `ws`. */
AMX_API int amx_magnetic_iter_guarded(const amx_desc* d, const void* B0, const void* gas0,
                      const void* bvk_or_null, void* bvk1, const void* dt_dev, void* change_out,
                      int launch_index, void* ws, void* stream);
/* _magnetic_field_update.py:127-141 while-condition on the ALL-REDUCED norms (device, 2 values of
 * desc->dtype) after launch `launch_index`; updates the loop state. */
AMX_API int amx_magnetic_check(const amx_desc* d, const void* change_dev, int launch_index, void* ws,
                      void* stream);
/* amx_magnetic_final reading the converged iterate from whichever of bv0 / bv1 the loop ended in and
 * writing the gas rows to gas_out (may be gas_in itself); a no-op when an emitting pair launch ended
 * the loop */
AMX_API int amx_magnetic_final_guarded(const amx_desc* d, const void* gas_in, void* gas_out, const void* bv0,
                      const void* bv1, void* B_out, void* alpha_out, void* ws, void* stream);
/* Two evaluations per launch for slab ranks (amx_pair.cu; 3-D fp32, periodic y / z): launch number L
 * reads the iterate in bvk (NULL: R_0, only when evals_before == 0) and writes R_{evals_before+2}
 * into bvk2; change4_out (device, 4 values) receives this rank's norms of both evaluations
 * (|dB|, |dv| of the first, then of the second).  The ghost planes of bvk / B0 / gas0 must hold
 * TWO valid planes per side.  amx_magnetic_pair_supported returns 1 when the descriptor qualifies.
 * emit_gas / emit_B (nullable, together): the launch is expected to end the loop and also writes the
 * write-back of its result (rho, v_n, p_2 -> emit_gas [ndim+2 rows]; B_n -> emit_B); pass emit = 1 to
 * amx_magnetic_check_pair for that launch, and amx_magnetic_final_guarded then does nothing if the
 * loop did end there. */
AMX_API int amx_magnetic_pair_supported(const amx_desc* d);
AMX_API int amx_magnetic_pair_guarded(const amx_desc* d, const void* B0, const void* gas0,
                      const void* bvk_or_null, void* bvk2, const void* dt_dev, void* change4_out,
                      int evals_before, void* emit_gas_or_null, void* emit_B_or_null, void* ws,
                      void* stream);
/* while-condition after EACH of the two evaluations, on the all-reduced norms; `buffer_index` is the
 * iterate buffer (0 / 1) the pair wrote.  If the loop ends on the first evaluation, the control
 * block asks for that evaluation alone: amx_magnetic_fix_guarded (same arguments as the pair) runs
 * only then and overwrites bvk2 with R_{evals_before+1}. */
AMX_API int amx_magnetic_check_pair(const amx_desc* d, const void* change4_dev, int evals_before,
                      int buffer_index, int emit, void* ws, void* stream);
AMX_API int amx_magnetic_fix_guarded(const amx_desc* d, const void* B0, const void* gas0,
                      const void* bvk_or_null, void* bvk1, const void* dt_dev, int evals_before,
                      void* ws, void* stream);
/* host look at the loop state (synchronises `stream`): still iterating?, while-body iterations so far */
AMX_API int amx_magnetic_state(const amx_desc* d, void* ws, void* stream, int32_t* active_out,
                      int32_t* iterations_out);

/* Ghost shell of the gas state on a slab rank: in unsplit MHD with a reflective side the
 * reference's ghost gas cells follow their own recurrence (evolve_state.py:456-469 without a ghost fill,
 * _magnetic_field_update.py:208-240 with the 3-vector flip rule) and enter the CFL reduction
 * (_timestep_estimator.py:118-128 runs over the padded array).  shell_inout: this rank's padded gas array
 * [ndim+2][n0+4][n1+4][n2+4] (only its ghost cells are used; planes beyond an AMX_BC_HALO side are skipped:
 * they belong to the neighbour).  mode 0: after a hydro half step u0 -> u1 (internal layout); mode 1: after
 * the fixed point (u0 = gas before the write-back, bv0 / bv1 = iterate buffers, the final one is taken from
 * the control block in ws); mode 2: wave bounds only; mode 3: shell <- ghost fill of u1 (initial state, split
 * sweeps).  alpha_out (nullable, 3 values of dtype): max |u_axis| + c over the shell cells. */
AMX_API int amx_shell_update(const amx_desc* d, int mode, void* shell_inout, const void* u0, const void* u1,
                      const void* bv0, const void* bv1, void* alpha_out, void* ws, void* stream);

/* ---- peer-memory plumbing of a slab run (one process per GPU, arrays in peer-mapped memory) ----
 * Which stage entry points honour amx_desc::peer_* for this descriptor (the kernels that would run):
 * bit 0 amx_hydro_stage, bit 1 amx_magnetic_iter* / _pair_guarded / _fix_guarded, bit 2
 * amx_magnetic_final*.  A caller still exchanges the halos of arrays written by the others itself. */
AMX_API int amx_peer_push_supported(const amx_desc* d);
/* MAX all-reduce of `count` <= 8 values of `dtype` (in place, device memory) over the ranks of the run
 * through peer memory: replaces the NCCL all_reduce of the wave bounds (_timestep_estimator.py:118-128 on
 * a decomposed grid) and of the fixed-point change norms (_magnetic_field_update.py:127-141); count = 0
 * is a barrier (it closes a halo exchange made of in-kernel stores).  One launch, no host round trip.
 * `comm_bufs[r]` = this process's mapping of rank r's communication buffer (amx_peer_comm_bytes() bytes,
 * zeroed once before the first call, same buffer for every call); every rank must make the same
 * sequence of calls.  max is order-independent: results are bitwise equal on all ranks. */
AMX_API size_t amx_peer_comm_bytes(void);
AMX_API int amx_peer_allreduce_max(void* const* comm_bufs, int rank, int world, void* vals_inout, int count,
                      int dtype, void* stream);

/* ---- jax's default PRNG on the device (turbulent driving) ----
 * jax.random.normal(key, shape) for a raw Threefry key (key0, key1) = jax.random.key_data(key): `count` =
 * prod(shape) samples of `dtype` written to `out` (device) in C order.  partitionable = 1: jax >= 0.5's
 * default counter layout (jax_threefry_partitionable), 0: the older default.  Replaces the two draws of
 * _physics_modules/_turbulent_forcing/_turbulent_forcing.py:50-51; key handling (key(42), split) is host-side
 * integer work (astronomix_b200/threefry.py). */
AMX_API int amx_threefry_normal(uint32_t key0, uint32_t key1, int64_t count, void* out, int dtype,
                      int partitionable, void* stream);
/* samples [offset, offset + count) of the same draw of `total` samples: what a slab rank needs of a global
 * (3, n, n, n) draw */
AMX_API int amx_threefry_normal_range(uint32_t key0, uint32_t key1, int64_t total, int64_t offset, int64_t count,
                      void* out, int dtype, int partitionable, void* stream);

/* _turbulent_forcing.py:15-65 up to the inverse FFT, fused: the rank's k_x slab [x0, x0 + nl) x n x n of the three
 * complex coefficient arrays (noise drawn in the kernel from the two Threefry keys of the step, spectrum
 * k^6 exp(-8 k / k_peak), DC mode zero, compressive part projected out) -> out_complex [3][nl][n][n] interleaved
 * (re, im) of `dtype`.  x0 = 0, nl = n: the whole grid. */
AMX_API int amx_forcing_spectrum(uint32_t re_key0, uint32_t re_key1, uint32_t im_key0, uint32_t im_key1, int n,
                      int x0, int nl, double box_size, void* out_complex, int dtype, int partitionable,
                      void* stream);
/* _turbulent_forcing.py:103-104: sums2_out[0] = 0.5 sum rho |w|^2, sums2_out[1] = sum rho v.w over `cells` cells
 * (double, deterministic two-pass reduction; w3 = [3][cells]; scratch >= 2048 doubles). */
AMX_API int amx_forcing_sums(const void* rho, const void* vx, const void* vy, const void* vz, const void* w3,
                      int64_t cells, void* scratch, void* sums2_out, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMX_H_ */
