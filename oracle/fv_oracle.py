"""CPU oracle of astronomix's finite-volume (M)HD time step -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference's algorithm for the hot path of
SURVEY.md §8.  It exists to *check* the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``astronomix_b200/`` imports it and the product path
raises when the CUDA library is missing; there is no CPU fallback.

How it is pinned.  (1) Against outputs of the reference itself run here: the
unmodified sources under /root/reference/astronomix are executed on NumPy through
``oracle/ref_on_numpy.py`` + ``oracle/jaxshim`` (jax, equinox and astropy cannot be installed),
``oracle/gen_reference_vectors.py`` / ``gen_reference_aux_vectors.py`` record 138 input/output
vectors of the reference's public ``time_integration`` (every solver, limiter, dimension,
boundary kind, both schemes, both precisions, MHD options, snapshots, driven turbulence), and
``tests/test_reference_vectors.py`` requires this restatement to reproduce each of them
(bit for bit in practice; bound 1e-13 fp64 / 2e-6 fp32).  (2) Against the one numerical fixture
the reference ships for this path, ``tests/mhd_tests/final_state_orszag_tang.npz``: run in that
file's configuration (SPLIT + MUSCL-Hancock, HLL, minmod, periodic ghost cells, implicit midpoint
at tol 1e-10, fp64, 3001 steps) it reproduces the file to ~1e-13 relative L1
(``oracle/pin_orszag_tang.py`` -> ``tests/golden/pin_orszag_tang.json``, re-checked from a late
checkpoint by ``tests/test_oracle_pin.py``).  Not certified by either: XLA's own instruction
selection (FMA contraction), a round-off-level effect the parity tolerances allow for.

Conventions.  Arrays are ``[var, x(, y(, z))]``.  ``roll(a, k, ax)[i] = a[i-k]``.
Scalar typing follows the reference under ``jax.jit``: SimulationParams
fields (gamma, C_cfl, dt_max, t_end) and ``dt``/``time`` are scalars of the working dtype,
SimulationConfig fields (grid_spacing) are Python doubles that round on first contact
with an array.  Every function cites the reference lines it follows (paths relative to
/root/reference/astronomix/).
"""
from __future__ import annotations

import numpy as np

# option codes: option_classes/simulation_config.py:20-93
MINMOD, OSHER, DOUBLE_MINMOD, SUPERBEE, VAN_ALBADA, VAN_ALBADA_PP = 0, 1, 2, 3, 4, 5
UNSPLIT, SPLIT = 0, 1
HLL, HLLC, HLLC_LM, LAX_FRIEDRICHS, HYBRID_HLLC, AM_HLLC = 0, 1, 2, 3, 4, 5
RK2_SSP, MUSCL = 0, 1
OPEN_BOUNDARY, REFLECTIVE_BOUNDARY, PERIODIC_BOUNDARY, MHD_JET_BOUNDARY = 0, 1, 2, 3
GAS_STATE, VELOCITY_ONLY, MAGNETIC_FIELD_ONLY = 0, 1, 2
GHOST_CELLS, PERIODIC_ROLL = 0, 1
IMPLICIT_MIDPOINT, IMPLICIT_EULER = 0, 1
SINGLE_PRECISION, DOUBLE_PRECISION = 0, 1


class Opts:
    """The static options the path reads (subset of SimulationConfig + SimulationParams)."""

    def __init__(self, dim, mhd=False, riemann_solver=HLL, limiter=MINMOD, split=UNSPLIT,
                 time_integrator=RK2_SSP, boundary_handling=GHOST_CELLS, boundaries=None,
                 grid_spacing=1.0, gamma=5 / 3, C_cfl=0.4, dt_max=1e-3, t_end=0.2,
                 fv_magnetic_integrator=IMPLICIT_MIDPOINT, numerical_precision=SINGLE_PRECISION,
                 exact_end_time=True, use_max_adaptive_timestep=True, first_order_fallback=False,
                 num_ghost_cells=2):
        self.dim = dim
        self.mhd = mhd
        self.riemann_solver = riemann_solver
        self.limiter = limiter
        self.split = split
        self.time_integrator = time_integrator
        self.boundary_handling = boundary_handling
        # [(left, right)] per axis
        self.boundaries = boundaries if boundaries is not None else [(OPEN_BOUNDARY, OPEN_BOUNDARY)] * dim
        self.grid_spacing = float(grid_spacing)
        self.gamma = gamma
        self.C_cfl = C_cfl
        self.dt_max = dt_max
        self.t_end = t_end
        self.fv_magnetic_integrator = fv_magnetic_integrator
        self.numerical_precision = numerical_precision
        self.exact_end_time = exact_end_time
        self.use_max_adaptive_timestep = use_max_adaptive_timestep
        self.first_order_fallback = first_order_fallback
        self.num_ghost_cells = num_ghost_cells
        self.rho_i = 0
        self.p_i = dim + 1  # registered_variables.py:131-170 (gas rows)

    @classmethod
    def from_config(cls, config, params):
        from astronomix_b200.option_classes import boundary_codes
        return cls(dim=config.dimensionality, mhd=config.mhd, riemann_solver=config.riemann_solver,
                   limiter=config.limiter, split=config.split, time_integrator=config.time_integrator,
                   boundary_handling=config.boundary_handling, boundaries=boundary_codes(config),
                   grid_spacing=config.grid_spacing, gamma=params.gamma, C_cfl=params.C_cfl,
                   dt_max=params.dt_max, t_end=params.t_end,
                   fv_magnetic_integrator=config.fv_magnetic_integrator,
                   numerical_precision=config.numerical_precision, exact_end_time=config.exact_end_time,
                   use_max_adaptive_timestep=config.use_max_adaptive_timestep,
                   first_order_fallback=config.first_order_fallback,
                   num_ghost_cells=config.num_ghost_cells)


# ------------------------------------------------------------------ stencil primitives
def roll(a, shift, axis):
    """_stencil_operations/_stencil_operations.py:20-30 (custom_roll == jnp.roll)."""
    return np.roll(a, shift, axis=axis)


def stencil_add(a, indices, factors, axis):
    """_stencil_operations.py:33-58: sum_j f_j * a[i + idx_j], evaluated as 0 + f0*a0 + f1*a1."""
    out = 0
    for f, idx in zip(factors, indices):
        out = out + a.dtype.type(f) * roll(a, -idx, axis)
    return out


# ------------------------------------------------------------------ fluid equations
def speed_of_sound(rho, p, gamma):
    """_fluid_equations/_equations.py:269-280."""
    return np.sqrt(gamma * p / rho)


def get_absolute_velocity(s, o):
    """_equations.py:165-199 (note: 1e-20 only in 3-D, Q2)."""
    if o.dim == 1:
        return np.abs(s[1])
    if o.dim == 2:
        return np.sqrt(s[1] ** 2 + s[2] ** 2)
    return np.sqrt(s[1] ** 2 + s[2] ** 2 + s[3] ** 2 + s.dtype.type(1e-20))


def total_energy_from_primitives(rho, u, p, gamma):
    """_equations.py:253-266."""
    return p / (gamma - 1) + rho.dtype.type(0.5) * rho * u ** 2


def pressure_from_energy(E, rho, u, gamma):
    """_equations.py:202-250: (gamma-1) * rho * (E/rho - 0.5 u^2)."""
    e = E / rho - rho.dtype.type(0.5) * u ** 2
    return (gamma - 1) * rho * e


def conserved_state_from_primitive(s, gamma, o):
    """_equations.py:104-156."""
    rho = s[o.rho_i]
    u = get_absolute_velocity(s, o)
    E = total_energy_from_primitives(rho, u, s[o.p_i], gamma)
    U = s.copy()
    U[o.p_i] = E
    for k in range(1, o.dim + 1):
        U[k] = rho * s[k]
    return U


def primitive_state_from_conserved(U, gamma, o):
    """_equations.py:24-94 (1e-20 in 2-D and 3-D, Q2)."""
    rho = U[o.rho_i]
    E = U[o.p_i]
    t = U.dtype.type
    if o.dim == 1:
        v = [U[1] / rho]
        u = v[0]
    elif o.dim == 2:
        v = [U[1] / rho, U[2] / rho]
        u = np.sqrt(v[0] ** 2 + v[1] ** 2 + t(1e-20))
    else:
        v = [U[1] / rho, U[2] / rho, U[3] / rho]
        u = np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2 + t(1e-20))
    s = U.copy()
    s[o.p_i] = pressure_from_energy(E, rho, u, gamma)
    for k in range(o.dim):
        s[k + 1] = v[k]
    return s


def euler_flux(s, gamma, o, axis):
    """_fluid_equations/_fluxes.py:25-93."""
    rho = s[o.rho_i]
    p = s[o.p_i]
    utotal = get_absolute_velocity(s, o)
    E = total_energy_from_primitives(rho, utotal, p, gamma)
    F = s.copy()
    F[o.p_i] = F[o.p_i] + E
    for k in range(1, o.dim + 1):
        F[k] = s[k] * rho
    F = s[axis] * F
    F[axis] = F[axis] + p
    return F


# ------------------------------------------------------------------ limiters
def _minmod(a, b):
    """_finite_volume/_state_evolution/limiters.py:14-15."""
    return a.dtype.type(0.5) * (np.sign(a) + np.sign(b)) * np.minimum(np.abs(a), np.abs(b))


def _minmod3(a, b, c):
    """limiters.py:18-25."""
    same = (np.sign(a) == np.sign(b)) & (np.sign(b) == np.sign(c))
    return np.where(same, np.sign(a) * np.minimum(np.abs(a), np.minimum(np.abs(b), np.abs(c))),
                    a.dtype.type(0.0))


def _double_minmod(a, b):
    """limiters.py:28-32."""
    return np.where(a * b > 0, _minmod3((a + b) / 2, 2 * a, 2 * b), a.dtype.type(0.0))


def _maxmod(a, b):
    """limiters.py:35-37."""
    return a.dtype.type(0.5) * (np.sign(a) + np.sign(b)) * np.maximum(np.abs(a), np.abs(b))


def _superbee(a, b):
    """limiters.py:40-44."""
    return np.where(a * b > 0, _minmod(_maxmod(a, b), _minmod(2 * a, 2 * b)), a.dtype.type(0.0))


def _van_albada(bd, fd, o):
    """limiters.py:48-64 (epsilon = 3*dx, dimensional, Q10)."""
    eps = 3 * o.grid_spacing
    return ((fd ** 2 + eps) * bd + (bd ** 2 + eps) * fd) / (fd ** 2 + bd ** 2 + 2 * eps)


def calculate_limited_gradients(s, o, axis):
    """limited_gradients.py:22-87 (Cartesian branch)."""
    dx = o.grid_spacing
    bd = stencil_add(s, (0, -1), (1.0, -1.0), axis) / dx
    fd = stencil_add(s, (1, 0), (1.0, -1.0), axis) / dx
    if o.limiter == MINMOD:
        return _minmod(bd, fd)
    if o.limiter == SUPERBEE:
        return _superbee(bd, fd)
    if o.limiter == DOUBLE_MINMOD:
        return _double_minmod(bd, fd)
    if o.limiter in (VAN_ALBADA, VAN_ALBADA_PP):
        return _van_albada(bd, fd, o)
    if o.limiter == OSHER:
        eps = 1e-11
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(np.abs(bd) > eps, fd / (bd + eps), np.zeros_like(bd))
        return np.maximum(0, np.minimum(1.3, g)) * bd
    raise ValueError("Unknown limiter.")


# ------------------------------------------------------------------ reconstruction
def reconstruct_unsplit_single(s, o, axis):
    """reconstruction.py:399-431."""
    g = calculate_limited_gradients(s, o, axis)
    d = g * o.grid_spacing / 2
    left_center = s - d
    right_center = s + d
    return roll(right_center, 1, axis), left_center


def reconstruct_split(s, dt, gamma, o, axis):
    """reconstruction.py:40-129 (MUSCL-Hancock predictor with the literal-1 quirk, Q1)."""
    rho = s[o.rho_i]
    p = s[o.p_i]
    u = s[axis]
    g = calculate_limited_gradients(s, o, axis)
    c = speed_of_sound(rho, p, gamma)
    # projected = einsum('ba..., a... -> b...', A_W, g) with
    # A_W = u*I ; A_W[rho, axis] = rho ; A_W[p, 1] = rho c^2 ; A_W[axis, p] = 1/rho   (:66-83)
    nv = s.shape[0]
    A = {}
    for b in range(nv):
        A[(b, b)] = u
    A[(o.rho_i, axis)] = rho
    A[(o.p_i, 1)] = rho * c ** 2
    A[(axis, o.p_i)] = 1 / rho
    proj = np.zeros_like(s)
    for (b, a), coeff in sorted(A.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        proj[b] = proj[b] + coeff * g[a]
    pred = s - dt / 2 * proj
    half = o.grid_spacing / 2
    prim_left = pred - half * g
    prim_right = pred + half * g
    return roll(prim_right, 1, axis), prim_left


# ------------------------------------------------------------------ Riemann solvers
def hll_solver(L, R, gamma, o, axis):
    """_riemann_solver/hll.py:36-109."""
    rho_L, u_L, p_L = L[o.rho_i], L[axis], L[o.p_i]
    rho_R, u_R, p_R = R[o.rho_i], R[axis], R[o.p_i]
    c_L = speed_of_sound(rho_L, p_L, gamma)
    c_R = speed_of_sound(rho_R, p_R, gamma)
    F_L = euler_flux(L, gamma, o, axis)
    F_R = euler_flux(R, gamma, o, axis)
    S_R = np.maximum(np.maximum(u_L + c_L, u_R + c_R), 0)
    S_L = np.minimum(np.minimum(u_L - c_L, u_R - c_R), 0)
    U_L = conserved_state_from_primitive(L, gamma, o)
    U_R = conserved_state_from_primitive(R, gamma, o)
    return (S_R * F_L - S_L * F_R + S_L * S_R * (U_R - U_L)) / (S_R - S_L)


def hllc_solver(L, R, gamma, o, axis, hllc_lm=False, low_mach_dissipation_control=False):
    """hll.py:123-261 (HLLC, HLLC-LM; low-Mach dissipation control for AM-HLLC)."""
    t = L.dtype.type
    rho_L, u_L, p_L = L[o.rho_i], L[axis], L[o.p_i]
    rho_R, u_R, p_R = R[o.rho_i], R[axis], R[o.p_i]
    c_L = speed_of_sound(rho_L, p_L, gamma)
    c_R = speed_of_sound(rho_R, p_R, gamma)
    F_L = euler_flux(L, gamma, o, axis)
    F_R = euler_flux(R, gamma, o, axis)
    sL, sR = np.sqrt(rho_L), np.sqrt(rho_R)
    u_hat = (sL * u_L + sR * u_R) / (sL + sR)
    c_hat_sq = (c_L ** 2 * sL + c_R ** 2 * sR) / (sL + sR) + t(0.5) * (sL * sR / (sL + sR) ** 2) * (u_R - u_L) ** 2
    c_hat = np.sqrt(c_hat_sq)
    S_L = np.minimum(u_L - c_L, u_hat - c_hat)
    S_R = np.maximum(u_R + c_R, u_hat + c_hat)
    S_star = (p_R - p_L + rho_L * u_L * (S_L - u_L) - rho_R * u_R * (S_R - u_R)) / (
        rho_L * (S_L - u_L) - rho_R * (S_R - u_R))
    U_L = conserved_state_from_primitive(L, gamma, o)
    U_R = conserved_state_from_primitive(R, gamma, o)
    Us_L = U_L.copy()
    Us_L[axis] = rho_L * S_star
    Us_L[o.p_i] = Us_L[o.p_i] + (S_star - u_L) * (rho_L * S_star + p_L / (S_L - u_L))
    Us_L = Us_L * (S_L - u_L) / (S_L - S_star)
    Us_R = U_R.copy()
    Us_R[axis] = rho_R * S_star
    Us_R[o.p_i] = Us_R[o.p_i] + (S_star - u_R) * (rho_R * S_star + p_R / (S_R - u_R))
    Us_R = Us_R * (S_R - u_R) / (S_R - S_star)
    if o.riemann_solver == HLLC_LM or hllc_lm:
        Ma_local = np.maximum(np.abs(u_L / c_L), np.abs(u_R / c_R))
        phi = np.sin(np.minimum(1, Ma_local / 0.1) * np.pi / 2)
        S_Ls, S_Rs = S_L * phi, S_R * phi
    else:
        S_Ls, S_Rs = S_L, S_R
    bulk = t(0.5) * (F_L + F_R)
    diss = t(0.5) * (S_Ls * (Us_L - U_L) + np.abs(S_star) * (Us_L - Us_R) + S_Rs * (Us_R - U_R))
    if low_mach_dissipation_control:
        aL = get_absolute_velocity(L, o)
        aR = get_absolute_velocity(R, o)
        f = np.minimum(1, np.maximum(np.abs(aL / c_L), np.abs(aR / c_R)))
        diss[1:1 + o.dim] = f * diss[1:1 + o.dim]
    F_star = bulk + diss
    F = np.where(S_L >= 0, F_L, F_star)
    F = np.where(S_R <= 0, F_R, F)
    return F


def am_hllc_solver(L, R, s, gamma, o, axis):
    """hll.py:268-326 (AM-HLLC / HYBRID-HLLC blend on a compression sensor)."""
    d = o.grid_spacing
    a = speed_of_sound(s[o.rho_i], s[o.p_i], gamma)
    div_v = sum(stencil_add(s[i + 1], (1, -1), (1.0, -1.0), i) / (2 * d) for i in range(o.dim))
    g = np.where(div_v < -0.05 * a / d, 1, 0).astype(s.dtype)
    lm = o.riemann_solver == AM_HLLC
    F_hllc = hllc_solver(L, R, gamma, o, axis, low_mach_dissipation_control=lm)
    F_lm = hllc_solver(L, R, gamma, o, axis, hllc_lm=True, low_mach_dissipation_control=lm)
    return g * F_lm + (1 - g) * F_hllc


def lax_friedrichs_solver(L, R, s, gamma, o, axis):
    """_riemann_solver/_lax_friedrichs.py:21-78 (one global alpha per axis, Q16)."""
    U_L = conserved_state_from_primitive(L, gamma, o)
    U_R = conserved_state_from_primitive(R, gamma, o)
    c = speed_of_sound(s[o.rho_i], s[o.p_i], gamma)
    alpha = np.max(np.abs(s[axis]) + c)
    F_L = euler_flux(L, gamma, o, axis)
    F_R = euler_flux(R, gamma, o, axis)
    t = s.dtype.type
    return t(0.5) * (F_L + F_R) - t(0.5) * alpha * (U_R - U_L)


def riemann_solver(L, R, s, gamma, o, axis):
    """_riemann_solver/_riemann_solver.py:37-87."""
    rs = o.riemann_solver
    if rs == HLL:
        return hll_solver(L, R, gamma, o, axis)
    if rs in (HLLC, HLLC_LM):
        return hllc_solver(L, R, gamma, o, axis)
    if rs in (AM_HLLC, HYBRID_HLLC):
        return am_hllc_solver(L, R, s, gamma, o, axis)
    if rs == LAX_FRIEDRICHS:
        return lax_friedrichs_solver(L, R, s, gamma, o, axis)
    raise ValueError("Riemann solver not supported.")


# ------------------------------------------------------------------ boundaries
def _set_along_axis(s, axis, set_index, get_index):
    """_geometry/boundaries.py:22-40."""
    sl_set = [slice(None)] * s.ndim
    sl_get = [slice(None)] * s.ndim
    sl_set[axis] = set_index
    sl_get[axis] = get_index
    s[tuple(sl_set)] = s[tuple(sl_get)]


def _set_specific_var_along_axis(s, axis, set_index, get_index, var_index, factor):
    """boundaries.py:43-70; an out-of-range var_index is silently dropped as under jax (Q12)."""
    if var_index >= s.shape[0]:
        return
    sl_set = [slice(None)] * s.ndim
    sl_get = [slice(None)] * s.ndim
    sl_set[0] = var_index
    sl_get[0] = var_index
    sl_set[axis] = set_index
    sl_get[axis] = get_index
    s[tuple(sl_set)] = s.dtype.type(factor) * s[tuple(sl_get)]


def boundary_handler(s, o, type_handled=GAS_STATE):
    """boundaries.py:74-441: open, then reflective, then periodic fills, axis by axis."""
    s = s.copy()
    ng = o.num_ghost_cells
    axes = range(1, o.dim + 1)
    for ax in axes:  # open (:270-293)
        left, right = o.boundaries[ax - 1]
        if left == OPEN_BOUNDARY:
            for i in range(ng):
                _set_along_axis(s, ax, i, ng)
        if right == OPEN_BOUNDARY:
            for i in range(-ng, 0):
                _set_along_axis(s, ax, i, -ng - 1)
    for ax in axes:  # reflective (:322-355; 1-D variant :358-385 is the same arithmetic for ng=2)
        left, right = o.boundaries[ax - 1]
        if left == REFLECTIVE_BOUNDARY:
            for i in range(ng):
                if o.dim == 1:
                    g_, s_ = 2, i
                else:
                    g_, s_ = i + ng, ng - i - 1
                _set_along_axis(s, ax, s_, g_)
                _set_specific_var_along_axis(s, ax, s_, g_, ax, -1.0)
        if right == REFLECTIVE_BOUNDARY:
            for i in range(ng):
                if o.dim == 1:
                    g_, s_ = -3, -1 - i
                else:
                    g_, s_ = -i - ng - 1, -ng + i
                _set_along_axis(s, ax, s_, g_)
                _set_specific_var_along_axis(s, ax, s_, g_, ax, -1.0)
    for ax in axes:  # periodic (:298-319)
        left, right = o.boundaries[ax - 1]
        if left == PERIODIC_BOUNDARY and right == PERIODIC_BOUNDARY:
            for i in range(ng):
                _set_along_axis(s, ax, i, i - 2 * ng)
                _set_along_axis(s, ax, -i - 1, 2 * ng - i - 1)
    if o.dim == 2 and o.boundaries[1][0] == MHD_JET_BOUNDARY:
        _jet_left_boundary(s, o, type_handled)
    return s


def _jet_left_boundary(s, o, type_handled):
    """boundaries.py:180-188, :399-441 (2-D, left y side only): the y ghost rows are an open copy; then the
    cells [:, 0:ng, N/2-h : N/2+h] of the padded array -- ghost rows of the X-left side, a window of
    PADDED y indices -- are overwritten with the inflow constants of the array's kind."""
    ng = o.num_ghost_cells
    for i in range(ng):
        _set_along_axis(s, 2, i, ng)
    n = s.shape[1] - 2 * ng  # config.num_cells
    h = int(0.025 / o.grid_spacing)
    B0 = 200 ** 0.5
    to_set = {GAS_STATE: [5 / 3, 800.0, 0.0, 1.0], VELOCITY_ONLY: [800.0, 0.0, 0.0],
              MAGNETIC_FIELD_ONLY: [B0, 0.0, 0.0]}[type_handled]
    s[:, 0:ng, (n // 2 - h):(n // 2 + h)] = np.asarray(to_set, dtype=s.dtype)[:, None, None]


def pad(s, o):
    """time_stepping/_utils.py:65-75 (edge mode, hard-coded 2 cells, Q8)."""
    return np.pad(s, ((0, 0),) + ((2, 2),) * o.dim, mode="edge")


def unpad(s, o):
    """_utils.py:19-61."""
    ng = o.num_ghost_cells
    sl = (slice(None),) + (slice(ng, -ng),) * o.dim
    return s[sl].copy()


# ------------------------------------------------------------------ hydro step
def _bc(s, o, type_handled=GAS_STATE):
    return boundary_handler(s, o, type_handled) if o.boundary_handling == GHOST_CELLS else s


def evolve_state_along_axis(s, dt, gamma, o, axis):
    """_state_evolution/evolve_state.py:53-163 (one split sweep, Cartesian)."""
    s = _bc(s, o)
    U = conserved_state_from_primitive(s, gamma, o)
    if o.first_order_fallback:
        L, R = roll(s, 1, axis), s
    else:
        L, R = reconstruct_split(s, dt, gamma, o, axis)
    F = riemann_solver(L, R, s, gamma, o, axis)
    dU = 1 / o.grid_spacing * stencil_add(F, (0, 1), (1.0, -1.0), axis) * dt
    U = U + dU
    return _bc(primitive_state_from_conserved(U, gamma, o), o)


def evolve_gas_state_split(s, dt, gamma, o):
    """evolve_state.py:168-325: x | x(1/2) y x(1/2) | x(1/2) y(1/2) z y(1/2) x(1/2)."""
    if o.dim == 1:
        seq = [(1, dt)]
    elif o.dim == 2:
        seq = [(1, dt / 2), (2, dt), (1, dt / 2)]
    else:
        seq = [(1, dt / 2), (2, dt / 2), (3, dt), (2, dt / 2), (1, dt / 2)]
    for ax, h in seq:
        s = evolve_state_along_axis(s, h, gamma, o, ax)
    return s


def reconstruct_unsplit_pp(s, gamma, o):
    """reconstruction.py:134-394 with limiter VAN_ALBADA_PP: van-Albada half-differences along every
    axis, scaled per cell so that the interface states keep density, pressure (alpha_density,
    kappa_pressure) and internal energy (beta) positive (Pang & Wu).  Returns the left / right
    interface states of every axis.  The expressions are written exactly as the reference writes
    them, so NumPy's trailing-axis broadcasting reproduces its index pairing: in A1 the per-AXIS
    factor C_d * alpha_d * drho_d lines up with velocity COMPONENT d of the velocity differences
    (A1 = sum_j (C_j alpha_j drho_j sum_i d_i v_j)^2), and in 1-D the sums run over the grid."""
    t = s.dtype.type
    dim = o.dim
    grads = np.stack([calculate_limited_gradients(s, o, axis) for axis in range(1, dim + 1)])
    diff = grads * o.grid_spacing / 2  # [axis][var][grid]
    eps = 1e-14
    rho, p = s[o.rho_i], s[o.p_i]
    c = speed_of_sound(rho, p, gamma)
    alpha_lax = np.stack([np.max(np.abs(s[axis]) + c) for axis in range(1, dim + 1)]).astype(s.dtype)
    C = alpha_lax / np.sum(alpha_lax)
    q = (t(2) if o.mhd else t(1)) / t(o.C_cfl)
    drho, dp = diff[:, o.rho_i], diff[:, o.p_i]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        drho_prot = np.where(np.abs(drho) > eps, drho, t(eps))
        dp_prot = np.where(np.abs(dp) > eps, dp, t(eps))
        alpha_density = np.where(np.abs(drho) > eps,
                                 np.minimum(rho / (np.abs(drho_prot) * (1 + eps)), 1), 1).astype(s.dtype)
        kappa_pressure = np.where(np.abs(dp) > eps,
                                  np.minimum(p / (np.abs(dp_prot) * (1 + eps)), 1), 1).astype(s.dtype)
        Cb = C.reshape((dim,) + (1,) * dim)
        if dim == 1:
            dv = diff[:, 1]  # velocity_index is an int in 1-D: [axis][grid]
            A1 = np.sum(np.sum(Cb * alpha_density * drho * dv, axis=0) ** 2, axis=0)
            A2 = np.sum(Cb * np.sum(dv ** 2, axis=1), axis=0)
        else:
            dv = diff[:, 1:1 + dim]  # [axis][component][grid]
            A1 = np.sum(np.sum(Cb * alpha_density * drho * dv, axis=0) ** 2, axis=0)
            A2 = np.sum(Cb * np.sum(dv ** 2, axis=1), axis=0)
        dvs = diff[:, 1:1 + dim]
        vsum = np.sum(np.sum(dvs ** 2, axis=1), axis=0)
        A1 = np.where(vsum > eps, A1, t(eps))
        A2 = np.where(vsum > eps, A2, t(eps))
        beta = np.where(
            vsum > eps,
            np.minimum(np.sqrt(((q - 2) ** 2 * rho * p) / ((gamma - 1) * (2 * A1 + (q - 2) * rho ** 2 * A2))), 1),
            1).astype(s.dtype)
    diff = diff.copy()
    diff[:, o.rho_i] = drho * alpha_density
    diff[:, o.p_i] = dp * kappa_pressure
    diff[:, 1:1 + dim] = dvs * beta
    Ls, Rs = [], []
    for axis in range(1, dim + 1):
        left_center = s - diff[axis - 1]
        right_center = s + diff[axis - 1]
        Ls.append(roll(right_center, 1, axis))
        Rs.append(left_center)
    return Ls, Rs


def evolve_gas_state_unsplit_inner(s, dt, gamma, o):
    """evolve_state.py:339-414 (one forward-Euler stage; all axes from the same state)."""
    s = _bc(s, o)
    U = conserved_state_from_primitive(s, gamma, o)
    if o.limiter == VAN_ALBADA_PP:
        pls, prs = reconstruct_unsplit_pp(s, gamma, o)
    for axis in range(1, o.dim + 1):
        s = _bc(s, o)
        if o.limiter == VAN_ALBADA_PP:
            L, R = pls[axis - 1], prs[axis - 1]
        else:
            L, R = reconstruct_unsplit_single(s, o, axis)
        F = riemann_solver(L, R, s, gamma, o, axis)
        dU = 1 / o.grid_spacing * stencil_add(F, (0, 1), (1.0, -1.0), axis) * dt
        U = U + dU
    return _bc(primitive_state_from_conserved(U, gamma, o), o)


def evolve_gas_state_unsplit(s, dt, gamma, o):
    """evolve_state.py:419-488 (RK2-SSP; combine without BC afterwards, Q11)."""
    if o.time_integrator != RK2_SSP:
        raise ValueError("Only the RK2 SSP time integrator is currently supported for the unsplit scheme.")
    s1 = evolve_gas_state_unsplit_inner(s, dt, gamma, o)
    s2 = evolve_gas_state_unsplit_inner(s1, dt, gamma, o)
    U0 = conserved_state_from_primitive(s, gamma, o)
    U2 = conserved_state_from_primitive(s2, gamma, o)
    U = s.dtype.type(0.5) * (U0 + U2)
    return primitive_state_from_conserved(U, gamma, o)


def evolve_gas(s, dt, gamma, o):
    if o.split == UNSPLIT:
        return evolve_gas_state_unsplit(s, dt, gamma, o)
    if o.time_integrator != MUSCL and not o.first_order_fallback:
        raise ValueError("Only MUSCL is supported for split reconstruction.")
    return evolve_gas_state_split(s, dt, gamma, o)


# ------------------------------------------------------------------ magnetic update
def cross(a, b):
    """_magnetic_update/_vector_maths.py:6-12."""
    r = np.zeros_like(a)
    r[0] = a[1] * b[2] - a[2] * b[1]
    r[1] = a[2] * b[0] - a[0] * b[2]
    r[2] = a[0] * b[1] - a[1] * b[0]
    return r


def _cdiff(f, axis, dx):
    t = f.dtype.type
    return t(0.5) * (np.roll(f, -1, axis=axis) - np.roll(f, 1, axis=axis)) / dx


def curl3D(f, dx):
    """_vector_maths.py:55-78."""
    c = np.zeros_like(f)
    c[0] = c[0] + (_cdiff(f[2], 1, dx) - _cdiff(f[1], 2, dx))
    c[1] = c[1] + (_cdiff(f[0], 2, dx) - _cdiff(f[2], 0, dx))
    c[2] = c[2] + (_cdiff(f[1], 0, dx) - _cdiff(f[0], 1, dx))
    return c


def curl2D(f, dx):
    """_vector_maths.py:81-100."""
    c = np.zeros_like(f)
    c[0] = c[0] + _cdiff(f[2], 1, dx)
    c[1] = c[1] + (-_cdiff(f[2], 0, dx))
    c[2] = c[2] + (_cdiff(f[1], 0, dx) - _cdiff(f[0], 1, dx))
    return c


def divergence(f, dx, dim):
    """_vector_maths.py:15-52 (central differences; diagnostics only)."""
    out = 0
    for k in range(dim):
        out = out + (np.roll(f[k], -1, axis=k) - np.roll(f[k], 1, axis=k)) / (2 * dx)
    return out


def _norm0(a):
    """jnp.linalg.norm(a, axis=0): sqrt(sum |a|^2) over the component axis."""
    return np.sqrt(np.sum(np.abs(a) ** 2, axis=0))


def magnetic_update(B, gas, dt, o, return_iterations=False):
    """_magnetic_update/_magnetic_field_update.py:37-242 (Pang-Wu system-B fixed point)."""
    dx = o.grid_spacing
    t = B.dtype.type
    if o.dim == 2:
        v = np.zeros((3,) + gas.shape[1:], dtype=B.dtype)
        v[0:2] = gas[1:3]
        curl = curl2D
    elif o.dim == 3:
        v = gas[1:4].copy()
        curl = curl3D
    else:
        raise ValueError("No 1D MHD.")
    rho = gas[o.rho_i]

    def phi(vv, BB):
        E = cross(BB, vv)
        return curl(E, dx), cross(BB, curl(BB, dx)) / rho

    B = _bc(B, o, MAGNETIC_FIELD_ONLY)
    v = _bc(v, o, VELOCITY_ONLY)
    B0, v0 = B, v
    phiA, phiB = phi((v0 + v) / 2, (B0 + B) / 2)
    B1 = _bc(B - dt * phiA, o, MAGNETIC_FIELD_ONLY)
    v1 = _bc(v - dt * phiB, o, VELOCITY_ONLY)
    tol = 1e-5 if o.numerical_precision == SINGLE_PRECISION else 1e-10
    Bk, vk, it = B0, v0, 0
    while max(np.max(_norm0(Bk - B1)), np.max(_norm0(vk - v1))) > tol and it < 1000:
        Bk, vk = B1, v1
        if o.fv_magnetic_integrator == IMPLICIT_MIDPOINT:
            aB, av = (Bk + B) / 2, (vk + v) / 2
        else:
            aB, av = Bk, vk
        phiA, phiB = phi(av, aB)
        B1 = _bc(B - dt * phiA, o, MAGNETIC_FIELD_ONLY)
        v1 = _bc(v - dt * phiB, o, VELOCITY_ONLY)
        it += 1
    Bn, vn = B1, v1
    g53 = 5 / 3  # hard-coded in the reference (Q4)
    E = total_energy_from_primitives(rho, _norm0(v), gas[o.p_i], g53)
    E2 = E - t(0.5) * rho * _norm0(v) ** 2 + t(0.5) * rho * _norm0(vn) ** 2
    p2 = pressure_from_energy(E2, rho, _norm0(vn), g53)
    gas = gas.copy()
    gas[1:1 + o.dim] = vn[:o.dim]
    gas[o.p_i] = p2
    if return_iterations:
        return Bn, gas, it
    return Bn, gas


# ------------------------------------------------------------------ step + CFL + loop
def evolve_state_fv(s, dt, gamma, o, iters=None):
    """evolve_state.py:497-600 (hydro, or Strang: hydro(dt/2) . magnetic(dt) . hydro(dt/2))."""
    if not o.mhd:
        return evolve_gas(s, dt, gamma, o)
    if o.dim == 1:
        raise ValueError("MHD currently not supported in 1D.")
    gas, B = s[:-3], s[-3:]
    gas = evolve_gas(gas, dt / 2, gamma, o)
    B, gas, it = magnetic_update(B, gas, dt, o, return_iterations=True)
    if iters is not None:
        iters.append(it)
    gas = evolve_gas(gas, dt / 2, gamma, o)
    return np.concatenate((gas, B), axis=0)


def cfl_time_step(s, o, gamma, C_cfl, dt_max):
    """_timestep_estimation/_timestep_estimator.py:95-218 (hydro sound speed only, Q3)."""
    dx = o.grid_spacing
    rho, p = s[o.rho_i], s[o.p_i]
    if o.split == UNSPLIT:
        c = speed_of_sound(rho, p, gamma)
        alpha = np.zeros((o.dim,), dtype=s.dtype)
        for axis in range(1, o.dim + 1):
            alpha[axis - 1] = np.max(np.abs(s[axis]) + c)
        return C_cfl * 1 / np.sum(alpha / dx)
    m = None
    for axis in range(1, o.dim + 1):
        sl_l = [slice(None)] * s.ndim
        sl_r = [slice(None)] * s.ndim
        sl_l[axis] = slice(None, -1)
        sl_r[axis] = slice(1, None)
        Ls, Rs = s[tuple(sl_l)], s[tuple(sl_r)]
        cL = speed_of_sound(Ls[o.rho_i], Ls[o.p_i], gamma)
        cR = speed_of_sound(Rs[o.rho_i], Rs[o.p_i], gamma)
        w = max(np.max(np.abs(np.abs(Ls[axis]) + cL)), np.max(np.abs(np.abs(Rs[axis]) + cR)))
        m = w if m is None else max(m, w)
    dt = C_cfl * dx / m
    if o.use_max_adaptive_timestep:
        dt = min(dt, dt_max)
    return s.dtype.type(dt)


def prepare_state(s0, o):
    """time_stepping/time_integration.py:274-286: pad + initial ghost fill of the gas rows (Q9)."""
    s = np.ascontiguousarray(s0)
    if o.boundary_handling == GHOST_CELLS:
        s = pad(s, o)
        if o.mhd:
            s[:-3] = boundary_handler(s[:-3], o)
        else:
            s = boundary_handler(s, o)
    return s


def time_integration(s0, o, max_steps=None, return_info=False, checkpoint_at=None):
    """time_integration.py:239-845, FORWARDS adaptive-dt loop without snapshots.

    ``max_steps`` truncates the loop (parity tests fix the step count);
    ``checkpoint_at`` = step index at which the padded state and time are recorded.
    """
    t_ = s0.dtype.type
    gamma, C_cfl, dt_max, t_end = t_(o.gamma), t_(o.C_cfl), t_(o.dt_max), t_(o.t_end)
    s = prepare_state(s0, o)
    time = t_(0.0)
    steps, iters, dts, ckpt = 0, [], [], None
    while time < t_end and (max_steps is None or steps < max_steps):
        if checkpoint_at is not None and steps == checkpoint_at:
            ckpt = (s.copy(), time)
        dt = t_(cfl_time_step(s, o, gamma, C_cfl, dt_max))
        if o.exact_end_time:
            dt = min(dt, t_end - time)
        s = evolve_state_fv(s, dt, gamma, o, iters)
        time = t_(time + dt)
        dts.append(float(dt))
        steps += 1
    out = unpad(s, o) if o.boundary_handling == GHOST_CELLS else s
    if return_info:
        return out, {"steps": steps, "time": float(time), "iterations": iters, "dts": dts,
                     "padded": s, "checkpoint": ckpt}
    return out


# ------------------------------------------------------------------ snapshots and totals (row f3)
def total_quantities(s, o, box_size, gamma, centers=None):
    """_fluid_equations/total_quantities.py:31-176 on an unpadded state (Cartesian; gas energy only;
    total mass in 2-D / 3-D is sum(rho) * box_size**dim, as the reference writes it)."""
    t = s.dtype.type
    dim, dx = o.dim, o.grid_spacing
    u = get_absolute_velocity(s, o)
    rho, p = s[o.rho_i], s[o.p_i]
    vol = (dx * np.ones_like(rho)) if dim == 1 else dx ** dim
    q = {"total_mass": np.sum(rho * vol) if dim == 1 else np.sum(rho) * box_size ** dim,
         "internal_energy": np.sum(p / (gamma - 1) * vol),
         "kinetic_energy": np.sum(t(0.5) * rho * u ** 2 * vol),
         "total_energy": np.sum(total_energy_from_primitives(rho, u, p, gamma) * vol)}
    if centers is not None:
        d = centers - box_size / 2
        r_hat = d / np.linalg.norm(d, axis=-1, keepdims=True)
        uu = s[1] if dim == 1 else s[1:1 + dim]
        u_radial = np.sum(np.moveaxis(uu, 0, -1) * r_hat, axis=-1)
        q["radial_momentum"] = np.sum(rho * u_radial * vol)
    if o.mhd and dim == 3:
        q["magnetic_divergence"] = np.max(np.abs(divergence(s[-3:], dx, 3)))
    return {k: t(v) for k, v in q.items()}


def time_integration_snapshots(s0, o, num_snapshots, box_size=1.0, centers=None, timepoints=None):
    """time_integration.py:397-570, :628-737: the adaptive loop with a snapshot at the START of every
    step whose time has reached ``current_checkpoint * t_end / num_snapshots`` (one per step at most).
    ``timepoints`` (use_specific_snapshot_timepoints, :542-554, :671-678, :742-750): a snapshot when
    |time - timepoints[cp]| < 1e-12 at the start of a step, dt clipped to timepoints[cp] - time instead of
    t_end - time (jax clamps the index cp to the last entry), and one more after the step that reaches t_end.
    Returns (final unpadded state, dict of time_points / states / totals, number of steps)."""
    t_ = s0.dtype.type
    gamma, C_cfl, dt_max, t_end = t_(o.gamma), t_(o.C_cfl), t_(o.dt_max), t_(o.t_end)
    s = prepare_state(s0, o)
    time, cp, steps = t_(0.0), 0, 0
    rec = {"time_points": np.zeros(num_snapshots, dtype=s0.dtype), "states": np.zeros((num_snapshots,) + s0.shape,
                                                                                    dtype=s0.dtype)}
    tps = None if timepoints is None else np.asarray(timepoints, dtype=s0.dtype)

    def take(cp):
        if cp < num_snapshots:  # (.at[].set past the end is dropped)
            u = unpad(s, o) if o.boundary_handling == GHOST_CELLS else s
            rec["time_points"][cp] = time
            rec["states"][cp] = u
            for k, v in total_quantities(u, o, box_size, gamma, centers).items():
                rec.setdefault(k, np.zeros(num_snapshots, dtype=s0.dtype))[cp] = v
        return cp + 1

    while time < t_end:
        if tps is not None:
            if abs(time - tps[min(cp, len(tps) - 1)]) < 1e-12:
                cp = take(cp)
        elif time >= t_(t_(cp) * t_end) / t_(num_snapshots):
            cp = take(cp)
        dt = t_(cfl_time_step(s, o, gamma, C_cfl, dt_max))
        if tps is not None:
            dt = min(dt, t_(tps[min(cp, len(tps) - 1)] - time))
            if not dt > 0:
                raise RuntimeError("snapshot_timepoints must increase and end at t_end (the reference would spin on dt = 0)")
        elif o.exact_end_time:
            dt = min(dt, t_end - time)
        s = evolve_state_fv(s, dt, gamma, o)
        time = t_(time + dt)
        steps += 1
        if tps is not None and abs(time - t_end) < 1e-12:
            cp = take(cp)
    out = unpad(s, o) if o.boundary_handling == GHOST_CELLS else s
    rec["current_checkpoint"] = cp
    return out, rec, steps


# ------------------------------------------------------------------ turbulent driving (row f1)
def forcing_field(noise, n, box_size):
    """_turbulent_forcing.py:15-82 from the complex noise (3, n, n, n): spectrum k^6 exp(-8 k / k_peak),
    DC mode removed, compressive part projected out, real part of the inverse FFT."""
    k1 = 2.0 * np.pi * np.fft.fftfreq(n, d=box_size / n).astype(noise.real.dtype)
    kx, ky, kz = k1.reshape(n, 1, 1), k1.reshape(1, n, 1), k1.reshape(1, 1, n)
    k2 = kx ** 2 + ky ** 2 + kz ** 2
    kk = np.sqrt(k2)
    kpk = 4.0 * np.pi / box_size
    sqrtP = np.sqrt(kk ** 6 * np.exp(-8.0 * kk / kpk))
    cw = [sqrtP * noise[i] for i in range(3)]
    for c in cw:
        c[0, 0, 0] = 0.0
    k2_safe = np.where(k2 == 0.0, 1.0, k2)
    div_k = (kx * cw[0] + ky * cw[1] + kz * cw[2]) / k2_safe
    div_k[0, 0, 0] = 0.0
    cw = [cw[0] - kx * div_k, cw[1] - ky * div_k, cw[2] - kz * div_k]
    return [np.real(np.fft.ifftn(c)).astype(noise.real.dtype) for c in cw]


def apply_forcing(s, dt, noise, energy_injection_rate, o, box_size):
    """_turbulent_forcing.py:84-125: amp from a*amp^2 + b*amp + c = 0 with c = -Edot dt / dV."""
    t = s.dtype.type
    n = s.shape[-1]
    w = forcing_field(noise, n, box_size)
    dV = o.grid_spacing ** 3
    rho, v = s[o.rho_i], s[1:4]
    tempa = t(0.5) * np.sum(rho * (w[0] ** 2 + w[1] ** 2 + w[2] ** 2))
    tempb = np.sum(rho * v[0] * w[0] + rho * v[1] * w[1] + rho * v[2] * w[2])
    tempc = t(-energy_injection_rate * dt / dV)
    disc = tempb ** 2 - t(4.0) * tempa * tempc
    amp = (-tempb + np.sqrt(disc)) / (t(2.0) * tempa) if (disc >= 0 and abs(tempa) > 1e-10) else t(0.0)
    out = s.copy()
    for i in range(3):
        out[1 + i] = s[1 + i] + t(amp) * w[i]
    return out


def time_integration_forced(s0, o, noises, energy_injection_rate, box_size=1.0):
    """time_integration.py:628-737 with turbulent_forcing: dt from the unforced state (clipped to
    t_end), forcing with that dt, finite-volume step.  ``noises[k]`` is the noise of step k; the
    loop stops when they run out.  PERIODIC_ROLL only (the reference's shapes require it)."""
    assert o.boundary_handling == PERIODIC_ROLL and o.dim == 3
    t_ = s0.dtype.type
    gamma, C_cfl, dt_max, t_end = t_(o.gamma), t_(o.C_cfl), t_(o.dt_max), t_(o.t_end)
    s, time, steps = np.ascontiguousarray(s0), t_(0.0), 0
    while time < t_end and steps < len(noises):
        dt = t_(cfl_time_step(s, o, gamma, C_cfl, dt_max))
        if o.exact_end_time:
            dt = min(dt, t_end - time)
        s = apply_forcing(s, dt, noises[steps], energy_injection_rate, o, box_size)
        s = evolve_state_fv(s, dt, gamma, o)
        time = t_(time + dt)
        steps += 1
    return s, steps, float(time)


def resume_padded(s_padded, time, o, max_steps=None):
    """Continue the loop of ``time_integration`` from a recorded padded state and time."""
    t_ = s_padded.dtype.type
    gamma, C_cfl, dt_max, t_end = t_(o.gamma), t_(o.C_cfl), t_(o.dt_max), t_(o.t_end)
    s, time, steps = s_padded.copy(), t_(time), 0
    while time < t_end and (max_steps is None or steps < max_steps):
        dt = t_(cfl_time_step(s, o, gamma, C_cfl, dt_max))
        if o.exact_end_time:
            dt = min(dt, t_end - time)
        s = evolve_state_fv(s, dt, gamma, o)
        time = t_(time + dt)
        steps += 1
    return (unpad(s, o) if o.boundary_handling == GHOST_CELLS else s), steps
