"""Snapshot / diagnostic path of time_integration.

Reference: time_stepping/time_integration.py:297-369 (arrays), :397-570 (a snapshot is taken at the START
of a step whenever ``time >= current_checkpoint * t_end / num_snapshots``, one per step at most),
:811-830 (return), _fluid_equations/total_quantities.py:31-176 (the totals), data_classes/
simulation_snapshot_data.py (SnapshotData).  The loop itself stays on the device: the segments
between two snapshots are single ``amx_integrate_from`` calls that resume at the current time and
stop as soon as the next snapshot time has been passed (dt is clipped against t_end only, as in the
reference).  The totals are reductions over the device state (torch; plumbing, not the hot path).
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import numpy as np
import torch

from . import _lib
from . import option_classes as oc
from .helpers import get_helper_data


class SnapshotData(NamedTuple):
    """data_classes/simulation_snapshot_data.py:5-47 (same fields, same order)."""
    time_points: object = None
    states: object = None
    final_state: object = None
    total_mass: object = None
    total_energy: object = None
    internal_energy: object = None
    kinetic_energy: object = None
    gravitational_energy: object = None
    radial_momentum: object = None
    magnetic_divergence: object = None
    runtime: float = 0.0
    num_iterations: int = 0
    current_checkpoint: int = 0


def _absolute_velocity(s, dim):
    """_equations.py:165-199 (1e-20 only in 3-D, Q2)."""
    if dim == 1:
        return s[1].abs()
    if dim == 2:
        return torch.sqrt(s[1] ** 2 + s[2] ** 2)
    return torch.sqrt(s[1] ** 2 + s[2] ** 2 + s[3] ** 2 + 1e-20)


class Diagnostics:
    """total_quantities.py:31-176 on an unpadded device state."""

    def __init__(self, config, params, registered_variables, dtype, device):
        self.c, self.rv, self.dim = config, registered_variables, config.dimensionality
        self.gamma = torch.tensor(params.gamma, dtype=dtype, device=device)
        self.dV = config.grid_spacing ** self.dim
        np_dtype = np.float32 if dtype == torch.float32 else np.float64
        h = get_helper_data(config._replace(return_snapshots=True), dtype=np_dtype)
        self.cell_volumes = torch.from_numpy(h.cell_volumes).to(device) if self.dim == 1 else None
        self.centers = torch.from_numpy(np.ascontiguousarray(h.geometric_centers)).to(device)

    def _sum(self, q):
        return torch.sum(q * self.cell_volumes) if self.dim == 1 else torch.sum(q * self.dV)

    def total_mass(self, s):  # :155-176: in 2-D / 3-D the reference multiplies by box_size**dim
        if self.dim == 1:
            return torch.sum(s[0] * self.cell_volumes)
        return torch.sum(s[0]) * self.c.box_size ** self.dim

    def internal_energy(self, s):
        return self._sum(s[self.rv.pressure_index] / (self.gamma - 1))

    def kinetic_energy(self, s):
        return self._sum(0.5 * s[0] * _absolute_velocity(s, self.dim) ** 2)

    def total_energy(self, s):  # gas energy only: the reference does not add B^2/2
        u = _absolute_velocity(s, self.dim)
        return self._sum(s[self.rv.pressure_index] / (self.gamma - 1) + 0.5 * s[0] * u ** 2)

    def radial_momentum(self, s):
        d = self.centers - self.c.box_size / 2
        r_hat = d / torch.linalg.norm(d, dim=-1, keepdim=True)
        u = s[1] if self.dim == 1 else s[1:1 + self.dim]
        u_radial = torch.sum(torch.movedim(u, 0, -1) * r_hat, dim=-1) if self.dim > 1 else torch.sum(u * r_hat, dim=-1)
        return self._sum(s[0] * u_radial)

    def magnetic_divergence(self, s):
        """max |divergence3D(B)| (_vector_maths.py:34-52: central differences with periodic wrap)."""
        if self.dim != 3:
            raise NotImplementedError("the reference evaluates divergence3D, which needs a 3-D grid")
        B = s[-3:]
        div = sum((torch.roll(B[k], -1, dims=k) - torch.roll(B[k], 1, dims=k)) / (2 * self.c.grid_spacing)
                  for k in range(3))
        return torch.max(torch.abs(div))


def integrate_with_snapshots(t, kind, config, params, registered_variables, snapshot_callable, from_device):
    """The snapshot / callback variants of time_integration on the device tensor ``t`` (unpadded)."""
    if config.snapshot_settings.return_gravitational_energy and config.self_gravity:
        raise NotImplementedError("self-gravity is outside the finite-volume hot path")
    lib = _lib.load()
    tdt = t.dtype
    npdt = np.float32 if tdt == torch.float32 else np.float64
    desc = _lib.make_desc(config, params, t.shape[1:], _lib.AMX_F32 if tdt == torch.float32 else _lib.AMX_F64)
    nbytes = C.c_size_t()
    _lib.check(lib.amx_workspace_bytes(C.byref(desc), C.byref(nbytes)))
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=t.device)
    ws_ptr = C.c_void_p((ws.data_ptr() + 255) & ~255)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    stats = _lib.AmxStats()
    cur, nxt = t.clone(), torch.empty_like(t)
    N = int(config.num_snapshots)
    st = config.snapshot_settings
    snaps = config.return_snapshots
    diag = Diagnostics(config, params, registered_variables, tdt, t.device) if snaps else None
    want = {}
    if snaps:
        want = {"total_mass": st.return_total_mass, "total_energy": st.return_total_energy,
                "internal_energy": st.return_internal_energy, "kinetic_energy": st.return_kinetic_energy,
                "radial_momentum": st.return_radial_momentum,
                "magnetic_divergence": st.return_magnetic_divergence and config.mhd}
    series = {k: torch.zeros(N, dtype=tdt, device=t.device) for k, v in want.items() if v}
    time_points = np.zeros(N, dtype=npdt)
    states = torch.zeros((N,) + tuple(t.shape), dtype=tdt, device=t.device) if snaps and st.return_states else None
    t_end = npdt(params.t_end)
    time, cp, iters, launches, mag_iters = npdt(0.0), 0, 0, 0, 0

    def take(cp):
        if snaps and cp < N:
            time_points[cp] = time
            if states is not None:
                states[cp].copy_(cur)
            for k in series:
                series[k][cp] = getattr(diag, k)(cur)
        elif not snaps:
            snapshot_callable(float(time), cur, registered_variables)
        return cp + 1

    if config.use_specific_snapshot_timepoints and snaps:
        # time_integration.py:542-554, :671-678, :742-750: a snapshot at the start of a step when
        # |time - snapshot_timepoints[cp]| < 1e-12, dt clipped to snapshot_timepoints[cp] - time INSTEAD of
        # t_end - time (the index is clamped to the last entry, as jax clamps it), and one more snapshot after
        # the step that reaches t_end.  A segment up to the next time point is one call of the device-resident
        # loop with that point as the time the last dt is clipped to.
        tps = np.asarray(params.snapshot_timepoints, dtype=npdt).reshape(-1)
        desc_tp = _lib.make_desc(config._replace(exact_end_time=True), params, t.shape[1:],
                                 _lib.AMX_F32 if tdt == torch.float32 else _lib.AMX_F64)
        while time < t_end:
            target = tps[min(cp, len(tps) - 1)]
            if abs(time - target) < 1e-12:
                cp = take(cp)
                target = tps[min(cp, len(tps) - 1)]
            if not target > time:
                raise ValueError("snapshot_timepoints must increase and end at t_end (the reference's loop would "
                                 "not advance: dt = snapshot_timepoints[current_checkpoint] - time <= 0)")
            with torch.cuda.device(t.device):
                _lib.check(lib.amx_integrate_from(C.byref(desc_tp), C.c_void_p(cur.data_ptr()),
                                                  C.c_void_p(nxt.data_ptr()), float(time), float("inf"),
                                                  float(target), -1, ws_ptr, stream, C.byref(stats)))
            if stats.steps == 0:
                raise RuntimeError("time integration made no progress")
            cur, nxt = nxt, cur
            time = npdt(stats.time)
            iters += stats.steps
            launches += stats.kernel_launches
            mag_iters += stats.magnetic_iterations
            if abs(time - t_end) < 1e-12:
                cp = take(cp)
    while time < t_end and not (config.use_specific_snapshot_timepoints and snaps):
        thr = npdt(npdt(cp) * t_end) / npdt(N)  # cp * t_end / num_snapshots in the working dtype
        if time >= thr:
            cp = take(cp)
            thr = npdt(npdt(cp) * t_end) / npdt(N)
        # The next test of the snapshot condition is at the start of the NEXT step: one step runs in any
        # case.  Usually time < thr and a single call steps until thr has been passed (or t_end is
        # reached); if one step jumped over two snapshot times (time >= thr already), the reference
        # takes the second snapshot one step later, so exactly one step is made.
        one_step = bool(time >= thr)
        with torch.cuda.device(t.device):
            _lib.check(lib.amx_integrate_from(C.byref(desc), C.c_void_p(cur.data_ptr()), C.c_void_p(nxt.data_ptr()),
                                              float(time), float("inf") if one_step else float(thr),
                                              float(params.t_end), 1 if one_step else -1, ws_ptr, stream,
                                              C.byref(stats)))
        if stats.steps == 0:
            raise RuntimeError("time integration made no progress")
        cur, nxt = nxt, cur
        time = npdt(stats.time)
        iters += stats.steps
        launches += stats.kernel_launches
        mag_iters += stats.magnetic_iterations
    info = (iters, float(time), mag_iters, launches)
    if not snaps:
        return cur, None, info
    out = {k: v.cpu().numpy() for k, v in series.items()}
    data = SnapshotData(
        time_points=time_points, states=None if states is None else from_device(states, kind),
        final_state=from_device(cur, kind) if st.return_final_state else None,
        total_mass=out.get("total_mass"), total_energy=out.get("total_energy"),
        internal_energy=out.get("internal_energy"), kinetic_energy=out.get("kinetic_energy"),
        gravitational_energy=None, radial_momentum=out.get("radial_momentum"),
        magnetic_divergence=out.get("magnetic_divergence"), runtime=0.0, num_iterations=iters,
        current_checkpoint=cp)
    return cur, data, info
