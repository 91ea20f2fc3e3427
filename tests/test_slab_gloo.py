"""World-size-2 (gloo, CPU) test of the slab driver's host logic: decomposition, halo exchange
pairing, all-reduces, step sequencing.  The kernels are replaced by a test-only backend built
on the oracle; the decomposed run must equal the single-domain oracle run bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from astronomix_b200 import problems as pr
from astronomix_b200 import option_classes as oc
from astronomix_b200.slab import NG, SlabDriver
from oracle import fv_oracle as fo


class OracleBackend:
    """Same interface as slab.AmxBackend, computed with the oracle on the local padded slab
    treated as periodic: only the ghost planes come out wrong, and they are never used."""

    def __init__(self, config, params, dtype):
        self.o = fo.Opts.from_config(config, params)
        self.o.boundary_handling = fo.PERIODIC_ROLL
        self.t = np.float64 if dtype == torch.float64 else np.float32
        self.gamma = self.t(params.gamma)
        self.launches = 0

    def _np(self, x):
        return x.numpy()

    def alpha(self, gas, out):
        g = self._np(gas)[:, NG:-NG]
        c = fo.speed_of_sound(g[0], g[self.o.p_i], self.gamma)
        for k in range(self.o.dim):
            out[k] = float(np.max(np.abs(g[k + 1]) + c))

    def dt_from_alpha(self, alpha, time, t_end, dt_out):
        a = self._np(alpha)[: self.o.dim].astype(self.t)
        if self.o.split == fo.UNSPLIT:
            dt = self.t(self.o.C_cfl) * 1 / np.sum(a / self.o.grid_spacing)
        else:
            dt = self.t(self.o.C_cfl) * self.o.grid_spacing / np.max(a)
            if self.o.use_max_adaptive_timestep:
                dt = min(dt, self.t(self.o.dt_max))
        if time is not None:
            dt = min(dt, self.t(t_end) - self.t(time.item()))
        dt_out.fill_(float(dt))

    def hydro_stage(self, gas_in, gas0, gas_out, dt, scale, alpha_lf, alpha_out):
        h = self.t(dt.item()) * self.t(scale)
        s = fo.evolve_gas_state_unsplit_inner(self._np(gas_in), h, self.gamma, self.o)
        if gas0 is not None:
            U = self.t(0.5) * (fo.conserved_state_from_primitive(self._np(gas0), self.gamma, self.o)
                               + fo.conserved_state_from_primitive(s, self.gamma, self.o))
            s = fo.primitive_state_from_conserved(U, self.gamma, self.o)
        gas_out[:, NG:-NG] = torch.from_numpy(s[:, NG:-NG].copy())
        if alpha_out is not None:
            self.alpha(gas_out, alpha_out)

    def split_sweep(self, gas_in, gas_out, axis, dt, scale, alpha_lf, alpha_out):
        h = self.t(dt.item()) * self.t(scale)
        s = fo.evolve_state_along_axis(self._np(gas_in), h, self.gamma, self.o, axis)
        gas_out[:, NG:-NG] = torch.from_numpy(s[:, NG:-NG].copy())
        if alpha_out is not None:
            self.alpha(gas_out, alpha_out)

    def magnetic_iter(self, B0, gas0, bvk, bvk1, dt, change_out):
        o, dx, h = self.o, self.o.grid_spacing, self.t(dt.item())
        B, g = self._np(B0), self._np(gas0)
        v = np.zeros_like(B)
        v[: o.dim] = g[1:1 + o.dim]
        curl = fo.curl2D if o.dim == 2 else fo.curl3D
        if bvk is None:
            Bk, vk = B, v
            aB, av = (B + B) / 2, (v + v) / 2
        else:
            Bk, vk = self._np(bvk)[:3], self._np(bvk)[3:]
            aB, av = ((Bk + B) / 2, (vk + v) / 2) if o.fv_magnetic_integrator == fo.IMPLICIT_MIDPOINT else (Bk, vk)
        B1 = B - h * curl(fo.cross(aB, av), dx)
        v1 = v - h * (fo.cross(aB, curl(aB, dx)) / g[0])
        inner = (slice(None), slice(NG, -NG))
        change_out[0] = float(np.max(fo._norm0((Bk - B1)[inner])))
        change_out[1] = float(np.max(fo._norm0((vk - v1)[inner])))
        bvk1[:3, NG:-NG] = torch.from_numpy(B1[inner].copy())
        bvk1[3:, NG:-NG] = torch.from_numpy(v1[inner].copy())

    # device-resident loop state of libamx (Ctl), restated on the host
    def magnetic_iter_guarded(self, B0, gas0, bvk, bvk1, dt, change_out, launch_index):
        assert (launch_index == 0) == (bvk is None)
        if launch_index == 0:
            self.active, self.iters, self.final, self.fix, self.emitted = True, 0, 0, None, False
        if launch_index > 0 and not self.active:
            change_out.zero_()  # a skipped launch leaves zeroed slots behind
            return
        self.magnetic_iter(B0, gas0, bvk, bvk1, dt, change_out)

    def magnetic_check(self, change, launch_index):
        if not self.active:
            return
        tol = 1e-5 if self.o.numerical_precision == fo.SINGLE_PRECISION else 1e-10
        m = self.t(max(change[0].item(), change[1].item()))
        if m > self.t(tol) and self.iters < 1000:
            self.iters += 1
        else:
            self.active, self.final = False, launch_index & 1

    def magnetic_state(self):
        return self.active, self.iters

    # pair launches (amx_pair.cu): two evaluations on the padded local slab; the intermediate iterate is
    # right on planes -1 .. n (its outermost plane is the periodic-wrap garbage of the first evaluation),
    # which is exactly what the second evaluation needs on planes 0 .. n-1
    pairs_enabled = False

    def pair_supported(self):
        return self.pairs_enabled and self.o.dim == 3

    def _eval_padded(self, B0, gas0, Bk, vk):
        o, dx = self.o, self.o.grid_spacing
        B, g = self._np(B0), self._np(gas0)
        v = np.zeros_like(B)
        v[: o.dim] = g[1:1 + o.dim]
        curl = fo.curl2D if o.dim == 2 else fo.curl3D
        if Bk is None:
            Bk, vk = B, v
        aB, av = ((Bk + B) / 2, (vk + v) / 2) if o.fv_magnetic_integrator == fo.IMPLICIT_MIDPOINT else (Bk, vk)
        h = self.t(self._dt)
        return B - h * curl(fo.cross(aB, av), dx), v - h * (fo.cross(aB, curl(aB, dx)) / g[0]), Bk, vk

    def magnetic_pair_guarded(self, B0, gas0, bvk, bvk2, dt, change4_out, evals_before, emit_gas=None,
                              emit_B=None):
        assert (evals_before == 0) == (bvk is None)
        if evals_before == 0:
            self.active, self.iters, self.final, self.fix, self.emitted = True, 0, 0, None, False
        if evals_before > 0 and not self.active:
            change4_out.zero_()
            return
        self._dt = dt.item()
        inner = (slice(None), slice(NG, -NG))
        Bk = None if bvk is None else self._np(bvk)[:3]
        vk = None if bvk is None else self._np(bvk)[3:]
        B1, v1, Bk, vk = self._eval_padded(B0, gas0, Bk, vk)
        B2, v2, _, _ = self._eval_padded(B0, gas0, B1, v1)
        change4_out[0] = float(np.max(fo._norm0((Bk - B1)[inner])))
        change4_out[1] = float(np.max(fo._norm0((vk - v1)[inner])))
        change4_out[2] = float(np.max(fo._norm0((B1 - B2)[inner])))
        change4_out[3] = float(np.max(fo._norm0((v1 - v2)[inner])))
        bvk2[:3, NG:-NG] = torch.from_numpy(B2[inner].copy())
        bvk2[3:, NG:-NG] = torch.from_numpy(v2[inner].copy())
        if emit_gas is not None:  # the write-back of this launch's result, valid if the loop ends here
            self.magnetic_final(gas0, bvk2, emit_B, None, gas_out=emit_gas)

    def magnetic_check_pair(self, change4, evals_before, buffer_index, emit=False):
        if not self.active:
            return
        tol = self.t(1e-5 if self.o.numerical_precision == fo.SINGLE_PRECISION else 1e-10)
        m1 = self.t(max(change4[0].item(), change4[1].item()))
        m2 = self.t(max(change4[2].item(), change4[3].item()))
        if not (m1 > tol and self.iters < 1000):
            self.active, self.final, self.fix = False, buffer_index, evals_before
            return
        self.iters += 1
        if m2 > tol and self.iters < 1000:
            self.iters += 1
        else:
            self.active, self.final, self.emitted = False, buffer_index, bool(emit)

    def magnetic_fix_guarded(self, B0, gas0, bvk, bvk1, dt, evals_before):
        if self.fix is None or self.fix != evals_before:
            return
        self.fix = None
        scratch = torch.zeros(2, dtype=bvk1.dtype)
        self.magnetic_iter(B0, gas0, bvk, bvk1, dt, scratch)

    def magnetic_final_guarded(self, gas, gas_out, bv0, bv1, B_out, alpha_out):
        if getattr(self, "emitted", False):
            return
        self.magnetic_final(gas, (bv0, bv1)[self.final], B_out, alpha_out, gas_out=gas_out)

    def magnetic_final(self, gas, bv_final, B_out, alpha_out, gas_out=None):
        o, t = self.o, self.t
        if gas_out is not None and gas_out is not gas:
            gas_out.copy_(gas)
            gas = gas_out
        g, X = self._np(gas), self._np(bv_final)
        v = np.zeros_like(X[:3])
        v[: o.dim] = g[1:1 + o.dim]
        vn, rho = X[3:], g[0]
        g53 = 5 / 3
        E = fo.total_energy_from_primitives(rho, fo._norm0(v), g[o.p_i], g53)
        E2 = E - t(0.5) * rho * fo._norm0(v) ** 2 + t(0.5) * rho * fo._norm0(vn) ** 2
        p2 = fo.pressure_from_energy(E2, rho, fo._norm0(vn), g53)
        g[1:1 + o.dim] = vn[: o.dim]
        g[o.p_i] = p2
        B_out.copy_(bv_final[:3])
        if alpha_out is not None:
            self.alpha(gas, alpha_out)


def _worker(rank, world, port, case, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim, n, mhd, split, steps = case[:5]
        pairs = len(case) > 5 and case[5]
        kw = dict(split=oc.SPLIT, time_integrator=oc.MUSCL) if split else {}
        config, params, rv = pr.make(dim, n, mhd=mhd, dtype=np.float64, riemann_solver=oc.HLLC, **kw)
        s = pr.smooth_random(dim, n, mhd=mhd, seed=17, dtype=np.float64)
        backend = OracleBackend(config, params, torch.float64)
        backend.pairs_enabled = bool(pairs)
        drv = SlabDriver(config, params, rv, (n,) * dim, torch.float64, torch.device("cpu"), backend=backend)
        assert drv.pairs == bool(pairs)
        drv.pair_greedy = pairs == "greedy"  # odd evaluation counts then end on the first half of a pair
        drv.scatter(torch.from_numpy(s))
        info = drv.run(max_steps=steps)
        full = drv.gather()
        if rank == 0:
            np.save(out_path, full.numpy())
            np.save(out_path + ".iters.npy", np.array([info.steps, info.magnetic_iterations, drv.collectives]))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case", [(2, 16, False, False, 3), (2, 16, True, False, 3), (3, 8, True, False, 2),
                                  (2, 16, True, True, 2), (3, 8, True, False, 3, "pairs"),
                                  (3, 8, True, False, 3, "greedy")])
def test_two_rank_slab_equals_single_domain(case, tmp_path):
    dim, n, mhd, split, steps = case[:5]
    out = str(tmp_path / "slab.npy")
    mp.spawn(_worker, args=(2, _free_port(), case, out), nprocs=2, join=True)
    got = np.load(out)
    nsteps, iters, colls = np.load(out + ".iters.npy")
    kw = dict(split=oc.SPLIT, time_integrator=oc.MUSCL) if split else {}
    config, params, rv = pr.make(dim, n, mhd=mhd, dtype=np.float64, riemann_solver=oc.HLLC, **kw)
    s = pr.smooth_random(dim, n, mhd=mhd, seed=17, dtype=np.float64)
    ref, info = fo.time_integration(s, fo.Opts.from_config(config, params), max_steps=steps, return_info=True)
    assert nsteps == info["steps"] == steps
    if mhd:
        assert iters == sum(info["iterations"])
    assert colls > 0
    assert np.array_equal(got, ref), np.abs(got - ref).max()


def test_single_rank_driver_wraps_locally():
    """world = 1 (no process group): the driver's own periodic wrap."""
    config, params, rv = pr.make(2, 16, mhd=True, dtype=np.float64)
    s = pr.smooth_random(2, 16, mhd=True, seed=2, dtype=np.float64)
    drv = SlabDriver(config, params, rv, (16, 16), torch.float64, torch.device("cpu"),
                     backend=OracleBackend(config, params, torch.float64))
    drv.scatter(torch.from_numpy(s))
    drv.run(max_steps=3)
    ref = fo.time_integration(s, fo.Opts.from_config(config, params), max_steps=3)
    assert np.array_equal(drv.gather().numpy(), ref)


# ---- turbulent driving on slab ranks (forcing.SlabForcing): distributed inverse FFT + global sums ----
def _forcing_worker(rank, world, port, n, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from astronomix_b200.forcing import SlabForcing
        from oracle import threefry as tf
        tfc = oc.TurbulentForcingConfig(turbulent_forcing=True)
        config, params, rv = pr.make(3, n, mhd=True, dtype=np.float64, boundary_handling=oc.PERIODIC_ROLL,
                                     turbulent_forcing_config=tfc)
        _, sk1, sk2 = tf.split(tf.seed_key(42), 3)
        noise = torch.complex(torch.from_numpy(tf.normal(sk1, (3, n, n, n), np.float64)),
                              torch.from_numpy(tf.normal(sk2, (3, n, n, n), np.float64)))
        f = SlabForcing(config, params, rv, torch.float64, torch.device("cpu"), rank, world,
                        noise_fn=lambda step: noise)
        s = torch.from_numpy(pr.smooth_random(3, n, mhd=True, seed=9, dtype=np.float64))
        nl = n // world
        loc = s[:, rank * nl:(rank + 1) * nl].clone()
        f.apply(loc, torch.tensor(1.0e-3, dtype=torch.float64))
        parts = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(parts, loc.contiguous())
        if rank == 0:
            np.save(out_path, torch.cat(parts, dim=1).numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_forcing_equals_single_domain_forcing(world, tmp_path):
    """k_x-slab spectrum + (2-D FFT, all-to-all, 1-D FFT) inverse transform + all-reduced sums against the oracle's
    whole-domain forcing on the same (restated Threefry) noise: round-off agreement (another FFT factorisation)."""
    n = 16
    out = str(tmp_path / "forced.npy")
    mp.spawn(_forcing_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    got = np.load(out)
    from oracle import threefry as tf
    tfc = oc.TurbulentForcingConfig(turbulent_forcing=True)
    config, params, rv = pr.make(3, n, mhd=True, dtype=np.float64, boundary_handling=oc.PERIODIC_ROLL,
                                 turbulent_forcing_config=tfc)
    _, sk1, sk2 = tf.split(tf.seed_key(42), 3)
    noise = tf.normal(sk1, (3, n, n, n), np.float64) + 1j * tf.normal(sk2, (3, n, n, n), np.float64)
    s = pr.smooth_random(3, n, mhd=True, seed=9, dtype=np.float64)
    ref = fo.apply_forcing(s, np.float64(1.0e-3), noise, params.turbulent_forcing_params.energy_injection_rate,
                           fo.Opts.from_config(config, params), config.box_size)
    assert pr.rel_l1(got, ref, 3, True) <= 1e-12
    assert np.array_equal(got[0], s[0]) and np.array_equal(got[4:], s[4:])


def _forced_driver_worker(rank, world, port, n, steps, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from astronomix_b200.forcing import SlabForcing
        from oracle import threefry as tf
        tfc = oc.TurbulentForcingConfig(turbulent_forcing=True)
        config, params, rv = pr.make(3, n, mhd=True, dtype=np.float64, boundary_handling=oc.PERIODIC_ROLL,
                                     turbulent_forcing_config=tfc)
        params = params._replace(t_end=1.0e9)
        noises, key = [], tf.seed_key(42)
        for _ in range(steps):
            key, sk1, sk2 = tf.split(key, 3)
            noises.append(torch.complex(torch.from_numpy(tf.normal(sk1, (3, n, n, n), np.float64)),
                                        torch.from_numpy(tf.normal(sk2, (3, n, n, n), np.float64))))
        s = pr.smooth_random(3, n, mhd=True, seed=21, dtype=np.float64)
        drv = SlabDriver(config, params, rv, (n,) * 3, torch.float64, torch.device("cpu"),
                         backend=OracleBackend(config, params, torch.float64))
        drv.forcing = SlabForcing(config, params, rv, torch.float64, torch.device("cpu"), rank, world,
                                  noise_fn=lambda step: noises[step])
        drv.scatter(torch.from_numpy(s))
        drv.run(max_steps=steps)
        full = drv.gather()
        if rank == 0:
            np.save(out_path, full.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_driven_loop_equals_single_domain(tmp_path):
    """dt from the unforced state -> driving (slab FFT, all-reduced sums) -> MHD step, on two slab ranks against
    the oracle's whole-domain loop (time_integration.py:699-711): the ghost planes of the forced state are
    exchanged again, not taken from what the previous step's kernels stored."""
    n, steps = 8, 2
    out = str(tmp_path / "forced_loop.npy")
    mp.spawn(_forced_driver_worker, args=(2, _free_port(), n, steps, out), nprocs=2, join=True)
    got = np.load(out)
    from oracle import threefry as tf
    tfc = oc.TurbulentForcingConfig(turbulent_forcing=True)
    config, params, rv = pr.make(3, n, mhd=True, dtype=np.float64, boundary_handling=oc.PERIODIC_ROLL,
                                 turbulent_forcing_config=tfc)
    params = params._replace(t_end=1.0e9)
    noises, key = [], tf.seed_key(42)
    for _ in range(steps):
        key, sk1, sk2 = tf.split(key, 3)
        noises.append(tf.normal(sk1, (3, n, n, n), np.float64) + 1j * tf.normal(sk2, (3, n, n, n), np.float64))
    s = pr.smooth_random(3, n, mhd=True, seed=21, dtype=np.float64)
    ref, nsteps, _ = fo.time_integration_forced(s, fo.Opts.from_config(config, params), noises,
                                                params.turbulent_forcing_params.energy_injection_rate, config.box_size)
    assert nsteps == steps
    assert pr.rel_l1(got, ref, 3, True) <= 1e-11
