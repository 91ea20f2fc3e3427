"""Slab-decomposed multi-GPU driver: one process per GPU, x axis split across ranks.

The reference has no explicit communication: users shard the ``[var, x, y, z]`` array with a
jax mesh and GSPMD inserts collective-permutes for the rolls and all-reduces for the maxima
(tests/multi_gpu/scaling.py:136-146, arena/arena_tests/scaling/scaling.py:93-98).
Here the same partitioning is explicit:

  * every rank owns ``n_x / P`` planes plus 2 ghost planes per side (``[rows][nxl+4][ny][nz]``);
  * before each kernel that reads x-neighbours the ghost planes are refreshed by a batched NCCL
    send/recv with the two neighbour ranks (periodic wrap = rank 0 <-> rank P-1; open /
    reflective physical boundaries are filled locally on the edge ranks);
  * the per-axis wave bounds (CFL dt, Lax-Friedrichs alpha) and the two fixed-point change norms
    are all-reduced with MAX -- max is order-independent, so N-GPU results are bitwise identical
    to 1-GPU results;
  * dt and the simulation time stay on the device (scalars of the working dtype).

All arithmetic is done by libamx through the stage-level C-ABI entry points (``AmxBackend``);
the driver itself only moves halos and scalars, which is what the gloo tests exercise on CPU
with a test-only backend.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch
import torch.distributed as dist

from . import _lib
from . import option_classes as oc

NG = 2  # ghost planes per side (config.num_ghost_cells, Q8)


class StepInfo(NamedTuple):
    steps: int
    time: float
    magnetic_iterations: int
    launches: int


class AmxBackend:
    """Stage-level kernels of libamx on local slab tensors (CUDA)."""

    def __init__(self, config, params, local_n, dtype, device, arithmetic=None):
        _lib.require_device()
        self.lib = _lib.load()
        self.device = device
        self.dtype = dtype
        code = _lib.AMX_F32 if dtype == torch.float32 else _lib.AMX_F64
        self.desc = _lib.make_desc(config, params, local_n, code, arithmetic=arithmetic, halo_x=True)
        nbytes = C.c_size_t()
        _lib.check(self.lib.amx_workspace_bytes(C.byref(self.desc), C.byref(nbytes)))
        # only the control block of the workspace is used by the stage-level entry points
        self.ws = torch.zeros(min(nbytes.value, 1 << 20) + 256, dtype=torch.uint8, device=device)
        self.ws_ptr = C.c_void_p((self.ws.data_ptr() + 255) & ~255)
        self.launches = 0

    def _s(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def set_shell_sides(self, left_code, right_code):
        """The stage kernels see AMX_BC_HALO on both x sides (the driver fills the ghost planes of a physical side
        itself); the ghost shell has to know which x side of this rank is a physical boundary: its ghost cells
        there are shell cells, whereas the planes beyond a halo side belong to the neighbour."""
        self.desc_shell = _lib.AmxDesc.from_buffer_copy(self.desc)
        self.desc_shell.peer_flags = 0
        if left_code is not None:
            self.desc_shell.bc[0][0] = int(left_code)
        if right_code is not None:
            self.desc_shell.bc[0][1] = int(right_code)

    def shell_update(self, mode, shell, u0=None, u1=None, bv0=None, bv1=None, alpha_out=None):
        """ghost shell of the gas state on this rank (include/amx.h amx_shell_update)"""
        _lib.check(self.lib.amx_shell_update(C.byref(self.desc_shell), int(mode), self._p(shell), self._p(u0), self._p(u1),
                                             self._p(bv0), self._p(bv1), self._p(alpha_out), self.ws_ptr, self._s()))
        self.launches += 2 if alpha_out is not None else 1

    # ---- peer memory (include/amx.h "peer-memory plumbing")
    def set_peers(self, flags, delta_lo, delta_hi):
        """Arrays handed to the stage entry points from now on live in peer-mapped memory: kernels also
        store their edge planes into the neighbours' ghost planes.  Returns amx_peer_push_supported."""
        self.desc.peer_flags = int(flags)
        self.desc.peer_delta_lo = int(delta_lo)
        self.desc.peer_delta_hi = int(delta_hi)
        return int(self.lib.amx_peer_push_supported(C.byref(self.desc))) if flags else 0

    def peer_comm_bytes(self):
        return int(self.lib.amx_peer_comm_bytes())

    def peer_allreduce_max(self, comm_ptrs, rank, world, vals, count):
        code = _lib.AMX_F32 if self.dtype == torch.float32 else _lib.AMX_F64
        _lib.check(self.lib.amx_peer_allreduce_max(comm_ptrs, int(rank), int(world), self._p(vals), int(count), code,
                                                   self._s()))
        self.launches += 1

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def alpha(self, gas, out):
        _lib.check(self.lib.amx_alpha(C.byref(self.desc), self._p(gas), self._p(out), self.ws_ptr, self._s()))
        self.launches += 2

    def dt_from_alpha(self, alpha, time, t_end, dt_out):
        _lib.check(self.lib.amx_dt_from_alpha(C.byref(self.desc), self._p(alpha), self._p(time), float(t_end),
                                              self._p(dt_out), self._s()))
        self.launches += 1

    def hydro_stage(self, gas_in, gas0, gas_out, dt, scale, alpha_lf, alpha_out):
        _lib.check(self.lib.amx_hydro_stage(C.byref(self.desc), self._p(gas_in), self._p(gas0), self._p(gas_out),
                                            self._p(dt), float(scale), self._p(alpha_lf), self._p(alpha_out),
                                            self.ws_ptr, self._s()))
        self.launches += 2 if alpha_out is not None else 1

    def split_sweep(self, gas_in, gas_out, axis, dt, scale, alpha_lf, alpha_out):
        _lib.check(self.lib.amx_split_sweep(C.byref(self.desc), self._p(gas_in), self._p(gas_out), int(axis),
                                            self._p(dt), float(scale), self._p(alpha_lf), self._p(alpha_out),
                                            self.ws_ptr, self._s()))
        self.launches += 2 if alpha_out is not None else 1

    def magnetic_iter(self, B0, gas0, bvk, bvk1, dt, change_out):
        _lib.check(self.lib.amx_magnetic_iter(C.byref(self.desc), self._p(B0), self._p(gas0), self._p(bvk),
                                              self._p(bvk1), self._p(dt), self._p(change_out), self.ws_ptr,
                                              self._s()))
        self.launches += 2

    def magnetic_iter_guarded(self, B0, gas0, bvk, bvk1, dt, change_out, launch_index):
        _lib.check(self.lib.amx_magnetic_iter_guarded(C.byref(self.desc), self._p(B0), self._p(gas0), self._p(bvk),
                                                      self._p(bvk1), self._p(dt), self._p(change_out),
                                                      int(launch_index), self.ws_ptr, self._s()))
        self.launches += 2

    def magnetic_check(self, change, launch_index):
        _lib.check(self.lib.amx_magnetic_check(C.byref(self.desc), self._p(change), int(launch_index), self.ws_ptr,
                                               self._s()))
        self.launches += 1

    def magnetic_final_guarded(self, gas, gas_out, bv0, bv1, B_out, alpha_out):
        _lib.check(self.lib.amx_magnetic_final_guarded(C.byref(self.desc), self._p(gas), self._p(gas_out),
                                                       self._p(bv0), self._p(bv1), self._p(B_out),
                                                       self._p(alpha_out), self.ws_ptr, self._s()))
        self.launches += 2 if alpha_out is not None else 1

    def pair_supported(self):
        return bool(self.lib.amx_magnetic_pair_supported(C.byref(self.desc)))

    def magnetic_pair_guarded(self, B0, gas0, bvk, bvk2, dt, change4_out, evals_before, emit_gas=None,
                              emit_B=None):
        _lib.check(self.lib.amx_magnetic_pair_guarded(C.byref(self.desc), self._p(B0), self._p(gas0), self._p(bvk),
                                                      self._p(bvk2), self._p(dt), self._p(change4_out),
                                                      int(evals_before), self._p(emit_gas), self._p(emit_B),
                                                      self.ws_ptr, self._s()))
        self.launches += 3

    def magnetic_check_pair(self, change4, evals_before, buffer_index, emit=False):
        _lib.check(self.lib.amx_magnetic_check_pair(C.byref(self.desc), self._p(change4), int(evals_before),
                                                    int(buffer_index), int(bool(emit)), self.ws_ptr, self._s()))
        self.launches += 1

    def magnetic_fix_guarded(self, B0, gas0, bvk, bvk1, dt, evals_before):
        _lib.check(self.lib.amx_magnetic_fix_guarded(C.byref(self.desc), self._p(B0), self._p(gas0), self._p(bvk),
                                                     self._p(bvk1), self._p(dt), int(evals_before), self.ws_ptr,
                                                     self._s()))
        self.launches += 1

    def magnetic_state(self):
        """(still iterating, while-body iterations so far); synchronises the stream"""
        active, iters = C.c_int32(), C.c_int32()
        _lib.check(self.lib.amx_magnetic_state(C.byref(self.desc), self.ws_ptr, self._s(), C.byref(active),
                                               C.byref(iters)))
        return bool(active.value), int(iters.value)

    def magnetic_final(self, gas, bv_final, B_out, alpha_out):
        _lib.check(self.lib.amx_magnetic_final(C.byref(self.desc), self._p(gas), self._p(bv_final), self._p(B_out),
                                               self._p(alpha_out), self.ws_ptr, self._s()))
        self.launches += 2 if alpha_out is not None else 1


class SlabDriver:
    """Time loop of time_integration.py:239-845 on an x-slab decomposition."""

    def __init__(self, config, params, registered_variables, global_n, dtype, device, backend=None,
                 group=None, arithmetic=None):
        if config.boundary_settings is None:
            raise ValueError("finalize_config first")
        self.config, self.params, self.rv = config, params, registered_variables
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dim = config.dimensionality
        if self.dim < 2:
            raise NotImplementedError("slab decomposition needs at least 2 dimensions")
        self.global_n = tuple(int(v) for v in global_n)
        if self.global_n[0] % self.world:
            raise ValueError("n_x must be divisible by the number of ranks")
        self.nxl = self.global_n[0] // self.world
        if self.nxl < 2 * NG:
            raise ValueError("slab too thin: need at least 4 planes per rank")
        self.local_n = (self.nxl,) + self.global_n[1:]
        codes = oc.boundary_codes(config)
        # MHD jet inflow (boundaries.py:399-441): constants in a window of the x-LEFT ghost rows, i.e. on rank 0 only;
        # the y-left side that carries the code is an open copy.  The kernels test for those cells before any
        # boundary rule, so rank 0 needs nothing but the shell; the other ranks see a plain open side.
        self.jet = any(oc.MHD_JET_BOUNDARY in codes[a] for a in range(self.dim))
        if self.jet and (config.riemann_solver == oc.LAX_FRIEDRICHS or config.limiter == oc.VAN_ALBADA_PP):
            raise NotImplementedError("the MHD jet boundary with LAX_FRIEDRICHS / VAN_ALBADA_PP is not built")
        # the reference's ghost gas cells set dt in this configuration (DESIGN.md "ghost shell"): every rank
        # carries the shell of its own cells (y / z faces of its x range, the x faces on the edge ranks)
        self.shell_mode = config.boundary_handling == oc.GHOST_CELLS and (self.jet or (
            config.mhd and config.split == oc.UNSPLIT
            and any(oc.REFLECTIVE_BOUNDARY in codes[a] for a in range(self.dim))))
        self.dtype, self.device = dtype, device
        self.ng = self.dim + 2
        self.nv = registered_variables.num_vars
        self.mhd = bool(config.mhd)
        self.bcx = oc.boundary_codes(config)[0]
        if config.boundary_handling == oc.PERIODIC_ROLL:
            self.bcx = (oc.PERIODIC_BOUNDARY, oc.PERIODIC_BOUNDARY)
        # kernels that need the global wave bounds of every stage's INPUT state: the Lax-Friedrichs flux
        # (_lax_friedrichs.py:61-66) and the unsplit VAN_ALBADA_PP reconstruction (reconstruction.py:186-193,
        # C = alpha / sum(alpha)); amx_api.cu's Engine makes the same choice
        self.lf = (config.riemann_solver == oc.LAX_FRIEDRICHS
                   or (config.limiter == oc.VAN_ALBADA_PP and config.split == oc.UNSPLIT))
        self.tol = 1e-5 if config.numerical_precision == oc.SINGLE_PRECISION else 1e-10
        self.backend = backend if backend is not None else AmxBackend(config, params, self.local_n, dtype, device,
                                                                     arithmetic)
        if self.jet and self.rank > 0 and hasattr(self.backend, "desc"):
            self.backend.desc.slab_flags |= 1  # same kernels as rank 0, but none of this rank's cells is an inflow cell
        shape = (self.nxl + 2 * NG,) + self.global_n[1:]
        import os
        # Peer-to-peer halo exchange: the slab arrays live in symmetric memory (every rank maps its
        # neighbours' buffers over NVLink), a rank stores its edge planes straight into the
        # neighbours' ghost planes and one device-side barrier closes the exchange.  No staging
        # copies, no NCCL launch; AMX_SLAB_P2P=0 (or a failed rendezvous) keeps NCCL send/recv.
        self._symm = []  # (base tensor, handle)
        self._peer_views = {}
        self.p2p = (self.world > 1 and torch.device(device).type == "cuda"
                    and os.environ.get("AMX_SLAB_P2P", "1") != "0")
        if self.p2p:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self._symm_mem = symm_mem
                self._symm_group = group if group is not None else dist.group.WORLD
            except Exception as e:  # pragma: no cover - depends on the torch build
                print(f"[slab] symmetric memory unavailable ({e!r}); NCCL send/recv halo exchange", flush=True)
                self.p2p = False

        # One symmetric allocation holds every slab array, so that ONE pair of byte deltas (own pointer ->
        # the same element in the lower / upper neighbour's copy) serves all of them: the kernels that write an
        # array also store its edge planes into the neighbours' ghost planes (amx_desc::peer_*), and a halo
        # exchange of such an array is only a barrier.  A second, tiny symmetric buffer carries the MAX
        # all-reduces / barriers (amx_peer_allreduce_max).
        rows_of = [self.nv, self.nv, self.ng, self.ng] + ([6, 6, self.ng] if self.mhd else [])
        plane = 1
        for v in shape:
            plane *= int(v)
        self._arena = self._arena_h = None
        self._comm = None
        self.push = 0          # amx_peer_push_supported bits of the backend
        self._pushed = {}      # data_ptr of an array whose producer stored its halos -> sync count at that time
        self._sync_seq = 0     # collectives (all-reduces, barriers) enqueued so far
        if self.p2p:
            try:
                total = sum(rows_of) * plane
                self._arena = self._symm_mem.empty((total,), dtype=dtype, device=device)
                self._arena.zero_()
                self._arena_h = self._symm_mem.rendezvous(self._arena, self._symm_group)
                self._symm.append((self._arena, self._arena_h))
            except Exception as e:  # pragma: no cover
                print(f"[slab] symmetric-memory rendezvous failed ({e!r}); NCCL send/recv halo exchange", flush=True)
                self.p2p = False
                self._arena = None
        arrays, off = [], 0
        for rows in rows_of:
            if self._arena is not None:
                arrays.append(self._arena[off:off + rows * plane].view((rows,) + shape))
                off += rows * plane
            else:
                arrays.append(torch.zeros((rows,) + shape, dtype=dtype, device=device))
        self.S = [arrays[0], arrays[1]]
        self.A, self.G = arrays[2], arrays[3]
        self.X = [arrays[4], arrays[5]] if self.mhd else None
        self.Gf = arrays[6] if self.mhd else None  # gas after the magnetic write-back
        if self.p2p and os.environ.get("AMX_SLAB_PUSH", "1") != "0" and hasattr(self.backend, "set_peers"):
            try:
                h = self._arena_h
                left, right = (self.rank - 1) % self.world, (self.rank + 1) % self.world
                periodic = self.bcx[0] == oc.PERIODIC_BOUNDARY
                flags = (1 if (periodic or self.rank > 0) else 0) | (2 if (periodic or self.rank < self.world - 1) else 0)
                ptrs = [int(q) for q in h.buffer_ptrs]
                nbytes = self.backend.peer_comm_bytes()
                self._comm = self._symm_mem.empty((max(nbytes, 2048),), dtype=torch.uint8, device=device)
                self._comm.zero_()
                hc = self._symm_mem.rendezvous(self._comm, self._symm_group)
                self._comm_h = hc
                self._comm_ptrs = (C.c_void_p * self.world)(*[int(q) for q in hc.buffer_ptrs])
                torch.cuda.synchronize(device)
                dist.barrier(group=group)  # every rank's communication buffer is zeroed before anybody signals
                self.push = self.backend.set_peers(flags, ptrs[left] - ptrs[self.rank], ptrs[right] - ptrs[self.rank])
                self._push_flags = flags
                # every rank must make the same sequence of collectives: agree on what the kernels push
                agreed = torch.tensor([(self.push >> k) & 1 for k in range(3)], dtype=torch.int32, device=device)
                dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=group)
                self.push = sum(int(v) << k for k, v in enumerate(agreed.tolist()))
            except Exception as e:  # pragma: no cover
                print(f"[slab] peer-memory plumbing unavailable ({e!r}); copies + NCCL all-reduce", flush=True)
                self._comm = None
                self.push = 0
        self.forcing = None  # forcing.SlabForcing: turbulent driving between dt and the step
        self.shell = None
        if self.shell_mode:
            if not hasattr(self.backend, "shell_update"):
                raise NotImplementedError("this backend does not carry the ghost shell (unsplit MHD with a reflective side)")
            self.shell = torch.zeros((self.ng, self.nxl + 2 * NG) + tuple(v + 2 * NG for v in self.global_n[1:]),
                                     dtype=dtype, device=device)
            self.alpha_shell = torch.zeros(3, dtype=dtype, device=device)
            self.backend.set_shell_sides(self.bcx[0] if self.rank == 0 else None,
                                         self.bcx[1] if self.rank == self.world - 1 else None)
        self.alpha = torch.zeros(3, dtype=dtype, device=device)
        self.change = torch.zeros(2, dtype=dtype, device=device)
        self.change4 = torch.zeros(4, dtype=dtype, device=device)
        self.dt = torch.zeros((), dtype=dtype, device=device)
        self.time = torch.zeros((), dtype=dtype, device=device)
        self.collectives = 0
        self._halo = {}
        self.expect_evals = 3 if dtype == torch.float32 else 7  # fixed-point evaluations expected per step
        import os
        # AMX_SLAB_PACK=0: per-row messages without staging copies (A/B switch, see exchange)
        self.pack_halos = os.environ.get("AMX_SLAB_PACK", "1") != "0"
        self._verify_push = os.environ.get("AMX_SLAB_PUSH_VERIFY", "0") == "1"
        # two fixed-point evaluations per launch where the kernel supports the problem (AMX_SLAB_PAIRS=0: off)
        self.pair_greedy = False  # tests: pair launches regardless of the expected count (exercises the re-run)
        # The pair kernel computes the intermediate iterate on ghost planes -1 / n with the stencil; that is
        # only what the reference has there when x is periodic (the while-body refills the ghosts of every
        # iterate with _boundary_handler, _magnetic_field_update.py:124-125,161-163): physical x sides keep
        # one evaluation per launch.
        self.pairs = (self.mhd and os.environ.get("AMX_SLAB_PAIRS", "1") != "0"
                      and self.bcx == (oc.PERIODIC_BOUNDARY, oc.PERIODIC_BOUNDARY)
                      and bool(getattr(self.backend, "pair_supported", lambda: False)()))

    # ------------------------------------------------------------------ communication
    def _halo_buffers(self, t):
        """(send left, send right, recv right, recv left) staging buffers for arrays of t's row count"""
        key = (t.shape[0], t.dtype)
        if key not in self._halo:
            shape = (t.shape[0], NG) + tuple(t.shape[2:])
            self._halo[key] = [torch.empty(shape, dtype=t.dtype, device=t.device) for _ in range(4)]
        return self._halo[key]

    def _exchange_p2p(self, t, left, right, send_l, send_r):
        """Store this rank's edge planes of ``t`` (a row slice of a symmetric-memory array) into the
        neighbours' ghost planes and close with a barrier over the group.  Write-after-read on the
        ghost planes is safe without a leading barrier: between a kernel that reads an array's
        ghosts and the next exchange of that array there is always another exchange or all-reduce,
        which the writer cannot pass before the reader has reached it."""
        n = self.nxl
        for base, h in self._symm:
            off = t.data_ptr() - base.data_ptr()
            if 0 <= off < base.numel() * base.element_size():
                break
        else:
            return False
        off_e = off // base.element_size()

        def peer(rank):
            key = (base.data_ptr(), rank)
            if key not in self._peer_views:
                self._peer_views[key] = h.get_buffer(rank, tuple(base.shape), base.dtype)
            return self._peer_views[key][off_e:off_e + t.numel()].view(t.shape)

        if send_l:   # my first active planes are the left neighbour's right ghost planes
            peer(left)[:, n + NG:].copy_(t[:, NG:2 * NG])
        if send_r:   # my last active planes are the right neighbour's left ghost planes
            peer(right)[:, :NG].copy_(t[:, n:n + NG])
        h.barrier(channel=0)
        self.collectives += 1
        self._sync_seq += 1
        return True

    def _check_pushed(self, t, left, right, send_l, send_r):
        """AMX_SLAB_PUSH_VERIFY=1 (diagnostic): the ghost planes the neighbours' kernels stored here must equal
        the neighbours' edge planes read back through the peer mapping."""
        self._barrier()
        base, h = self._symm[0]
        off_e = (t.data_ptr() - base.data_ptr()) // base.element_size()
        n = self.nxl

        def peer(rank):
            return h.get_buffer(rank, tuple(base.shape), base.dtype)[off_e:off_e + t.numel()].view(t.shape)

        bad = []
        if send_l and not torch.equal(t[:, :NG], peer(left)[:, n:n + NG]):
            bad.append(("left ghosts", float((t[:, :NG] - peer(left)[:, n:n + NG]).abs().max())))
        if send_r and not torch.equal(t[:, n + NG:], peer(right)[:, NG:2 * NG]):
            bad.append(("right ghosts", float((t[:, n + NG:] - peer(right)[:, NG:2 * NG]).abs().max())))
        name = {self.A.data_ptr(): "A", self.G.data_ptr(): "G"}.get(t.data_ptr(), f"rows={t.shape[0]}")
        print(f"[slab push verify] rank {self.rank} array {name} rows {t.shape[0]} seq {self._sync_seq}: "
              f"{'OK' if not bad else bad}", flush=True)
        self._barrier()

    def _allreduce_max(self, t):
        if self.world > 1:
            if self._comm is not None:  # one kernel through peer memory instead of an NCCL launch
                self.backend.peer_allreduce_max(self._comm_ptrs, self.rank, self.world, t, t.numel())
            else:
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self.collectives += 1
            self._sync_seq += 1

    def _barrier(self):
        self.backend.peer_allreduce_max(self._comm_ptrs, self.rank, self.world, None, 0)
        self.collectives += 1
        self._sync_seq += 1

    def _wrote(self, t, bit):
        """``t`` was just written by a kernel of push class ``bit`` (1 stage, 2 iterate, 4 write-back): if that
        kernel stores halos itself, remember it so that exchange(t) shrinks to a barrier (or to nothing when a
        collective has passed since).  Safe against write-after-read: between any kernel that reads the ghost
        planes of an array and the next kernel that writes the array there is a collective on every rank."""
        if self.push & bit:
            self._pushed[t.data_ptr()] = self._sync_seq

    def exchange(self, t, flip_rows=()):
        """Refresh the x ghost planes of ``t`` ([rows][nxl+4][...]).  flip_rows: rows negated in
        the ghost planes of a reflective physical x boundary (gas state: (1,) = v_x; 3-vectors
        and the velocity rows seen by the magnetic step: the reference's Q12 rule)."""
        n = self.nxl
        left, right = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        periodic = self.bcx[0] == oc.PERIODIC_BOUNDARY
        at_left, at_right = self.rank == 0, self.rank == self.world - 1
        if self.world == 1:
            if periodic:
                t[:, :NG].copy_(t[:, n:n + NG])
                t[:, n + NG:].copy_(t[:, NG:2 * NG])
        else:
            send_l = periodic or not at_left
            send_r = periodic or not at_right
            seq = self._pushed.pop(t.data_ptr(), None)
            if seq is not None:
                # the producer stored the edge planes into the neighbours' ghost planes while it ran; they are
                # complete once every rank has passed a collective enqueued after its producer
                if seq == self._sync_seq:
                    self._barrier()
                if self._verify_push:
                    self._check_pushed(t, left, right, send_l, send_r)
                send_l = send_r = False
            elif self.p2p and self._exchange_p2p(t, left, right, send_l, send_r):
                send_l = send_r = False  # done
            # one packed buffer per direction (a single strided copy each) instead of one message
            # per row: 4 NCCL point-to-point operations per exchange, grouped into one launch.
            # The k-th send to a peer pairs with the k-th receive from it (with 2 ranks both
            # neighbours are the same peer): sends are posted as (to left, to right), receives as
            # (from right, from left).
            ops = []
            if not self.pack_halos:
                # one message per row straight from / into the strided array (each row's two planes
                # are contiguous): no staging copies on the device, 4 x rows operations in the group
                for r in range(t.shape[0]):
                    if send_l:
                        ops.append(dist.P2POp(dist.isend, t[r, NG:2 * NG], left, self.group))
                    if send_r:
                        ops.append(dist.P2POp(dist.isend, t[r, n:n + NG], right, self.group))
                for r in range(t.shape[0]):
                    if send_r:
                        ops.append(dist.P2POp(dist.irecv, t[r, n + NG:], right, self.group))
                    if send_l:
                        ops.append(dist.P2POp(dist.irecv, t[r, :NG], left, self.group))
                if ops:
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
                    self.collectives += 1
                send_l = send_r = False  # nothing staged
            buf = self._halo_buffers(t) if (send_l or send_r) else None
            ops = []
            if send_l:
                buf[0].copy_(t[:, NG:2 * NG])
                ops.append(dist.P2POp(dist.isend, buf[0], left, self.group))
            if send_r:
                buf[1].copy_(t[:, n:n + NG])
                ops.append(dist.P2POp(dist.isend, buf[1], right, self.group))
            if send_r:
                ops.append(dist.P2POp(dist.irecv, buf[2], right, self.group))
            if send_l:
                ops.append(dist.P2POp(dist.irecv, buf[3], left, self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                self.collectives += 1
            if send_r:
                t[:, n + NG:].copy_(buf[2])
            if send_l:
                t[:, :NG].copy_(buf[3])
        if periodic:
            return
        # physical (non-periodic) x boundaries are filled locally on the edge ranks
        # (_geometry/boundaries.py:270-355: open copies the first/last cell, reflective mirrors)
        for side, code, here in ((0, self.bcx[0], at_left), (1, self.bcx[1], at_right)):
            if not here:
                continue
            for i in range(NG):
                if side == 0:
                    dst = NG - 1 - i
                    src = NG if code == oc.OPEN_BOUNDARY else NG + i
                else:
                    dst = n + NG + i
                    src = n + NG - 1 if code == oc.OPEN_BOUNDARY else n + NG - 1 - i
                t[:, dst].copy_(t[:, src])
                if code == oc.REFLECTIVE_BOUNDARY:
                    for r in flip_rows:
                        if r < t.shape[0]:
                            t[r, dst].neg_()

    # ------------------------------------------------------------------ state in / out
    def scatter(self, global_state):
        """Global unpadded state (host or device tensor, identical on every rank) -> local slab."""
        x0 = self.rank * self.nxl
        loc = torch.as_tensor(global_state)[:, x0:x0 + self.nxl].to(self.device, self.dtype)
        self.S[0][:, NG:NG + self.nxl].copy_(loc)
        self.reset()
        self.time.zero_()
        self._init_shell(self.S[0][: self.ng])
        self._alpha_of(self.S[0][: self.ng])

    def reset(self):
        """S[0] holds a new initial state (active planes): restart the time loop from it."""
        self.cur = 0
        self._pushed.clear()  # nobody stored the halos of the new state

    def local_result(self):
        return self.S[self.cur][:, NG:NG + self.nxl]

    def gather(self):
        loc = self.local_result().contiguous()
        if self.world == 1:
            return loc.clone()
        parts = [torch.empty_like(loc) for _ in range(self.world)]
        dist.all_gather(parts, loc, group=self.group)
        return torch.cat(parts, dim=1)

    # ------------------------------------------------------------------ step pieces
    def _alpha_of(self, gas, with_shell=True):
        self.backend.alpha(gas, self.alpha)
        if self.shell is not None and with_shell:
            self.backend.shell_update(2, self.shell, alpha_out=self.alpha_shell)
            torch.maximum(self.alpha, self.alpha_shell, out=self.alpha)
        self._allreduce_max(self.alpha)

    def _init_shell(self, gas):
        """time_integration.py:274-286: the padded initial state's ghost gas cells are images of interior cells"""
        if self.shell is not None:
            self.backend.shell_update(3, self.shell, u1=gas)  # (source cells are this rank's own active cells)

    def _stage_alpha(self, want):
        """alpha_out argument + all-reduce for kernels that also produce wave bounds"""
        return self.alpha if want else None

    def _hydro(self, gas_in, gas_out, scale, final_alpha, tmpB):
        b = self.backend
        cfg = self.config
        if cfg.split == oc.UNSPLIT:
            if self.shell is not None and self.lf:
                # Lax-Friedrichs / PP stages want the bounds of the ghost-FILLED state, i.e. of the interior
                # (self.alpha holds those of the padded array, which dt needed)
                self._alpha_of(gas_in, with_shell=False)
            self.exchange(gas_in, flip_rows=(1,))
            b.hydro_stage(gas_in, None, self.A, self.dt, scale, self.alpha, self.alpha if self.lf else None)
            self._wrote(self.A, 1)
            if self.lf:
                self._allreduce_max(self.alpha)
            self.exchange(self.A, flip_rows=(1,))
            want = final_alpha or self.lf
            b.hydro_stage(self.A, gas_in, gas_out, self.dt, scale, self.alpha, self.alpha if want else None)
            self._wrote(gas_out, 1)
            if self.shell is not None:
                # evolve_state.py:456-469 at the ghost cells; the bounds the next CFL reads are those of the padded array
                b.shell_update(0, self.shell, u0=gas_in, u1=gas_out, alpha_out=self.alpha_shell if final_alpha else None)
                if final_alpha:
                    torch.maximum(self.alpha, self.alpha_shell, out=self.alpha)
            if want:
                self._allreduce_max(self.alpha)
            return
        seq = {2: [(1, .5), (2, 1.), (1, .5)], 3: [(1, .5), (2, .5), (3, 1.), (2, .5), (1, .5)]}[self.dim]
        src = gas_in
        for k, (axis, frac) in enumerate(seq):
            last = k == len(seq) - 1
            dst = gas_out if last else (tmpB if k & 1 else self.A)
            self.exchange(src, flip_rows=(1,))
            want = (last and final_alpha) or self.lf
            b.split_sweep(src, dst, axis, self.dt, scale * frac, self.alpha, self.alpha if want else None)
            if last and self.shell is not None:
                # a sweep ends with a ghost fill: the shell is the fill of the result (images / jet constants)
                b.shell_update(3, self.shell, u1=dst, alpha_out=self.alpha_shell if final_alpha else None)
                if final_alpha:
                    torch.maximum(self.alpha, self.alpha_shell, out=self.alpha)
            if want:
                self._allreduce_max(self.alpha)
            src = dst

    def _magnetic(self, B0, gas, B_out):
        """_magnetic_field_update.py:37-242; returns the number of while-body iterations.

        The loop state stays on the device: every launch is followed by the all-reduce of its two
        change norms and the device-side while-condition, launches past the end of the loop do
        nothing.  As many launches as the previous step needed (+1) are enqueued without a host
        round trip; the host looks once per step.  The iteration counts derive from all-reduced
        values, so every rank enqueues the same sequence of collectives."""
        b = self.backend
        # Q12: on a reflective x boundary the reference flips component 1 (y) of 3-vectors
        self.exchange(B0, flip_rows=(1,))
        self.exchange(gas, flip_rows=(2,))
        nl, c = 0, 0  # launches / evaluations enqueued; launch L writes iterate buffer X[L & 1]

        def launch(pair):
            nonlocal nl, c
            src = None if nl == 0 else self.X[(nl - 1) & 1]
            dst = self.X[nl & 1]
            if src is not None:
                self.exchange(src, flip_rows=(1, 4))
            if pair:
                # two evaluations, the intermediate iterate stays on chip (amx_pair.cu); if the loop
                # ends on the first of them, that one is re-run alone into the same buffer.  The
                # launch expected to end the loop also writes the write-back of its result
                # (gas -> Gf, B -> B_out), which makes magnetic_final a no-op when it does end there.
                emit = self.pairs and not self.lf and self.shell is None and c + 2 == self.expect_evals
                b.magnetic_pair_guarded(B0, gas, src, dst, self.dt, self.change4, c,
                                        self.Gf if emit else None, B_out if emit else None)
                self._wrote(dst, 2)
                if emit:
                    self._wrote(self.Gf, 2)
                    self._wrote(B_out, 2)
                self._allreduce_max(self.change4)
                b.magnetic_check_pair(self.change4, c, nl & 1, emit)
                b.magnetic_fix_guarded(B0, gas, src, dst, self.dt, c)
                c += 2
            else:
                b.magnetic_iter_guarded(B0, gas, src, dst, self.dt, self.change, nl)
                self._wrote(dst, 2)
                self._allreduce_max(self.change)
                b.magnetic_check(self.change, nl)
                c += 1
            nl += 1

        pairs = self.pairs
        while c < self.expect_evals:
            launch(pairs and (self.expect_evals - c >= 2 or self.pair_greedy))
        launch(pairs)  # slack
        active, it = b.magnetic_state()
        while active:
            launch(pairs)
            if not pairs:
                launch(False)
            active, it = b.magnetic_state()
        self.expect_evals = it + 1
        b.magnetic_final_guarded(gas, self.Gf, self.X[0], self.X[1], B_out, self.alpha if self.lf else None)
        self._wrote(self.Gf, 4)
        self._wrote(B_out, 4)
        if self.shell is not None and self.config.split == oc.UNSPLIT:
            # _magnetic_field_update.py:208-240 at the ghost cells (Q12 flip rule)
            b.shell_update(1, self.shell, u0=gas, bv0=self.X[0], bv1=self.X[1])
        if self.lf:
            self._allreduce_max(self.alpha)
        return it

    def step(self):
        """dt from the all-reduced wave bounds, one _evolve_state_fv, t += dt."""
        self.backend.dt_from_alpha(self.alpha, self.time if self.config.exact_end_time else None,
                                   self.params.t_end, self.dt)
        sin, sout = self.S[self.cur], self.S[self.cur ^ 1]
        if self.forcing is not None:
            # time_integration.py:699-711: the driving acts on the state between dt and the step; the ghost
            # planes the previous step's kernels stored are those of the unforced state
            self.forcing.apply(sin[:, NG:NG + self.nxl], self.dt)
            self._pushed.pop(sin.data_ptr(), None)
        iters = 0
        if not self.mhd:
            self._hydro(sin, sout, 1.0, True, self.G)
        else:
            gin, gout = sin[: self.ng], sout[: self.ng]
            self._hydro(gin, self.G, 0.5, False, self.X[0][: self.ng])
            iters = self._magnetic(sin[self.ng:], self.G, sout[self.ng:])
            self._hydro(self.Gf, gout, 0.5, True, self.X[0][: self.ng])
        self.time.add_(self.dt)
        self.cur ^= 1
        return iters

    def run(self, max_steps=None):
        """while t < t_end (time_integration.py:779-801); returns StepInfo."""
        steps, iters = 0, 0
        t_end = float(torch.tensor(self.params.t_end, dtype=self.dtype).item())
        while (max_steps is None or steps < max_steps) and float(self.time.item()) < t_end:
            iters += self.step()
            steps += 1
        return StepInfo(steps, float(self.time.item()), iters, getattr(self.backend, "launches", 0))


# ---------------------------------------------------------------------------------------- bench.py, N > 1
#: north-star weak-scaling series (SURVEY.md 8d, C5 geometry): 512^3 cells per GPU, slabs along x;
#: world size -> global grid.  8 GPUs = the 1024^3 domain as 128 x 1024^2 slabs.
WEAK_SERIES = {1: (512, 512, 512), 2: (1024, 512, 512), 4: (1024, 1024, 512), 8: (1024, 1024, 1024)}


def fill_blast_lattice(drv, cell_n, p_in=1000.0, p_out=0.1, B0=None, r_inj=0.1):
    """Initial state of the local slab, built on the device: the 3-D MHD blast of tests/mhd_tests/mhd_blast3D.py
    repeated with period 1 in every direction on a grid of spacing 1 / cell_n (a cubic cell_n^3 grid gives
    exactly problems.mhd_blast3d): every 512^3 block of the series holds the same problem, so dt and the
    fixed-point iteration counts equal the single-GPU run's."""
    import math

    import numpy as np
    B0 = 100.0 / math.sqrt(4.0 * math.pi) if B0 is None else B0
    dev, n = drv.device, drv.nxl
    dx = 1.0 / cell_n

    def d2(i0, m):
        c = (torch.arange(i0, i0 + m, device=dev, dtype=torch.float64) + 0.5) * dx
        c = torch.remainder(c, 1.0) - 0.5
        return c * c

    gy, gz = drv.global_n[1], drv.global_n[2]
    r2 = d2(drv.rank * n, n)[:, None, None] + d2(0, gy)[None, :, None] + d2(0, gz)[None, None, :]
    act = drv.S[0][:, NG:NG + n]
    act[0].fill_(1.0)
    act[1:4].zero_()
    act[4].copy_(torch.where(r2 < r_inj ** 2, p_in, p_out))
    del r2
    theta, phi = np.pi / 2, np.pi / 4
    act[5].fill_(float(B0 * np.sin(theta) * np.cos(phi)))
    act[6].fill_(float(B0 * np.sin(theta) * np.sin(phi)))
    act[7].fill_(float(B0 * np.cos(theta)))
    drv.reset()
    drv.time.zero_()
    drv._alpha_of(drv.S[0][: drv.ng])


def _timed_steps(drv, K, W, device, sample_clocks):
    import time as _time
    world, rank = dist.get_world_size(), dist.get_rank()
    for _ in range(W):
        drv.step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = None
    if sample_clocks and rank == 0:
        from bench import ClockSampler  # the driver script is on sys.path when this runs
        sampler = ClockSampler(device.index or 0)
        sampler.start()
        _time.sleep(0.3)
    dist.barrier()
    torch.cuda.synchronize()
    l0, c0 = drv.backend.launches, drv.collectives
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = _time.time()
    ev0.record()
    iters = 0
    for _ in range(K):
        iters += drv.step()
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t1 = _time.time()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # device time, max over ranks
    clocks = sampler.stop(t0, t1) if sampler is not None else None
    return float(ms.item()), iters, drv.backend.launches - l0, drv.collectives - c0, clocks


def _slab_equals_single(device, n=256, steps=2):
    """FAST-arithmetic check that rides with every multi-GPU bench run: the n^3 blast split into slabs gives
    the same bits as amx_integrate on one GPU (max is the only reduction on the path)."""
    import numpy as np

    from . import problems as pr
    from . import stepping as ti
    rank = dist.get_rank()
    config, params, rv = pr.make(3, n, mhd=True)
    params = params._replace(t_end=1.0e9)
    drv = SlabDriver(config, params, rv, (n, n, n), torch.float32, device)
    fill_blast_lattice(drv, n)
    info = drv.run(max_steps=steps)
    out = drv.gather()
    same = True
    if rank == 0:
        ref = ti.integrate_steps(torch.from_numpy(pr.mhd_blast3d(n, np.float32)).to(device), config, params, rv, steps)
        same = bool(torch.equal(out, ref)) and info.magnetic_iterations == ti.last_info.magnetic_iterations
    flag = torch.tensor([1 if same else 0], device=device)
    dist.broadcast(flag, 0)
    del drv
    torch.cuda.empty_cache()
    return bool(flag.item())


def bench_weak(n, K, W, device):
    """bench.py's N > 1 arm.  Headline: the north-star weak series, 512^3 cells per GPU (WEAK_SERIES; 8 GPUs =
    1024^3 as 128 x 1024^2 slabs).  Secondary: n^3 cells per rank (global n*world x n x n, round 1's
    geometry).  Both: 3-D MHD blast lattice, fp32, slabs along x."""
    import json
    import os

    from . import problems as pr

    world, rank = dist.get_world_size(), dist.get_rank()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    same = _slab_equals_single(device)

    def run(cell_n, global_n, with_e2e, sample_clocks):
        config, params, rv = pr.make(3, cell_n, mhd=True)
        params = params._replace(t_end=1.0e9)
        drv = SlabDriver(config, params, rv, global_n, torch.float32, device)
        fill_blast_lattice(drv, cell_n)
        ms, iters, launches, colls, clocks = _timed_steps(drv, K, W, device, sample_clocks)
        cells = global_n[0] * global_n[1] * global_n[2]
        res = {"ms": ms, "iters": iters, "launches": launches, "colls": colls, "clocks": clocks, "cells": cells,
               "value": cells * K / (ms * 1e-3), "push": drv.push, "p2p": drv.p2p,
               "local": (drv.nxl,) + tuple(global_n[1:])}
        if with_e2e:
            # end to end: pinned host slab -> device -> K steps -> the same pinned buffer
            n_loc = drv.nxl
            host = torch.empty((drv.nv, n_loc) + tuple(global_n[1:]), dtype=torch.float32, pin_memory=True)
            fill_blast_lattice(drv, cell_n)
            host.copy_(drv.S[0][:, NG:NG + n_loc])
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            drv.S[0][:, NG:NG + n_loc].copy_(host, non_blocking=True)
            drv.reset()
            drv.time.zero_()
            drv._alpha_of(drv.S[0][: drv.ng])
            for _ in range(K):
                drv.step()
            host.copy_(drv.local_result(), non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res["e2e_ms"] = float(t.item())
            res["host_bytes"] = host.numel() * 4
            del host
        del drv
        torch.cuda.empty_cache()
        return res

    def run_c5(nc, steps, warm):
        """BASELINE config 5: driven MHD turbulence on the nc^3 cube, slabs along x (rho 0.5, p 0.1, v 0,
        B = (0.1, 0.1, 0) / sqrt 2, energy injection rate 2, PERIODIC_ROLL; arena_tests/turbulent_forcing)."""
        from . import option_classes as oc
        from .forcing import SlabForcing
        tfc = oc.TurbulentForcingConfig(turbulent_forcing=True)
        config, params, rv = pr.make(3, nc, mhd=True, boundary_handling=oc.PERIODIC_ROLL, turbulent_forcing_config=tfc)
        params = params._replace(t_end=1.0e9, dt_max=0.1,
                                 turbulent_forcing_params=params.turbulent_forcing_params._replace(energy_injection_rate=2.0))
        drv = SlabDriver(config, params, rv, (nc, nc, nc), torch.float32, device)
        act = drv.S[0][:, NG:NG + drv.nxl]
        act.zero_()
        act[0].fill_(0.5)
        act[4].fill_(0.1)
        act[5].fill_(0.1 / 2 ** 0.5)
        act[6].fill_(0.1 / 2 ** 0.5)
        drv.reset()
        drv.time.zero_()
        drv._alpha_of(drv.S[0][: drv.ng])
        drv.forcing = SlabForcing(config, params, rv, torch.float32, device, drv.rank, drv.world)
        ms, iters, launches, colls, _ = _timed_steps(drv, steps, warm, device, False)
        vel = drv.local_result()[1:4]
        ek = torch.stack([(0.5 * 0.5 * vel.double() ** 2).sum()])
        dist.all_reduce(ek)
        finite = bool(torch.isfinite(drv.local_result()).all())
        res = {"workload": f"C5: 3-D driven MHD turbulence {nc}^3 on {world} GPUs (x slabs), fp32, HLL, minmod, "
                           "PERIODIC_ROLL; per step: Threefry noise, spectrum + projection, distributed inverse FFT "
                           "(cuFFT + NCCL all-to-all), amplitude from two all-reduced sums, then the MHD step",
               "value": nc ** 3 * steps / (ms * 1e-3), "unit": "cell-updates/s", "steps": steps, "ms_per_step": ms / steps,
               "magnetic_iterations_per_step": iters / steps, "finite": finite,
               "kinetic_energy_after": float(ek.item()) * config.grid_spacing ** 3,
               "time": float(drv.time.item())}
        del drv
        torch.cuda.empty_cache()
        return res

    series = WEAK_SERIES.get(world)
    if series is not None and os.environ.get("AMX_BENCH_SERIES", "1") != "0":
        main = run(512, series, True, True)
        second = run(n, (n * world, n, n), False, False)
        workload = (f"3D MHD blast lattice, 512^3 cells per GPU (north-star weak series: global "
                    f"{series[0]}x{series[1]}x{series[2]}, local slab {main['local'][0]}x{main['local'][1]}x{main['local'][2]}), "
                    "fp32, HLL, minmod, unsplit RK2-SSP, Pang-Wu implicit-midpoint magnetic step, periodic, C_cfl=0.4")
    else:
        main = run(n, (n * world, n, n), True, True)
        second = None
        workload = None
    bytes_cell = (10 * 5 + 13 + 19 * (main["iters"] / K)) * 4
    per_gpu = main["value"] / world
    roofline = {"bound": "hbm", "kernel": "whole step (per-kernel event timing is N=1 only)",
                "achieved": per_gpu * bytes_cell / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": per_gpu * bytes_cell / 1e9 / hbm, "traffic": None,
                "peak_source": "measured" if "hbm_gbs" in peaks else "fallback"}
    if main["push"]:
        halo = ("edge planes stored into the neighbours' ghost planes by the producing kernels themselves (P2P "
                "stores over NVLink, overlapped with the launch); an exchange is a barrier")
        red = "MAX all-reduce of 3 (dt) / 4 (fixed point) scalars in one kernel through peer memory"
    elif main["p2p"]:
        halo = "edge planes copied into the neighbours' ghost planes over NVLink (symmetric memory) + device-side barrier"
        red = "NCCL all-reduce MAX of 3 (dt) + 4 (fixed point) scalars"
    else:
        halo, red = "NCCL send/recv (packed)", "NCCL all-reduce MAX"
    cfg = {"collectives_per_step": main["colls"] / K, "halo": halo, "reductions": red,
           "slab_equals_single_gpu_bitwise_256": same}
    if os.environ.get("AMX_BENCH_C5", "1") != "0" and 1024 % world == 0:
        try:
            cfg["c5_driven_turbulence"] = run_c5(1024 if world == 8 else 512, 5, 2)
        except Exception as e:  # a secondary line must not take the headline down
            cfg["c5_driven_turbulence"] = {"error": repr(e)}
    if second is not None:
        cfg["secondary"] = {"workload": f"{n}^3 cells per rank, global {n * world}x{n}x{n}",
                            "value": second["value"], "ms_per_step": second["ms"] / K,
                            "collectives_per_step": second["colls"] / K,
                            "magnetic_iterations_per_step": second["iters"] / K}
    return {"ms": main["ms"], "e2e_ms": main["e2e_ms"], "h2d_bytes": main["host_bytes"],
            "d2h_bytes": main["host_bytes"], "launches": main["launches"], "clocks": main["clocks"],
            "roofline": roofline, "kernels": {}, "mag_iters": main["iters"], "cells": main["cells"],
            "workload": workload, "config": cfg}
