"""D3Q19Engine: device state + calls into liblbm_b200.so.

PyTorch is plumbing here (device memory, streams, torch.distributed rendezvous); every
kernel is hand-written CUDA behind the C ABI.  There is no fallback path: constructing an
engine without the shared library or without an sm_100 device raises.

HBM layout (see include/lbm_b200.h): x fastest, z slowest, SoA.
    g          [19, nzp, ny, nx]  post-collision populations, ping-pong pair
    rho/phase  [nzp, ny, nx]
    u, force   [3, nzp, ny, nx]
    flags      [nzp, ny, nx] u8 (solid | filter | les | near)
nzp = nz + 2*zghost.  The reference's logical index order ([q,i,j,k], [i,j,k,c]) is exposed by
`fields.py` as permuted views over these buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .config import LBMConfig
from .errors import BackendInitializationError, ComputeExecutionError


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class D3Q19Engine:
    def __init__(self, nx: int, ny: int, nz: int, *, compat: str = "physical",
                 periodic: Sequence[bool] = (True, True, True), walls: bool = False, force: bool = False,
                 phase: bool = False, les: bool = False, porous: bool = False, strict: bool = True,
                 config: Optional[LBMConfig] = None, device: int = 0, zghost: int = 0, z0: int = 0,
                 nz_global: Optional[int] = None, tau: Optional[float] = None, tau_air: Optional[float] = None,
                 gravity_lu: Optional[float] = None, cs_smag: Optional[float] = None,
                 porous_darcy: float = 0.0, porous_forch: float = 0.0, vec: int = 0, block: int = 0,
                 macro_fields: bool = True, drive: bool = False, drive_max_force: float = 0.12, drive_scale: float = 1.0,
                 mrt_magic: float = 0.0):
        if not torch.cuda.is_available():
            raise BackendInitializationError(
                "no CUDA device visible: pour_over_coffee_lbm_b200 runs on B200 (sm_100a) only, there is no CPU fallback",
                "b200", "NO_DEVICE")
        self.lib = L.lib()
        self.cfg = config or LBMConfig(NX=nx, NY=ny, NZ=nz_global or nz)
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.zghost, self.z0 = int(zghost), int(z0)
        self.nz_global = int(nz_global or nz)
        self.nzp = self.nz + 2 * self.zghost
        self.compat = L.COMPAT_REFERENCE if compat == "reference" else L.COMPAT_PHYSICAL
        self.compat_name = "reference" if self.compat == L.COMPAT_REFERENCE else "physical"
        self.device = torch.device("cuda", device)
        self.device_index = device
        self.periodic = tuple(bool(b) for b in periodic)
        feats = 0
        if walls: feats |= L.FEAT_WALLS
        if force: feats |= L.FEAT_FORCE
        if phase: feats |= L.FEAT_PHASE
        if les: feats |= L.FEAT_LES
        if porous: feats |= L.FEAT_POROUS
        if strict: feats |= L.FEAT_STRICT
        if drive: feats |= L.FEAT_DRIVE          # pressure-gradient drive fused into the step kernel (include/lbm_b200.h)
        self.drive = bool(drive)
        self.features = feats
        k_lu, beta_lu = self.cfg.forchheimer_parameters()
        c_darcy, c_forch = self.cfg.filter_constants()
        self.params = L.LbmParams(
            nx=self.nx, ny=self.ny, nz=self.nz, nz_global=self.nz_global, z0=self.z0, zghost=self.zghost,
            periodic=(1 if self.periodic[0] else 0) | (2 if self.periodic[1] else 0) | (4 if self.periodic[2] else 0),
            compat=self.compat, features=feats,
            tau_water=self.cfg.TAU_WATER if tau is None else tau,
            tau_air=self.cfg.TAU_AIR if tau_air is None else tau_air,
            gravity_lu=self.cfg.GRAVITY_LU if gravity_lu is None else gravity_lu,
            cs_smag=self.cfg.LES_CS if cs_smag is None else cs_smag, tau_min=0.55, tau_max=1.90,
            porous_darcy=porous_darcy, porous_forch=porous_forch,
            K_lu=k_lu, beta_lu=beta_lu, c_darcy=c_darcy, c_forch=c_forch, vec=vec, block=block,
            drive_max_force=drive_max_force, drive_scale=drive_scale, mrt_magic=mrt_magic)
        self._ctx = C.c_void_p()
        rc = self.lib.lbm_create(C.byref(self._ctx), device, C.byref(self.params))
        if rc != 0:
            raise BackendInitializationError(self.lib.lbm_last_error(None).decode(), "b200", "INIT_FAILED")

        dev = self.device
        shp = (self.nzp, self.ny, self.nx)
        with torch.cuda.device(dev):
            # zeros, not empty: the four-cell walls kernel pulls solid cells' words into lanes whose results nobody reads; NaN there
            # would only cost speed (slow paths of rcp / sqrt), but it would
            self.g = [torch.zeros((L.Q,) + shp, dtype=torch.float32, device=dev) for _ in range(2)]
            self.cur = 0                      # index of the buffer holding the newest populations
            # rho: one buffer; a ping-pong pair when the fused drive reads the previous step's density while this step's is written
            self.rho_buf = [torch.ones(shp, dtype=torch.float32, device=dev) for _ in range(2 if drive else 1)] if macro_fields else []
            self.rho_cur = 0
            self.ref_les = self.compat == L.COMPAT_REFERENCE and les
            nu = 2 if self.ref_les else 1
            self.u_buf = [torch.zeros((3,) + shp, dtype=torch.float32, device=dev) for _ in range(nu)] if macro_fields else []
            self.u_cur = 0
            self.body_force = torch.zeros((3,) + shp, dtype=torch.float32, device=dev) if force else None
            self.phase = torch.zeros(shp, dtype=torch.float32, device=dev) if phase else None
            self.flags = torch.full(shp, L.FLAG_LES, dtype=torch.uint8, device=dev) if walls else None
            self.solid = torch.zeros(shp, dtype=torch.uint8, device=dev) if walls else None
            self.filter_zone = torch.zeros(shp, dtype=torch.int32, device=dev) if walls else None
            self.les_mask = torch.ones(shp, dtype=torch.int32, device=dev) if walls else None
            self.blockage = None
        self.comm_stream = None
        self.rank, self.nranks = 0, 1
        self.steps_done = 0
        self.populations_generation = 0       # bumped by everything that rewrites g (fields.PopulationField keys its cache on it)
        if walls:
            self.pack_flags()          # all-fluid box: NEAR on open faces, work lists for the step kernels
        self.init_equilibrium(1.0, (0.0, 0.0, 0.0))

    # ---------------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise ComputeExecutionError(f"{what}: {self.lib.lbm_last_error(self._ctx).decode()}", "b200", "EXECUTION_FAILED")

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def u(self) -> torch.Tensor:
        return self.u_buf[self.u_cur]

    @property
    def rho(self) -> Optional[torch.Tensor]:
        """the newest density field (None when the engine was built without macroscopic fields)"""
        return self.rho_buf[self.rho_cur] if self.rho_buf else None

    @property
    def populations(self) -> torch.Tensor:
        """post-collision populations g[q, zp, y, x] (newest state)"""
        return self.g[self.cur]

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.lbm_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def populations_changed(self):
        """Call after writing `populations` / `g[...]` directly (torch ops): compat = physical behind walls keeps
        bounce-back copies in the solid cells' slots and must rebuild them (include/lbm_b200.h)."""
        self.populations_generation += 1
        self._check(self.lib.lbm_populations_changed(self._ctx), "lbm_populations_changed")

    def selftest_math(self):
        out = (C.c_ulonglong * 7)()
        self._check(self.lib.lbm_selftest_math(self._ctx, out, self.stream), "lbm_selftest_math")
        return tuple(int(v) for v in out)

    def launch_count(self) -> int:
        return int(self.lib.lbm_launch_count(self._ctx))

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.params, k, v)
        self._check(self.lib.lbm_set_params(self._ctx, C.byref(self.params)), "lbm_set_params")

    # ---- initialisation ------------------------------------------------------------------------
    def init_equilibrium(self, rho0: float = 1.0, u0=(0.0, 0.0, 0.0), rho: Optional[torch.Tensor] = None,
                         u: Optional[torch.Tensor] = None):
        """g <- f_eq(rho, u); both ping-pong buffers (LBMSolver.init_fields sets f and f_new)."""
        arr = (C.c_float * 3)(*[float(x) for x in u0])
        self.populations_generation += 1
        for buf in self.g:
            self._check(self.lib.lbm_init_equilibrium(self._ctx, _ptr(buf), _ptr(rho), _ptr(u), float(rho0), arr, self.stream),
                        "lbm_init_equilibrium")
        if self.rho is not None:
            for rb in self.rho_buf:
                if rho is None: rb.fill_(rho0)
                else: rb.copy_(rho)
            for ub in self.u_buf:
                if u is None:
                    for c in range(3): ub[c].fill_(float(u0[c]))
                else:
                    ub.copy_(u)
        if self.zghost:
            self.halo_exchange()

    def build_v60_geometry(self):
        """FilterPaperSystem.initialize_filter_geometry: solid, filter_zone, les_mask punch-out."""
        geom = (C.c_float * 5)(*self.cfg.v60_geometry_constants())
        self._check(self.lib.lbm_build_v60_geometry(self._ctx, _ptr(self.solid), _ptr(self.filter_zone), geom, self.stream),
                    "lbm_build_v60_geometry")
        self.les_mask[self.filter_zone == 1] = 0          # filter_paper.py:199-204
        self.pack_flags()

    def pack_flags(self):
        self.populations_generation += 1      # the exported f depends on the mask (bounce-back, inflow)
        self._check(self.lib.lbm_pack_flags(self._ctx, _ptr(self.flags), _ptr(self.solid), _ptr(self.filter_zone),
                                            _ptr(self.les_mask), self.stream), "lbm_pack_flags")
        if len(self.u_buf) == 2:      # keep the u ping-pong pair identical on cells the kernel never writes
            self.u_buf[1 - self.u_cur].copy_(self.u_buf[self.u_cur])
        if len(self.rho_buf) == 2:
            self.rho_buf[1 - self.rho_cur].copy_(self.rho_buf[self.rho_cur])

    def set_geometry_preserving_f(self, mutate):
        """Change the solid mask exactly as the reference would see it: the reference streams with the
        mask current at collision time, so convert g -> f with the old flags, mutate, f -> g with the new."""
        f = torch.empty_like(self.g[self.cur])
        self._check(self.lib.lbm_export_f(self._ctx, _ptr(self.g[self.cur]), _ptr(self.flags), _ptr(f), self.stream), "lbm_export_f")
        mutate()
        self.pack_flags()
        self._check(self.lib.lbm_import_f(self._ctx, _ptr(f), _ptr(self.flags), _ptr(self.g[self.cur]), self.stream), "lbm_import_f")
        self.g[1 - self.cur].copy_(self.g[self.cur])
        self.populations_generation += 1
        if self.zghost:
            self.halo_exchange()              # the ghost planes' populations were converted with the old mask

    # ---- the hot path --------------------------------------------------------------------------
    def _fields(self) -> L.LbmFields:
        if self.ref_les:
            u_src, u_dst = self.u_buf[self.u_cur], self.u_buf[1 - self.u_cur]
        else:
            u_src, u_dst = None, (self.u_buf[0] if self.u_buf else None)
        if len(self.rho_buf) == 2:    # fused drive: this step writes the other buffer and reads the newest one
            rho, rho_src = self.rho_buf[1 - self.rho_cur], self.rho_buf[self.rho_cur]
        else:
            rho, rho_src = self.rho, None
        return L.LbmFields(f_src=_ptr(self.g[self.cur]), f_dst=_ptr(self.g[1 - self.cur]), rho=_ptr(rho),
                           u_src=_ptr(u_src), u_dst=_ptr(u_dst), body_force=_ptr(self.body_force),
                           phase=_ptr(self.phase), blockage=_ptr(self.blockage), flags=_ptr(self.flags), rho_src=_ptr(rho_src))

    def step(self, nsteps: int = 1, write_macro_every: int = 1):
        """nsteps fused collide-stream updates (one kernel launch each)."""
        if self.rho is None:
            write_macro_every = 0
        f = self._fields()
        comm = C.c_void_p(self.comm_stream.cuda_stream) if self.comm_stream is not None else None
        self._check(self.lib.lbm_step(self._ctx, C.byref(f), int(nsteps), int(write_macro_every), self.stream, comm), "lbm_step")
        if nsteps % 2 == 1:
            self.cur = 1 - self.cur
        if self.ref_les and write_macro_every == 1 and nsteps % 2 == 1:
            self.u_cur = 1 - self.u_cur
        if len(self.rho_buf) == 2 and nsteps % 2 == 1:
            self.rho_cur = 1 - self.rho_cur
        self.steps_done += nsteps
        self.populations_generation += 1

    def macroscopic(self):
        f = self._fields()
        if self.ref_les:
            f.u_dst = _ptr(self.u_buf[self.u_cur])
        f.rho = _ptr(self.rho)
        self._check(self.lib.lbm_macroscopic(self._ctx, C.byref(f), self.stream), "lbm_macroscopic")

    def face_bc(self):
        f = self._fields()
        f.rho = _ptr(self.rho)
        self._check(self.lib.lbm_face_bc(self._ctx, C.byref(f), self.stream), "lbm_face_bc")

    # ---- restart checkpoint (absent in the reference) ------------------------------------
    _CKPT_FIELDS = ("rho", "body_force", "phase", "solid", "filter_zone", "les_mask", "blockage")

    def save_checkpoint(self, path: str) -> None:
        """Everything a bit-exact restart needs: the current population buffer, the macroscopic fields the kernels read
        back (rho; both u buffers of the reference-mode LES), the inputs (force, phase, geometry) and the step count.
        One file per slab (torch.save); geometry and parameters are checked on load."""
        torch.cuda.synchronize(self.device)
        blob = {"version": 1, "shape": (self.nx, self.ny, self.nz, self.zghost, self.z0, self.nz_global), "compat": self.compat,
                "features": self.features, "steps_done": self.steps_done, "g": self.g[self.cur].cpu(),
                "u": [u.cpu() for u in self.u_buf], "u_cur": self.u_cur}
        for name in self._CKPT_FIELDS:
            t = getattr(self, name, None)
            blob[name] = t.cpu() if t is not None else None
        torch.save(blob, path)

    def load_checkpoint(self, path: str) -> None:
        blob = torch.load(path, map_location="cpu", weights_only=True)      # tensors, ints, lists and tuples only
        if tuple(blob["shape"]) != (self.nx, self.ny, self.nz, self.zghost, self.z0, self.nz_global) or blob["compat"] != self.compat \
                or blob["features"] != self.features:
            raise ValueError("checkpoint was written for a different slab geometry, compat mode or feature set")
        for name in self._CKPT_FIELDS:
            t = getattr(self, name, None)
            if blob[name] is not None and t is None:
                if name != "blockage":
                    raise ValueError(f"checkpoint holds a '{name}' field this engine was built without")
                self.blockage = t = torch.zeros_like(self.rho)        # FilterPaperSystem's blockage, not allocated yet
            if blob[name] is not None:
                t.copy_(blob[name])
        if self.flags is not None:
            self.pack_flags()                      # flags, work lists, neighbour masks from solid / filter_zone / les_mask
        self.g[self.cur].copy_(blob["g"]); self.g[1 - self.cur].copy_(blob["g"])
        for dst, src in zip(self.u_buf, blob["u"]):
            dst.copy_(src)
        self.u_cur = blob["u_cur"] if len(self.u_buf) == 2 else 0
        self.steps_done = int(blob["steps_done"])
        self.populations_changed()
        if self.zghost:
            self.halo_exchange()

    # ---- reference `f` view ------------------------------------------------------------------
    def export_f(self) -> torch.Tensor:
        """The reference's pre-collision f in device layout [19, nzp, ny, nx] (exact data movement)."""
        out = torch.empty_like(self.g[self.cur])
        self._check(self.lib.lbm_export_f(self._ctx, _ptr(self.g[self.cur]), _ptr(self.flags), _ptr(out), self.stream), "lbm_export_f")
        return out

    def import_f(self, f: torch.Tensor):
        f = f.to(self.device, torch.float32).contiguous()
        self.populations_generation += 1
        self._check(self.lib.lbm_import_f(self._ctx, _ptr(f), _ptr(self.flags), _ptr(self.g[self.cur]), self.stream), "lbm_import_f")
        self.g[1 - self.cur].copy_(self.g[self.cur])
        if self.zghost:
            self.halo_exchange()

    # ---- neighbours that feed body_force ----------------------------------------------------
    def clear_body_force(self):
        self.body_force.zero_()

    def add_pressure_gradient_force(self, max_force: float = 0.12, scale: float = 1.0):
        self.exchange_field(scalar=self.rho)      # the gradient reads rho[k -+ 1] across a slab interface
        self._check(self.lib.lbm_pressure_gradient_force(self._ctx, _ptr(self.rho), _ptr(self.flags), _ptr(self.body_force),
                                                         float(max_force), float(scale), self.stream), "lbm_pressure_gradient_force")

    def field_statistics(self) -> torch.Tensor:
        """[max|u|, min rho, max rho, sum rho, kinetic energy, NaN count, Inf count, fluid cells] of the owned slab as an
        8-element f64 DEVICE tensor (one fused deterministic pass, no host sync; include/lbm_b200.h)."""
        if getattr(self, "_stats", None) is None:
            self._stats = torch.zeros(8, dtype=torch.float64, device=self.device)
        self._check(self.lib.lbm_field_statistics(self._ctx, _ptr(self.rho), _ptr(self.u), _ptr(self.flags) if self.flags is not None else None,
                                                  _ptr(self._stats), self.stream), "lbm_field_statistics")
        return self._stats

    def set_pressure_gradient_force(self, max_force: float = 0.12, scale: float = 1.0):
        """clear_body_force() + add_pressure_gradient_force() on the fluid cells in one pass (solid cells keep their
        old body_force, which no kernel reads)."""
        self.exchange_field(scalar=self.rho)
        self._check(self.lib.lbm_pressure_gradient_force_set(self._ctx, _ptr(self.rho), _ptr(self.flags), _ptr(self.body_force),
                                                             float(max_force), float(scale), self.stream), "lbm_pressure_gradient_force_set")

    def add_forchheimer_force(self, fmax: Optional[float] = None):
        fmax = 0.01 * self.cfg.SCALE_VELOCITY / self.cfg.DT if fmax is None else fmax
        self._check(self.lib.lbm_forchheimer_force(self._ctx, _ptr(self.u), _ptr(self.flags), _ptr(self.body_force), float(fmax),
                                                   self.stream), "lbm_forchheimer_force")

    def add_reaction_force(self, reaction: torch.Tensor):
        self._check(self.lib.lbm_add_reaction_force(self._ctx, _ptr(reaction), _ptr(self.flags), _ptr(self.body_force), self.stream),
                    "lbm_add_reaction_force")

    # ---- multiphase / pouring producers (csrc/lbm_producers.cu) -----------------------------------
    def chemical_potential(self, phi, laplacian_phi, mu, kappa: float):
        self._check(self.lib.lbm_chemical_potential(self._ctx, _ptr(phi), _ptr(laplacian_phi), _ptr(mu), float(kappa), self.stream),
                    "lbm_chemical_potential")

    def surface_tension(self, phi, mu, grad_phi, grad_mu, normal, curvature, surface_force, sigma: float, apply: bool = True):
        """compute_gradients + compute_curvature + compute_surface_tension_force (+ apply_surface_tension when `apply`)."""
        self._check(self.lib.lbm_surface_tension(self._ctx, _ptr(phi), _ptr(mu), _ptr(self.rho), _ptr(self.flags), _ptr(grad_phi), _ptr(grad_mu),
                                                 _ptr(normal), _ptr(curvature), _ptr(surface_force),
                                                 _ptr(self.body_force) if apply else None, float(sigma), self.stream), "lbm_surface_tension")

    def surface_tension_gradients(self, phi, mu, grad_phi, grad_mu, normal):
        """First launch of the chain (any zghost): compute_gradients."""
        self._check(self.lib.lbm_surface_tension_gradients(self._ctx, _ptr(phi), _ptr(mu), _ptr(grad_phi), _ptr(grad_mu), _ptr(normal), self.stream),
                    "lbm_surface_tension_gradients")

    def surface_tension_curvature_force(self, phi, grad_phi, normal, curvature, surface_force, sigma: float, apply: bool = True):
        """Second launch of the chain (any zghost; `normal`'s ghost planes must be current): curvature, force, body_force."""
        self._check(self.lib.lbm_surface_tension_curvature_force(self._ctx, _ptr(phi), _ptr(self.rho), _ptr(self.flags), _ptr(grad_phi), _ptr(normal),
                                                                 _ptr(curvature), _ptr(surface_force), _ptr(self.body_force) if apply else None,
                                                                 float(sigma), self.stream), "lbm_surface_tension_curvature_force")

    def surface_tension_body_force(self, phi, sigma: float, normal_outer=None, surface_force_outer=None):
        """body_force += surface tension / rho in ONE launch, no intermediate fields (bit-identical to surface_tension(apply=True))."""
        self._check(self.lib.lbm_surface_tension_body_force(self._ctx, _ptr(phi), _ptr(self.rho), _ptr(self.flags), _ptr(normal_outer),
                                                            _ptr(surface_force_outer), _ptr(self.body_force), float(sigma), self.stream),
                    "lbm_surface_tension_body_force")

    def apply_surface_tension(self, surface_force):
        self._check(self.lib.lbm_apply_surface_tension(self._ctx, _ptr(surface_force), _ptr(self.rho), _ptr(self.flags), _ptr(self.body_force),
                                                       self.stream), "lbm_apply_surface_tension")

    def phase_field_step(self, phi, phi_new, mu, mobility: float, dt: float, rho_water: float, rho_air: float):
        self._check(self.lib.lbm_phase_field_step(self._ctx, _ptr(phi), _ptr(phi_new), _ptr(mu), _ptr(self.u), _ptr(self.rho), _ptr(self.phase),
                                                  float(mobility), float(dt), float(rho_water), float(rho_air), self.stream),
                    "lbm_phase_field_step")

    def density_from_phase(self, phi, rho_water: float, rho_air: float):
        self._check(self.lib.lbm_density_from_phase(self._ctx, _ptr(phi), _ptr(self.rho), _ptr(self.phase), float(rho_water), float(rho_air),
                                                    self.stream), "lbm_density_from_phase")

    def pouring_force(self, pour: "L.LbmPour", body_force=None):
        bf = self.body_force if body_force is None else body_force
        self._check(self.lib.lbm_pouring_force(self._ctx, pour, _ptr(self.flags), _ptr(bf), self.stream), "lbm_pouring_force")

    def pouring_phase_change(self, pour: "L.LbmPour", phi):
        self._check(self.lib.lbm_pouring_phase_change(self._ctx, pour, _ptr(self.flags), _ptr(phi), self.stream), "lbm_pouring_phase_change")

    # ---- slabs ----------------------------------------------------------------------------------
    def attach_process_group(self, group=None):
        """Create the NCCL communicator of the z-slab chain; the unique id travels through
        torch.distributed (any backend)."""
        import torch.distributed as dist
        self.rank, self.nranks = dist.get_rank(group), dist.get_world_size(group)
        uid = (C.c_char * 128)()
        if self.rank == 0:
            self._check(self.lib.lbm_nccl_unique_id(C.cast(uid, C.c_void_p)), "lbm_nccl_unique_id")
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        uid = (C.c_char * 128).from_buffer_copy(box[0])
        with torch.cuda.device(self.device):
            self._check(self.lib.lbm_attach_nccl(self._ctx, C.cast(uid, C.c_void_p), self.rank, self.nranks), "lbm_attach_nccl")
            # high priority: the NCCL send/recv kernel must get SM slots while the interior kernel (tens of
            # thousands of queued CTAs on the compute stream) is still being dispatched, otherwise the halo
            # exchange only starts at the interior kernel's tail and nothing overlaps
            self.comm_stream = torch.cuda.Stream(self.device, priority=-1)

    def halo_exchange(self, with_u: bool = False):
        v = _ptr(self.u) if (with_u and self.u_buf) else None
        self._check(self.lib.lbm_halo_exchange(self._ctx, _ptr(self.g[self.cur]), v, self.stream), "lbm_halo_exchange")
        if self.drive:          # the fused drive reads the previous step's rho across the interface
            for rb in self.rho_buf:
                self.exchange_field(scalar=rb)

    def exchange_field(self, scalar: Optional[torch.Tensor] = None, vec3: Optional[torch.Tensor] = None):
        """Refresh the ghost planes of a scalar and / or 3-vector field from the z neighbours (NCCL inside the library)."""
        if self.zghost:
            self._check(self.lib.lbm_halo_exchange_field(self._ctx, _ptr(scalar), _ptr(vec3), self.stream), "lbm_halo_exchange_field")

    # ---- sizes ---------------------------------------------------------------------------------
    def cells(self) -> int:
        return self.nx * self.ny * self.nz

    def fluid_cells(self) -> int:
        if self.solid is None:
            return self.cells()
        own = self.solid[self.zghost:self.zghost + self.nz]
        return int((own == 0).sum().item())


def v60_fluid_cells_per_plane(cfg, device=0, chord_end_cost: float = 0.0) -> list:
    """Fluid-cell count of every z plane of the global V60 box `cfg` (FilterPaperSystem._setup_v60_geometry,
    filter_paper.py:206-286, evaluated by lbm_build_v60_geometry into a temporary u8 mask: 1 B per cell, no populations).
    Input of slab.partition_z_balanced.

    chord_end_cost > 0 returns the step kernel's COST per plane instead: fluid cells + chord_end_cost x (solid / fluid transitions
    along x).  A chord end costs the walls kernel about as much as 16 fluid cells (its sector partner in the list, the wall links,
    DRAM bursts it uses in part: V60 512^3 runs 3.4 ps per fluid cell above the all-fluid box, 360 000 chord ends).  Planes near the
    tip of the cone hold short chords, so slabs cut by fluid count alone leave the rank at the tip 10 % slower than the mean
    (8 GPUs, 1024^3: 2.03 ms against 1.84 ms for the same number of fluid cells on one GPU)."""
    lib = L.lib()
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    p = L.LbmParams(nx=cfg.NX, ny=cfg.NY, nz=cfg.NZ, nz_global=cfg.NZ, z0=0, zghost=0, periodic=0, compat=L.COMPAT_PHYSICAL,
                    features=L.FEAT_WALLS, tau_water=0.6, tau_air=0.8, gravity_lu=0.0, cs_smag=0.18, tau_min=0.55, tau_max=1.9)
    ctx = C.c_void_p()
    if lib.lbm_create(C.byref(ctx), dev.index or 0, C.byref(p)) != 0:
        raise BackendInitializationError(lib.lbm_last_error(None).decode(), "b200", "INIT_FAILED")
    try:
        with torch.cuda.device(dev):
            solid = torch.zeros((cfg.NZ, cfg.NY, cfg.NX), dtype=torch.uint8, device=dev)
            geom = (C.c_float * 5)(*cfg.v60_geometry_constants())
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if lib.lbm_build_v60_geometry(ctx, _ptr(solid), None, geom, stream) != 0:
                raise ComputeExecutionError(lib.lbm_last_error(ctx).decode(), "b200", "EXECUTION_FAILED")
            counts = (solid == 0).sum(dim=(1, 2))
            if chord_end_cost > 0.0:
                ends = torch.zeros_like(counts)
                for z0 in range(0, cfg.NZ, 64):             # in chunks: the comparison makes a temporary of the chunk's size
                    blk = solid[z0:z0 + 64]
                    ends[z0:z0 + 64] = (blk[:, :, 1:] != blk[:, :, :-1]).sum(dim=(1, 2))
                counts = counts.double() + float(chord_end_cost) * ends.double()
            counts = counts.tolist()
            del solid
    finally:
        lib.lbm_destroy(ctx)
    return counts


class ParticleState:
    """SoA particle arrays on the device (CoffeeParticleSystem fields, coffee_particles.py:30-60)."""

    def __init__(self, n: int, device):
        z3 = lambda: torch.zeros((3, n), dtype=torch.float32, device=device)
        z1 = lambda: torch.zeros((n,), dtype=torch.float32, device=device)
        self.n = n
        self.pos, self.vel = z3(), z3()
        self.radius, self.mass = z1(), z1()
        self.active = torch.zeros((n,), dtype=torch.int32, device=device)
        self.drag_new, self.drag_old, self.drag, self.u_fluid = z3(), z3(), z3(), z3()
        self.reynolds, self.cd = z1(), z1()
        self.cell = torch.zeros((3, n), dtype=torch.int32, device=device)

    def sort_by_cell(self, nx: int, ny: int) -> torch.Tensor:
        """Reorder the particles (every array, in place) by the linear index of their base cell, z slowest.  The coupling kernel runs
        one thread per particle: with neighbours in memory being neighbours in space, a warp's gathers and scatters share sectors and
        `__match_any_sync` finds peers to aggregate (1 M particles in the bed of a 1024^3 box: 1.36 ms in random order).  The
        reference creates its bed layer by layer (coffee_particles.py:220-412), i.e. already z-ordered; call this again every few
        hundred steps of a long run.  Returns the permutation that was applied (new[i] = old[perm[i]])."""
        key = (self.pos[2].clamp(min=0).long() * ny + self.pos[1].clamp(min=0).long()) * nx + self.pos[0].clamp(min=0).long()
        perm = torch.argsort(key, stable=True)
        for name in ("pos", "vel", "drag_new", "drag_old", "drag", "u_fluid", "cell"):
            t = getattr(self, name)
            t.copy_(t[:, perm])
        for name in ("radius", "mass", "active", "reynolds", "cd"):
            t = getattr(self, name)
            t.copy_(t[perm])
        return perm

    def struct(self) -> L.LbmParticles:
        return L.LbmParticles(pos=_ptr(self.pos), vel=_ptr(self.vel), radius=_ptr(self.radius), mass=_ptr(self.mass),
                              active=_ptr(self.active), drag_new=_ptr(self.drag_new), drag_old=_ptr(self.drag_old),
                              drag=_ptr(self.drag), u_fluid=_ptr(self.u_fluid), reynolds=_ptr(self.reynolds),
                              cd=_ptr(self.cd), cell=_ptr(self.cell), n=self.n)


def particles_couple(engine: D3Q19Engine, ps: ParticleState, reaction: torch.Tensor, relax: float = 0.8,
                     water_density: Optional[float] = None, water_viscosity: Optional[float] = None, sparse_clear: bool = False,
                     clear_interface: bool = True):
    """CoffeeParticleSystem.compute_two_way_coupling_forces + apply_under_relaxation (one kernel).  sparse_clear: `reaction` is
    written by this call only (and was zero before the first one): the previous call's deposits are cleared cell by cell instead of
    zeroing the whole field (lbm_particles_couple_sparse, include/lbm_b200.h); on a slab engine clear_interface = False skips the
    whole-plane clear of the interface planes (lbm_particles_couple_slab)."""
    cfg = engine.cfg
    rho_w = np.float32(cfg.WATER_DENSITY_90C if water_density is None else water_density)
    mu_w = np.float32(cfg.WATER_VISCOSITY_90C * cfg.WATER_DENSITY_90C if water_viscosity is None else water_viscosity)
    st = ps.struct()
    if sparse_clear:
        engine._check(engine.lib.lbm_particles_couple_slab(engine._ctx, _ptr(engine.u), _ptr(reaction), C.byref(st), float(rho_w), float(mu_w),
                                                           float(relax), 1 if clear_interface else 0, engine.stream), "lbm_particles_couple_slab")
    else:
        engine._check(engine.lib.lbm_particles_couple(engine._ctx, _ptr(engine.u), _ptr(reaction), C.byref(st), float(rho_w), float(mu_w),
                                                      float(relax), engine.stream), "lbm_particles_couple")


def particles_couple_slab(engine: D3Q19Engine, ps: ParticleState, reaction: torch.Tensor, relax: float = 0.8,
                          water_density: Optional[float] = None, water_viscosity: Optional[float] = None, sparse_clear: bool = False,
                          sync: str = "all", interface_guard: bool = False):
    """Two-way coupling on a z-slab engine: every rank holds all particles; a particle is computed by the rank whose slab holds
    its base cell (the kernel tests that itself on a slab engine).  Around
    it: ghost planes of u in (the trilinear gather reaches one plane up), the top ghost plane of the reaction field out and
    added to the rank above (the scatter reaches one plane up), then ONE packed all-reduce of the output arrays so that the
    replicated state stays identical (torch.distributed: NCCL on the device, gloo in tests/test_slab_gloo.py, where the kernel
    source runs CPU-emulated).

    interface_guard: a coffee bed sits in the lower part of the cone, planes away from most slab interfaces, and the integrator
    moves a particle by at most one lattice unit per step (coffee_particles.py:641-720).  After a full call the ranks agree (one
    4-byte all-reduce) on M = the smallest distance, in planes, between any owned particle's gather / scatter stencil and a slab
    interface; the next M - 1 calls then run WITHOUT the three exchanges -- nothing can have reached an interface -- and the call
    after that is a full one again with sync = "all", which brings every rank's copy of the per-particle outputs up to date before
    ownership can change.  (1 M particles in a V60 1024^3 box on 8 B200s: 0.69 -> 0.3 ms per coupling call.)"""
    from . import slab
    per_z = engine.periodic[2]
    guard_left = getattr(ps, "_interface_guard", 0) if interface_guard else 0
    active_all = ps.active

    def couple():
        # the kernel itself skips particles whose base cell lies in another slab: `active` needs no masking (a guarded call is
        # two launches and no torch op).  The interface planes of `reaction` are cleared whole only when a neighbour's deposits
        # were added to them since the last call.
        dirty = getattr(ps, "_interface_dirty", True)
        particles_couple(engine, ps, reaction, relax=relax, water_density=water_density, water_viscosity=water_viscosity,
                         sparse_clear=sparse_clear, clear_interface=dirty)
        ps._interface_dirty = False

    if guard_left > 0:
        couple()
        ps._interface_guard = guard_left - 1
        return
    owned = slab.particle_owner_mask(ps.pos[2], active_all, engine.z0, engine.nz, engine.nz_global)
    slab.exchange_planes(engine.u, engine.rank, engine.nranks, per_z)
    couple()
    ps._interface_dirty = True
    slab.reduce_ghost_up(reaction, engine.rank, engine.nranks, per_z)
    if sync == "state" and relax >= 0.0 and not interface_guard:
        # what the next step needs on whichever rank owns the particle then: the under-relaxed drag (the kernel leaves
        # drag_old == drag).  The diagnostics (drag_new, u_fluid, reynolds, cd, cell) stay valid on the owner only: 12 MB
        # instead of 68 MB per million particles and step.
        slab.allreduce_owned_packed([ps.drag], owned, active_all)
        ps.drag_old.copy_(ps.drag)
    else:
        outs = [ps.drag_new, ps.u_fluid, ps.reynolds, ps.cd, ps.cell] + ([ps.drag, ps.drag_old] if relax >= 0.0 else [])
        slab.allreduce_owned_packed(outs, owned, active_all)
    if interface_guard:
        ps._interface_guard = _interface_margin(engine, ps, owned)


def _interface_margin(engine: D3Q19Engine, ps: ParticleState, owned: torch.Tensor) -> int:
    """Calls that may skip the slab exchanges: min over the ranks of (planes between an owned particle's stencil [k, k + 1] and the
    rank's interfaces) - 3, capped; 0 when some particle is that close (then every call is a full one)."""
    import torch.distributed as dist
    cap = 1 << 20
    k = ps.pos[2].clamp(0.0, float(engine.nz_global - 2)).to(torch.int32)
    mine = owned != 0
    big = torch.full((), cap, dtype=torch.int32, device=k.device)
    lo = torch.where(mine, k - engine.z0, big).min() if engine.z0 > 0 or engine.periodic[2] else big
    hi = torch.where(mine, (engine.z0 + engine.nz - 1) - (k + 1), big).min() if engine.z0 + engine.nz < engine.nz_global or engine.periodic[2] else big
    m = torch.minimum(lo, hi).to(torch.int64).reshape(1)
    if engine.nranks > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MIN)
    margin = int(m.item()) - 2
    return min(max(margin - 1, 0), 4096)


def particles_fluid_forces_slab(engine: D3Q19Engine, ps: ParticleState, force: torch.Tensor, counters: torch.Tensor,
                                water_density: float, water_viscosity: float, gravity: float) -> None:
    """CoffeeParticleSystem.apply_fluid_forces (coffee_particles.py:547-639) on a z-slab engine: replicated particles, the rank whose
    slab holds the particle's base cell computes (the kernel reads u at that cell only, no ghost plane involved).  The kernel's side
    effects on the owner -- the force, a velocity it reset, a particle it deactivated, the error counter -- reach every rank through
    one packed all-reduce, so the replicated arrays stay identical."""
    import ctypes as C
    import torch.distributed as dist
    from . import slab
    active_all = ps.active
    owned = slab.particle_owner_mask(ps.pos[2], active_all, engine.z0, engine.nz, engine.nz_global)
    work = owned.clone()
    local = torch.zeros_like(counters)
    ps.active = work
    try:
        st = ps.struct()
        engine._check(engine.lib.lbm_particles_fluid_forces(engine._ctx, _ptr(engine.u), C.byref(st), _ptr(force), float(water_density),
                                                            float(water_viscosity), float(gravity), _ptr(local), engine.stream),
                      "lbm_particles_fluid_forces")
    finally:
        ps.active = active_all
    slab.allreduce_owned_packed([force, ps.vel, work], owned, active_all)      # work: 1 where the owner kept the particle
    active_all.mul_(torch.where(active_all != 0, work, torch.ones_like(work)))
    if engine.nranks > 1:
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
    counters.add_(local)


def particles_block_at_filter_slab(engine: D3Q19Engine, ps: ParticleState, accumulated: torch.Tensor, scale_length: float, noise: float,
                                   seed: int) -> None:
    """FilterPaperSystem.block_particles_at_filter (filter_paper.py:616-700) on a z-slab engine.  The reference takes the FIRST filter
    cell among the planes gz - 2 .. gz + 2 of the particle's column; that range can straddle a slab interface.  Every rank looks up
    the lowest such plane among the planes it OWNS, the ranks agree on the minimum (one all-reduce), and the rank that owns that plane
    runs the kernel for the particle (no earlier filter cell exists anywhere, so the kernel's own search ends on that plane); particles
    without a filter cell in range stay with the owner of their base cell, which leaves them unchanged.  One packed all-reduce brings
    the velocities back."""
    import ctypes as C
    import torch.distributed as dist
    from . import slab
    active_all = ps.active
    n_big = 1 << 30
    sl = torch.tensor(scale_length, dtype=torch.float32, device=ps.pos.device)
    g = torch.div(ps.pos, sl).to(torch.int32)                                   # (int)(pos / SCALE_LENGTH), f32 division as in the kernel
    gx, gy, gz = g[0].long(), g[1].long(), g[2].long()
    inside = (gx >= 0) & (gx < engine.nx) & (gy >= 0) & (gy < engine.ny) & (gz >= 0) & (gz < engine.nz_global) & (active_all != 0)
    first = torch.full_like(gz, n_big)
    flags = engine.flags                                                         # [nz + 2, ny, nx], plane zp = k - z0 + 1
    cx_, cy_ = gx.clamp(0, engine.nx - 1), gy.clamp(0, engine.ny - 1)
    for off in (2, 1, 0, -1, -2):                                                # descending: the last hit written is the lowest plane
        k = gz + off
        mine = inside & (k >= engine.z0) & (k < engine.z0 + engine.nz)
        zp = (k - engine.z0 + 1).clamp(0, engine.nz + 1)
        hit = mine & ((flags[zp, cy_, cx_] & L.FLAG_FILTER) != 0)
        first = torch.where(hit, k, first)
    if engine.nranks > 1:
        dist.all_reduce(first, op=dist.ReduceOp.MIN)
    has = first < n_big
    base_owner = slab.particle_owner_mask(ps.pos[2], active_all, engine.z0, engine.nz, engine.nz_global) != 0
    owner = torch.where(has, (first >= engine.z0) & (first < engine.z0 + engine.nz), base_owner) & (active_all != 0)
    owned = owner.to(torch.int32)
    ps.active = torch.where(has, owned, torch.zeros_like(owned))                 # the kernel runs for particles with a filter cell in range
    try:
        st = ps.struct()
        engine._check(engine.lib.lbm_particles_block_at_filter(engine._ctx, C.byref(st), _ptr(engine.flags), _ptr(accumulated),
                                                               float(scale_length), float(noise), int(seed) & 0xFFFFFFFF, engine.stream),
                      "lbm_particles_block_at_filter")
    finally:
        ps.active = active_all
    slab.allreduce_owned_packed([ps.vel], owned, active_all)


def particles_advance(engine: D3Q19Engine, ps: ParticleState, dt: float, center_x: float, center_y: float, bottom_z: float,
                      bottom_radius_lu: float, top_radius_lu: float, force: Optional[torch.Tensor] = None,
                      counters: Optional[torch.Tensor] = None) -> torch.Tensor:
    """CoffeeParticleSystem.update_particle_physics (coffee_particles.py:641-720) on the device.  `force` is the
    reference's self.force as a [3,n] tensor (zeroed by the call); returns the int32 counters tensor
    [coordinate_errors, boundary_violations] (accumulated into `counters` when given)."""
    cfg = engine.cfg
    cup = np.float32(cfg.CUP_HEIGHT / cfg.SCALE_LENGTH)
    if not cup > 0:
        cup = np.float32(50.0)
    b = L.LbmParticleBounds(center_x=float(center_x), center_y=float(center_y), bottom_z=float(bottom_z),
                            bottom_radius_lu=float(bottom_radius_lu), top_radius_lu=float(top_radius_lu), cup_height_lu=float(cup),
                            max_coordinate=float(max(cfg.NX, cfg.NY, cfg.NZ)), nz_minus_5=float(cfg.NZ - 5))
    if counters is None:
        counters = torch.zeros(2, dtype=torch.int32, device=ps.pos.device)
    st = ps.struct()
    engine._check(engine.lib.lbm_particles_advance(engine._ctx, C.byref(st), _ptr(force) if force is not None else None, C.byref(b),
                                                   float(dt), _ptr(counters), engine.stream), "lbm_particles_advance")
    return counters

