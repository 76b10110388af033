"""Facades of the physics modules whose work is fused into (or feeds) the D3Q19 step.

Names and call signatures follow the reference so main.py-style orchestration keeps working:
  BoundaryConditionManager   src/physics/boundary_conditions.py:326-556
  LESTurbulenceModel         src/physics/les_turbulence.py:40-389
  FilterPaperSystem          src/physics/filter_paper.py:47-926          (geometry + drag only)
  PressureGradientDrive      src/physics/pressure_gradient_drive.py:14-397 (force mode B / mixed)
  CoffeeParticleSystem       src/physics/coffee_particles.py:14-1245      (two-way coupling only)
  MultiphaseFlow3D           src/core/multiphase_3d.py:12-579             (surface tension + phase-field step)
  PrecisePouringSystem       src/physics/precise_pouring.py:12-404        (nozzle force + gradual phase change)
All device work goes through the C ABI; nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch

from . import _lib as L
from . import config as cfgmod
from .engine import ParticleState, particles_advance, particles_couple, _ptr
from .fields import ScalarCount, ScalarField, TensorField, VectorField


# ----------------------------------------------------------------------------------------------------
class BoundaryConditionManager:
    """Strategy order of the reference: bounce_back, filter_paper, top, bottom, outlet (:384-409).
    bounce-back and the filter damping live inside the fused step kernel; what remains observable is
    the density written on open faces (quirk Q5), done by lbm_face_bc."""

    def __init__(self, filter_system=None):
        self.filter_system = filter_system
        self.enabled = True

    def set_filter_system(self, filter_system):
        self.filter_system = filter_system

    def apply_all_boundaries(self, solver) -> None:
        if self.enabled and solver.engine.flags is not None and solver.engine.compat_name == "reference":
            solver.engine.face_bc()

    def apply(self, solver) -> bool:
        self.apply_all_boundaries(solver)
        return True

    def apply_fallback(self, solver) -> bool:
        try:
            self.apply_all_boundaries(solver)
            return True
        except Exception:
            return False

    def initialize_all_boundaries(self, geometry_system=None, filter_system=None, multiphase_system=None):
        if filter_system is not None:
            self.set_filter_system(filter_system)

    def get_boundary_info(self) -> Dict[str, str]:
        return {"bounce_back": "halfway bounce-back fused in the step kernel", "outlet": "open faces (rho extrapolation)",
                "top": "top face rho=1", "bottom": "bottom face rho extrapolation"}

    def get_priority_order(self) -> list:
        return ["bounce_back", "filter_paper", "top", "bottom", "outlet"]


# ----------------------------------------------------------------------------------------------------
class LESTurbulenceModel:
    """The Smagorinsky model is evaluated inside the step kernel (reference: FD on the previous step's u,
    physical: local non-equilibrium stress).  This object keeps the reference's configuration surface."""

    def __init__(self, solver):
        self._solver = solver
        self.cs = solver.config.LES_CS

    def set_mask(self, mask_field):
        if mask_field is not self._solver.les_mask:
            self._solver.les_mask.copy_from(mask_field)

    def set_phase_field(self, phase_field):
        pass        # the solver's own phase field is what the kernel reads (legacy/lbm_solver.py:101-105)

    def update_turbulent_viscosity(self, u_field=None):
        pass        # fused: nothing to do between steps


# ----------------------------------------------------------------------------------------------------
class FilterPaperSystem:
    def __init__(self, lbm_solver: Any):
        self.lbm = lbm_solver
        cfg = lbm_solver.config
        self.PAPER_THICKNESS, self.PAPER_POROSITY = cfg.PAPER_THICKNESS, cfg.PAPER_POROSITY
        self.PAPER_PORE_SIZE, self.PAPER_PERMEABILITY = 20e-6, 1e-12
        self.filter_zone = lbm_solver.filter_zone
        self.filter_bottom_z = None
        self.filter_thickness_lu = None
        e = lbm_solver.engine
        self._blockage = None
        self.filter_blockage = None

    def initialize_filter_geometry(self) -> None:
        """filter_paper.py:136-197: V60 solid mask, filter zones, Forchheimer parameters, LES punch-out."""
        cfg = self.lbm.config
        self.filter_bottom_z = 5.0
        self.filter_thickness_lu = max(1, int(self.PAPER_THICKNESS / cfg.SCALE_LENGTH))
        s = self.lbm
        e = s.engine

        def mutate():
            geom = e.cfg.v60_geometry_constants()
            import ctypes as C
            arr = (C.c_float * 5)(*geom)
            e._check(e.lib.lbm_build_v60_geometry(e._ctx, _ptr(e.solid), _ptr(e.filter_zone), arr, e.stream), "lbm_build_v60_geometry")
            e.les_mask[e.filter_zone == 1] = 0
        if e.steps_done > 0 and e.compat_name == "reference":
            e.set_geometry_preserving_f(mutate)
        else:
            mutate(); e.pack_flags()
        s._flags_dirty = False
        s.boundary_manager.set_filter_system(self)
        if self._blockage is None:
            self._blockage = torch.zeros_like(e.rho)
            e.blockage = self._blockage
            self.filter_blockage = ScalarField(lambda: self._blockage, e.zghost)

    setup_filter_geometry = initialize_filter_geometry

    def compute_forchheimer_resistance(self) -> None:
        """filter_paper.py:471-536: body_force += Forchheimer drag in the filter zone."""
        self.lbm.engine.add_forchheimer_force()

    def apply_filter_effects(self) -> None:
        pass        # filter_paper.py:538-614 is fused into the step kernel's u write-out (reference compat)

    def get_filter_statistics(self) -> Dict[str, Any]:
        z = self.lbm.engine.filter_zone
        return {"total_filter_nodes": int((z == 1).sum()), "filter_fraction": float((z == 1).float().mean())}

    def get_filter_inner_radius_at_height(self, z: float) -> float:
        """filter_paper.py:840-854: inner radius of the cone at height z (lattice units), f32 like the kernel."""
        cfg = self.lbm.config
        f = np.float32
        cup = f(cfg.CUP_HEIGHT / cfg.SCALE_LENGTH)
        top, bot = f(cfg.TOP_RADIUS / cfg.SCALE_LENGTH), f(cfg.BOTTOM_RADIUS / cfg.SCALE_LENGTH)
        h = (f(z) - f(self.filter_bottom_z if self.filter_bottom_z is not None else 5.0)) / cup
        h = max(f(0.0), min(f(1.0), h))
        return float(bot + (top - bot) * h)

    def get_coffee_bed_boundary(self) -> Dict[str, Any]:
        """filter_paper.py:856-900: the cone the coffee bed lives in (main.py:672-679 feeds it to the particle integrator)."""
        cfg = self.lbm.config
        bottom = self.filter_bottom_z if self.filter_bottom_z is not None else 5.0
        return {"center_x": cfg.NX * 0.5, "center_y": cfg.NY * 0.5, "bottom_z": bottom,
                "top_z": bottom + cfg.CUP_HEIGHT / cfg.SCALE_LENGTH,
                "top_radius_lu": cfg.TOP_RADIUS / cfg.SCALE_LENGTH, "bottom_radius_lu": cfg.BOTTOM_RADIUS / cfg.SCALE_LENGTH,
                "get_radius_at_height": self.get_filter_inner_radius_at_height}

    def _ensure_accumulated(self):
        if self._blockage is None:
            raise RuntimeError("FilterPaperSystem: call initialize_filter_geometry() first")
        if getattr(self, "_accumulated", None) is None:
            self._accumulated = torch.zeros_like(self._blockage)
            self.accumulated_particles = ScalarField(lambda: self._accumulated, self.lbm.engine.zghost)

    def update_dynamic_resistance(self) -> None:
        """filter_paper.py:703-746: blockage <- 0.95 blockage + 0.05 * 0.9 (1 - exp(-0.1 accumulated)); accumulated *= 0.999 in
        the filter zone (lbm_filter_dynamic_resistance).  The blockage field feeds the step kernel's filter damping."""
        if self._blockage is None:
            return
        self._ensure_accumulated()
        self.lbm._sync_flags()
        e = self.lbm.engine
        e._check(e.lib.lbm_filter_dynamic_resistance(e._ctx, _ptr(e.flags), _ptr(self._blockage), _ptr(self._accumulated), e.stream),
                 "lbm_filter_dynamic_resistance")

    def block_particles_at_filter(self, particle_positions=None, particle_velocities=None, particle_radii=None, particle_active=None,
                                  particle_count=None, particle_system=None, noise: float = 0.01, seed: Optional[int] = None) -> None:
        """filter_paper.py:616-700 (lbm_particles_block_at_filter).  The reference passes the five Taichi fields of the particle
        system; here the particle arrays belong to a `CoffeeParticleSystem`, so pass it as `particle_system` (or as the first
        positional argument).  The horizontal kick of the reference comes from Taichi's unseeded ti.random(); here it is a
        counter-based draw from (seed, particle) -- `seed` defaults to a per-call counter."""
        ps = particle_system if particle_system is not None else particle_positions
        if not hasattr(ps, "state"):
            raise TypeError("block_particles_at_filter needs the CoffeeParticleSystem that owns the particle arrays")
        self._ensure_accumulated()
        self.lbm._sync_flags()
        self._block_calls = getattr(self, "_block_calls", 0) + 1
        e = self.lbm.engine
        if e.zghost:
            from .engine import particles_block_at_filter_slab
            particles_block_at_filter_slab(e, ps.state, self._accumulated, float(np.float32(self.lbm.config.SCALE_LENGTH)), float(noise),
                                           int(self._block_calls if seed is None else seed))
            return
        st = ps.state.struct()
        import ctypes as C
        e._check(e.lib.lbm_particles_block_at_filter(e._ctx, C.byref(st), _ptr(e.flags), _ptr(self._accumulated),
                                                     float(np.float32(self.lbm.config.SCALE_LENGTH)), float(noise),
                                                     int(self._block_calls if seed is None else seed) & 0xFFFFFFFF, e.stream),
                 "lbm_particles_block_at_filter")

    def step(self, particle_system: Optional[Any] = None) -> None:
        """filter_paper.py:748-790: one filter time step.  The drag on the fluid is inside the fused step kernel
        (apply_filter_effects); then particle interception and the blockage update."""
        self.apply_filter_effects()
        if particle_system is not None:
            self.block_particles_at_filter(particle_system=particle_system)
        self.update_dynamic_resistance()

    def print_status(self) -> None:
        st = self.get_filter_statistics()
        print(f"FilterPaperSystem: {st['total_filter_nodes']:,} filter nodes ({100 * st['filter_fraction']:.2f} % of the box), "
              f"bottom z = {self.filter_bottom_z}, thickness = {self.filter_thickness_lu} lu")


# ----------------------------------------------------------------------------------------------------
class PressureGradientDrive:
    def __init__(self, lbm_solver: Any):
        self.lbm = lbm_solver
        self.MAX_PRESSURE_FORCE = 0.12           # pressure_gradient_drive.py:30 (host-mutable, main.py:913-925)
        self.ADJUSTMENT_RATE = 0.025             # :27
        self.force_drive_active = False
        self.mixed_drive_active = False
        self.density_drive_active = False

    def activate_force_drive(self, active: bool = True):
        """pressure_gradient_drive.py:81-86: switching one mode on (or off) clears the other two."""
        self.force_drive_active = bool(active); self.mixed_drive_active = False; self.density_drive_active = False

    def activate_mixed_drive(self, active: bool = True):
        self.mixed_drive_active = bool(active); self.force_drive_active = False; self.density_drive_active = False

    def activate_density_drive(self, active: bool = True):
        # method A writes rho, which the next macroscopic pass overwrites: inert in LBMSolver
        self.density_drive_active = bool(active); self.force_drive_active = False; self.mixed_drive_active = False

    def apply_force_drive(self):
        if self.force_drive_active:
            self.lbm.engine.add_pressure_gradient_force(self.MAX_PRESSURE_FORCE, 1.0)

    def apply_mixed_drive(self):
        if self.mixed_drive_active:
            self.lbm.engine.add_pressure_gradient_force(self.MAX_PRESSURE_FORCE, 0.5)

    def compute_pressure_gradient(self):
        """pressure_gradient_drive.py:124-177: the clamped acceleration field -cs^2 grad(rho)/rho, kept in
        `pressure_force` ([3, z, y, x] device tensor; zero on solid cells) without touching body_force."""
        e = self.lbm.engine
        if getattr(self, "_pressure_force", None) is None:
            self._pressure_force = torch.zeros_like(e.u)
            self.pressure_force = VectorField(lambda: self._pressure_force, e.zghost)
        self._pressure_force.zero_()
        e.exchange_field(scalar=e.rho)        # z-slabs: the gradient reads rho[k -+ 1] across the interface
        e._check(e.lib.lbm_pressure_gradient_force_set(e._ctx, _ptr(e.rho), _ptr(e.flags), _ptr(self._pressure_force),
                                                       float(self.MAX_PRESSURE_FORCE), 1.0, e.stream), "lbm_pressure_gradient_force_set")

    def initialize_target_density(self):
        """pressure_gradient_drive.py:54-72: z profile 0.4 -> 1 over the bottom 20 %, 1 in the middle, 1 -> 1.8 over the top
        20 % (global z).  f32 with the reference's constant folding: expressions of Python floats (1.0 - 0.8, 1.8 - 1.0, ...) fold
        in f64 and are rounded to f32 when they meet a kernel value.  One value per plane of the slab, ghost planes included."""
        e = self.lbm.engine
        f = np.float32
        k = np.arange(e.z0 - e.zghost, e.z0 + e.nz + e.zghost, dtype=np.float32)
        zr = k / f(e.nz_global)
        hi = f(1.0) + ((zr - f(0.8)) / f(1.0 - 0.8)) * f(1.8 - 1.0)
        lo = f(0.4) + (zr / f(0.2)) * f(1.0 - 0.4)
        prof = np.where(zr >= f(0.8), hi, np.where(zr <= f(0.2), lo, f(1.0))).astype(np.float32)
        self._target_profile = torch.from_numpy(prof).to(e.device)
        self._target_density = self._target_profile[:, None, None].expand(-1, e.ny, e.nx)

    @property
    def target_density(self):
        """the reference's target_density field (logical [i, j, k] order, owned planes)"""
        if getattr(self, "_target_profile", None) is None:
            self.initialize_target_density()
        e = self.lbm.engine
        if getattr(self, "_target_field", None) is None:
            self._target_full = self._target_profile[:, None, None].expand(-1, e.ny, e.nx).contiguous()
            self._target_field = ScalarField(lambda: self._target_full, e.zghost)
        return self._target_field

    def apply_density_drive(self):
        """pressure_gradient_drive.py:95-122 (method A, lbm_density_drive): nudges rho toward the target profile by 0.025 of the
        difference, at most 0.001 per call, clamped to [0.5, 2].  LBMSolver recomputes rho from f before it is used, so this is
        cosmetic there; bit-exact against the recorded reference run (tests/test_gpu_geometry_particles.py)."""
        if not self.density_drive_active:
            return
        e = self.lbm.engine
        if getattr(self, "_target_profile", None) is None:
            self.initialize_target_density()
        self.lbm._sync_flags()
        e._check(e.lib.lbm_density_drive(e._ctx, _ptr(e.rho), _ptr(e.flags), _ptr(self._target_profile), float(np.float32(self.ADJUSTMENT_RATE)),
                                         0.001, 0.5, 2.0, e.stream), "lbm_density_drive")

    def apply(self, step: int = 0):
        """pressure_gradient_drive.py:257-272"""
        if self.density_drive_active:
            self.apply_density_drive()
        if self.force_drive_active:
            self.lbm.engine.add_pressure_gradient_force(self.MAX_PRESSURE_FORCE, 1.0)
        elif self.mixed_drive_active:
            self.lbm.engine.add_pressure_gradient_force(self.MAX_PRESSURE_FORCE, 0.5)

    def get_status(self) -> Dict[str, Any]:
        return {"density_drive": bool(self.density_drive_active), "force_drive": bool(self.force_drive_active),
                "mixed_drive": bool(self.mixed_drive_active), "max_pressure_force": float(self.MAX_PRESSURE_FORCE)}

    def get_statistics(self) -> Dict[str, Any]:
        """pressure_gradient_drive.py:281-330: extrema of the fields the drive acts on (from the fused statistics pass)."""
        s = self.lbm.engine.field_statistics().tolist()
        return {"max_pressure": s[2], "min_pressure": s[1], "pressure_drop": s[2] - s[1], "max_velocity": s[0],
                "mean_density": s[3] / max(1.0, s[7] - s[5] - s[6])}

    get_enhanced_diagnostics = get_statistics
    compute_statistics = get_statistics

    def check_enhanced_stability(self) -> bool:
        s = self.lbm.engine.field_statistics().tolist()
        return bool(s[5] == 0 and s[6] == 0 and s[0] < 0.3 and s[1] > 0.1 and s[2] < 5.0)


# ----------------------------------------------------------------------------------------------------
class CoffeeParticleSystem:
    """Two-way coupling part of CoffeeParticleSystem (coffee_particles.py:1048-1212)."""

    def __init__(self, max_particles: int = 15000, solver=None, device=None):
        self.max_particles = int(max_particles)
        self._solver = solver
        dev = device if device is not None else (solver.engine.device if solver is not None else torch.device("cuda"))
        self.state = ParticleState(self.max_particles, dev)
        self.reaction_force_tensor = None
        self.particle_count = 0
        cfg = solver.config if solver is not None else None
        self.water_density = cfg.WATER_DENSITY_90C if cfg else 965.3
        self.water_viscosity = (cfg.WATER_VISCOSITY_90C * cfg.WATER_DENSITY_90C) if cfg else 3.15e-7 * 965.3
        self.coffee_density = cfg.COFFEE_BEAN_DENSITY if cfg else 1200.0
        self.gravity = 9.81
        if solver is not None:
            self.bind(solver)

    def bind(self, solver):
        self._solver = solver
        e = solver.engine
        self.reaction_force_tensor = torch.zeros_like(e.u)
        self.reaction_force_field = VectorField(lambda: self.reaction_force_tensor, e.zghost)

    # field-like accessors in the reference's [P,3] order
    def _pv(self, t):  # [3,n] -> [n,3]
        return TensorField.wrap(t.t())

    # tensors that also answer to_numpy / from_numpy / fill, the calls the reference's diagnostics make on these arrays
    position = property(lambda self: self._pv(self.state.pos))
    velocity = property(lambda self: self._pv(self.state.vel))
    drag_force = property(lambda self: self._pv(self.state.drag))
    drag_force_new = property(lambda self: self._pv(self.state.drag_new))
    drag_force_old = property(lambda self: self._pv(self.state.drag_old))
    fluid_velocity_at_particle = property(lambda self: self._pv(self.state.u_fluid))
    radius = property(lambda self: TensorField.wrap(self.state.radius))
    mass = property(lambda self: TensorField.wrap(self.state.mass))
    active = property(lambda self: TensorField.wrap(self.state.active))
    particle_reynolds = property(lambda self: TensorField.wrap(self.state.reynolds))
    drag_coefficient = property(lambda self: TensorField.wrap(self.state.cd))
    cell_index = property(lambda self: self._pv(self.state.cell))

    @property
    def particle_count(self) -> ScalarCount:          # the reference reads `particle_count[None]` (a 0-D Taichi field)
        return ScalarCount(self._particle_count)

    @particle_count.setter
    def particle_count(self, n: int) -> None:
        self._particle_count = int(n)

    def set_particles(self, pos, vel=None, radius=None, mass=None):
        """Inject explicit particle arrays ([P,3] positions in lattice units; SI radius/mass -- quirk Q9/Q10)."""
        st = self.state
        pos = torch.as_tensor(np.asarray(pos, np.float32))
        n = pos.shape[0]
        assert n <= self.max_particles
        dev = st.pos.device
        st.pos[:, :n] = pos.t().to(dev)
        if vel is not None: st.vel[:, :n] = torch.as_tensor(np.asarray(vel, np.float32)).t().to(dev)
        r = np.full(n, 3.25e-4, np.float32) if radius is None else np.asarray(radius, np.float32)
        st.radius[:n] = torch.as_tensor(r).to(dev)
        if mass is None:   # coffee_particles.py:199-200: volume = (4/3)*3.14159*r**3 ; mass = volume*rho_coffee
            mass = ((np.float32(4.0 / 3.0) * np.float32(3.14159)) * (r * r * r)) * np.float32(self.coffee_density)
        st.mass[:n] = torch.as_tensor(np.asarray(mass, np.float32)).to(dev)
        st.active.zero_(); st.active[:n] = 1
        self.particle_count = n

    # ---- validation helpers (coffee_particles.py:75-117) ----
    MIN_COORDINATE, MAX_VELOCITY, MAX_RADIUS, MIN_RADIUS = 0.0, 10.0, 0.01, 1e-5

    @property
    def MAX_COORDINATE(self) -> float:
        c = self._solver.config
        return float(max(c.NX, c.NY, c.NZ))

    def validate_coordinate(self, x, y, z) -> bool:
        m = self.MAX_COORDINATE
        return bool(all(np.isfinite(v) and 0.0 <= v <= m and abs(v) <= 1e6 for v in (x, y, z)))

    def validate_velocity(self, vx, vy, vz) -> bool:
        return bool(all(v == v for v in (vx, vy, vz)) and vx * vx + vy * vy + vz * vz <= self.MAX_VELOCITY ** 2)

    def validate_radius(self, radius) -> bool:
        return bool(radius == radius and self.MIN_RADIUS <= radius <= self.MAX_RADIUS)

    # ---- creation (host side, like the reference: these run once, in Python) ----
    def clear_all_particles(self) -> None:
        """coffee_particles.py:119-144"""
        st = self.state
        for t in (st.pos, st.vel, st.radius, st.mass, st.drag, st.drag_new, st.drag_old, st.u_fluid, st.reynolds, st.cd):
            t.zero_()
        st.active.zero_()
        if self.reaction_force_tensor is not None:
            self.reaction_force_tensor.zero_()
        if getattr(self, "error_counters", None) is not None:
            self.error_counters.zero_()
        self.particle_count = 0

    def generate_gaussian_particle_radius(self, mean_radius=None, std_dev_ratio: float = 0.3, rng=None) -> float:
        """coffee_particles.py:146-182: N(mean, 30 %) clipped to [0.5, 1.5] x mean."""
        cfg = self._solver.config
        mean = cfg.COFFEE_PARTICLE_RADIUS if mean_radius is None else mean_radius
        mean = max(self.MIN_RADIUS, min(self.MAX_RADIUS, mean))
        rng = rng if rng is not None else np.random
        r = rng.normal(mean, mean * std_dev_ratio)
        return float(np.clip(r, max(self.MIN_RADIUS, 0.5 * mean), min(self.MAX_RADIUS, 1.5 * mean)))

    def create_particle_with_physics(self, idx: int, px, py, pz, radius, vx=0.0, vy=0.0, vz=0.0) -> int:
        """coffee_particles.py:184-218: one particle with mass = (4/3) 3.14159 r^3 rho_coffee (f32); 1 on success."""
        if not (0 <= idx < self.max_particles and self.validate_coordinate(px, py, pz) and self.validate_velocity(vx, vy, vz)
                and self.validate_radius(radius)):
            return 0
        f = np.float32
        r = f(radius)
        mass = ((f(4.0) / f(3.0)) * f(3.14159)) * (r * (r * r)) * f(self.coffee_density)
        if not (mass == mass and mass > 0):
            return 0
        st = self.state
        st.pos[:, idx] = torch.tensor([px, py, pz], dtype=torch.float32)
        st.vel[:, idx] = torch.tensor([vx, vy, vz], dtype=torch.float32)
        st.radius[idx] = float(r); st.mass[idx] = float(mass); st.active[idx] = 1
        self.particle_count = max(self.particle_count, idx + 1)
        return 1

    def initialize_coffee_bed_confined(self, filter_paper_system, seed: Optional[int] = None) -> int:
        """coffee_particles.py:220-412: the coffee bed as layered discs above the filter surface inside the V60 cone
        (up to 2000 particles, 10-30 layers, radius concentrated toward the axis, Gaussian grain sizes).  The reference
        draws with the unseeded global NumPy generator one particle at a time; here each layer is drawn in one
        vectorised batch from `seed` (same distributions and acceptance tests), then uploaded once."""
        if self._solver is None:             # main.py:477 constructs the system bare; the filter system knows the solver
            self.bind(filter_paper_system.lbm)
        cfg = self._solver.config
        rng = np.random.default_rng(seed)
        self.clear_all_particles()
        b = filter_paper_system.get_coffee_bed_boundary()
        cx, cy, bz = float(b["center_x"]), float(b["center_y"]), float(b["bottom_z"])
        top_r, bot_r = float(b["top_radius_lu"]), float(b["bottom_radius_lu"])
        if not all(np.isfinite([cx, cy, bz, top_r, bot_r])) or not (0 <= cx <= cfg.NX and 0 <= cy <= cfg.NY and 0 <= bz <= cfg.NZ):
            return 0
        bed_bottom = bz + 2.0
        bed_h = max(5.0, min(30.0, max(0.005, min(0.05, getattr(cfg, "COFFEE_BED_HEIGHT_PHYS", 0.015))) / cfg.SCALE_LENGTH))
        bed_top = bed_bottom + bed_h
        if bed_top >= cfg.NZ - 5:
            bed_top = cfg.NZ - 5; bed_h = bed_top - bed_bottom
        target = min(2000, self.max_particles - 100)
        if target <= 0 or bed_h <= 0:
            return 0
        layers = max(10, min(int(bed_h / 2), 30))
        if target < layers:
            layers = max(1, target)
        per_layer = max(1, target // layers)
        cup_h = max(10.0, cfg.CUP_HEIGHT / cfg.SCALE_LENGTH)
        bed_r = getattr(cfg, "COFFEE_BED_TOP_RADIUS", top_r * 0.8) / cfg.SCALE_LENGTH
        pos = []
        for layer in range(layers):
            if len(pos) >= target:
                break
            z = bed_bottom + (layer / layers) * bed_h
            if z > bed_top:
                break
            ratio = min(1.0, max(0.0, (z - bz) / cup_h))
            layer_r = max(1.0, min(top_r * 1.5, bot_r + (top_r - bot_r) * ratio))
            bed_ratio = max(0.1, min(1.0, (bed_h - (z - bed_bottom)) / bed_h))
            eff_r = max(2.0, min(min(layer_r * 0.85, bed_r * (0.3 + 0.7 * bed_ratio)), 50.0))
            n_try = per_layer * 5
            ang = rng.uniform(0, 2 * np.pi, n_try)
            r = np.clip(rng.uniform(0, 1, n_try) ** 1.5 * eff_r * 0.9, 0.0, eff_r * 0.9)
            x = cx + r * np.cos(ang); y = cy + r * np.sin(ang); zf = z + rng.uniform(-0.5, 0.5, n_try)
            ok = (x >= 5.0) & (x <= cfg.NX - 5.0) & (y >= 5.0) & (y <= cfg.NY - 5.0) & (zf >= bed_bottom) & (zf <= bed_top)
            # _safe_cone_boundary_check (coffee_particles.py:414-440): inside 0.9 x the cone radius at that height
            hr = np.clip((zf - bz) / cup_h, 0.0, 1.0)
            ok &= np.hypot(x - cx, y - cy) <= (bot_r + (top_r - bot_r) * hr) * 0.9
            take = np.flatnonzero(ok)[:min(per_layer, target - len(pos))]
            pos.extend(np.stack([x[take], y[take], zf[take]], 1))
        n = len(pos)
        if n == 0:
            return 0
        mean = max(self.MIN_RADIUS, min(self.MAX_RADIUS, cfg.COFFEE_PARTICLE_RADIUS))
        radius = np.clip(rng.normal(mean, 0.3 * mean, n), max(self.MIN_RADIUS, 0.5 * mean), min(self.MAX_RADIUS, 1.5 * mean))
        self.set_particles(np.asarray(pos, np.float32), radius=radius.astype(np.float32))
        return n

    def get_particle_statistics(self) -> Dict[str, Any]:
        """coffee_particles.py:833-900: radius / position statistics of the valid active particles."""
        st = self.state
        cfg = self._solver.config
        n = self.particle_count
        act = (st.active[:n] == 1).cpu().numpy()
        r = st.radius[:n].cpu().numpy(); p = st.pos[:, :n].cpu().numpy().T
        ok = act & np.isfinite(r) & np.isfinite(p).all(1) & (r >= self.MIN_RADIUS) & (r <= self.MAX_RADIUS) & \
            (p[:, 0] >= 0) & (p[:, 0] <= cfg.NX) & (p[:, 1] >= 0) & (p[:, 1] <= cfg.NY) & (p[:, 2] >= 0) & (p[:, 2] <= cfg.NZ)
        rr = r[ok]
        valid, invalid = int(ok.sum()), int((act & ~ok).sum())
        # key for key what the reference returns (main.py:1046-1067 reads count / mean_radius / ...; recorded in reference_main_trace.json)
        return {"count": valid, "invalid_count": invalid, "coordinate_errors": self.coordinate_errors,
                "boundary_violations": self.boundary_violations, "mean_radius": float(rr.mean()) if rr.size else 0.0,
                "std_radius": float(rr.std()) if rr.size else 0.0, "min_radius": float(rr.min()) if rr.size else 0.0,
                "max_radius": float(rr.max()) if rr.size else 0.0, "positions": p[ok], "radii": rr,
                "success_rate": valid / max(1, valid + invalid) * 100}

    def compute_two_way_coupling_forces(self, fluid_u=None, relax: float = -1.0):
        """coffee_particles.py:1107-1154; relax >= 0 also applies the under-relaxation in the same kernel.
        On z-slabs the particle arrays are replicated on every rank and the owner computes (`_couple_on_slab`)."""
        e = self._solver.engine
        if e.zghost:
            self._couple_on_slab(relax)
            return
        particles_couple(e, self.state, self.reaction_force_tensor, relax=relax,
                         water_density=self.water_density, water_viscosity=self.water_viscosity)

    def _couple_on_slab(self, relax: float) -> None:
        """z-slabs: replicated particles, owner computes (engine.particles_couple_slab)."""
        from .engine import particles_couple_slab
        particles_couple_slab(self._solver.engine, self.state, self.reaction_force_tensor, relax=relax, water_density=self.water_density,
                              water_viscosity=self.water_viscosity)

    def apply_under_relaxation(self, relaxation_factor: float):
        """coffee_particles.py:1200-1212"""
        import ctypes as C
        e = self._solver.engine
        st = self.state.struct()
        e._check(e.lib.lbm_particles_under_relax(e._ctx, C.byref(st), float(relaxation_factor), e.stream), "lbm_particles_under_relax")

    def update_particle_physics(self, dt: float, center_x: float, center_y: float, bottom_z: float,
                                bottom_radius_lu: float, top_radius_lu: float) -> None:
        """coffee_particles.py:641-720 (explicit Euler + cone constraint), called by main.py:672-679 every step."""
        if getattr(self, "force_tensor", None) is None:
            self.force_tensor = torch.zeros_like(self.state.pos)
            self.error_counters = torch.zeros(2, dtype=torch.int32, device=self.state.pos.device)
        particles_advance(self._solver.engine, self.state, dt, center_x, center_y, bottom_z, bottom_radius_lu, top_radius_lu,
                          force=self.force_tensor, counters=self.error_counters)

    def apply_fluid_forces(self, fluid_u=None, fluid_v=None, fluid_w=None, fluid_density=None, pressure=None, dt: float = 0.0) -> None:
        """coffee_particles.py:547-639 (lbm_particles_fluid_forces): fills the integrator's force from the solver's velocity
        field.  The reference reads only `fluid_u` (the LBM vector field) of its six arguments; here the field is the bound
        solver's `u`."""
        import ctypes as C
        if getattr(self, "force_tensor", None) is None:
            self.force_tensor = torch.zeros_like(self.state.pos)
            self.error_counters = torch.zeros(2, dtype=torch.int32, device=self.state.pos.device)
        e = self._solver.engine
        if e.zghost:
            from .engine import particles_fluid_forces_slab
            particles_fluid_forces_slab(e, self.state, self.force_tensor, self.error_counters, float(self.water_density),
                                        float(self.water_viscosity), float(self.gravity))
            return
        st = self.state.struct()
        e._check(e.lib.lbm_particles_fluid_forces(e._ctx, _ptr(e.u), C.byref(st), _ptr(self.force_tensor), float(self.water_density),
                                                  float(self.water_viscosity), float(self.gravity), _ptr(self.error_counters), e.stream),
                 "lbm_particles_fluid_forces")

    def update_particles(self, dt: float) -> None:
        """coffee_particles.py:722-732: the public wrapper with the default V60 bounds."""
        cfg = self._solver.config
        self.update_particle_physics(dt, cfg.NX // 2, cfg.NY // 2, cfg.NZ // 4, cfg.BOTTOM_RADIUS / cfg.SCALE_LENGTH,
                                     cfg.TOP_RADIUS / cfg.SCALE_LENGTH)

    force = property(lambda self: self._pv(self.force_tensor) if getattr(self, "force_tensor", None) is not None else None)
    coordinate_errors = property(lambda self: int(self.error_counters[0]) if getattr(self, "error_counters", None) is not None else 0)
    boundary_violations = property(lambda self: int(self.error_counters[1]) if getattr(self, "error_counters", None) is not None else 0)

    def get_coupling_diagnostics(self) -> Dict[str, Any]:
        st = self.state
        act = st.active != 0
        n = int(act.sum())
        if n == 0:
            return {"active_particles": 0, "avg_reynolds": 0.0, "avg_drag_coeff": 0.0, "max_reaction_force": 0.0,
                    "coupling_quality": "no_particles"}
        r = self.reaction_force_tensor
        return {"active_particles": n, "avg_reynolds": float(st.reynolds[act].mean()), "max_reynolds": float(st.reynolds[act].max()),
                "avg_drag_coeff": float(st.cd[act].mean()), "max_reaction_force": float(torch.sqrt((r * r).sum(0)).max()),
                "coupling_quality": "active"}


# ----------------------------------------------------------------------------------------------------
class MultiphaseFlow3D:
    """src/core/multiphase_3d.py:12-579 -- the phase-field system as main.py drives it (main.py:622-631, 795-800, 839):
    `accumulate_surface_tension_pre_collision()` before the collision and `step()` after it, plus the initial-state
    helpers.  Device work: lbm_surface_tension (2 launches for the reference's 4 kernels), lbm_phase_field_step
    (2 launches for 4 kernels), lbm_density_from_phase, lbm_chemical_potential.  The live `step()` of the reference is
    the second definition (:389-407); the Cahn-Hilliard `step()` at :246-270 is shadowed by it and never runs."""

    def __init__(self, lbm_solver: Any, lazy_fields: bool = False):
        """lazy_fields = False: every call materialises grad_phi / normal / curvature / surface_force exactly when the reference
        does.  lazy_fields = True: the hot calls (accumulate_surface_tension_pre_collision, step) only update body_force -- one
        launch that touches the interface band, bit-identical body_force -- and the four diagnostic fields are computed when
        somebody reads them, from the phase field as it is at that moment (after a step() the reference's copies still
        reflect phi before the step's update)."""
        self.lbm = lbm_solver
        e = lbm_solver.engine
        cfg = lbm_solver.config
        # z-slabs: the chain runs as two calls with ghost-plane refreshes of phi / mu / normal (slab.exchange_planes); the
        # one-launch kernel reaches k -+ 2 and is single-slab only
        self.lazy_fields = bool(lazy_fields) and not e.zghost
        self._stale = False
        if e.body_force is None or e.phase is None or e.flags is None:
            raise ValueError("MultiphaseFlow3D needs a solver with body_force, phase and a flag field (force=True, phase=True, walls=True)")
        sc = lambda: torch.zeros_like(e.rho)
        vc = lambda: torch.zeros_like(e.body_force)
        self._phi, self._phi_new, self._mu, self._curv, self._lap = sc(), sc(), sc(), sc(), sc()
        self._normal, self._grad_phi, self._grad_mu, self._sf = vc(), vc(), vc(), vc()
        zg = e.zghost
        self.phi = ScalarField(lambda: self._phi, zg); self.phi_new = ScalarField(lambda: self._phi_new, zg)
        self.phi.owner = self.phi_new.owner = getattr(lbm_solver, "_solver", lbm_solver)
        self.mu = ScalarField(lambda: self._mu, zg)
        self.laplacian_phi = ScalarField(lambda: self._lap, zg)
        fresh = self._fresh
        self.curvature = ScalarField(lambda: fresh(self._curv), zg)
        self.normal = VectorField(lambda: fresh(self._normal), zg); self.grad_phi = VectorField(lambda: fresh(self._grad_phi), zg)
        self.grad_mu = VectorField(lambda: fresh(self._grad_mu), zg); self.surface_force = VectorField(lambda: fresh(self._sf), zg)
        # multiphase_3d.py:40-47
        self.INTERFACE_WIDTH = 2.0
        self.MOBILITY = 0.001
        self.SURFACE_TENSION_COEFF = cfg.SURFACE_TENSION_LU
        self.CAHN_NUMBER = 0.005
        self.BETA = 12.0 * self.SURFACE_TENSION_COEFF / self.INTERFACE_WIDTH
        self.KAPPA = 1.5 * self.SURFACE_TENSION_COEFF * self.INTERFACE_WIDTH

    def _ready(self):
        self.lbm._sync_flags()
        return self.lbm.engine

    def _ghosts(self, *tensors) -> None:
        """z-slabs: fill the ghost planes of whole-plane fields from the neighbours (no-op on a single slab)."""
        e = self.lbm.engine
        if not e.zghost:
            return
        from . import slab
        for t in tensors:
            slab.exchange_planes(t, e.rank, e.nranks, e.periodic[2])

    def _fresh(self, tensor):
        """Field access: materialise the diagnostic fields first if a lazy call skipped them."""
        if self._stale:
            self._surface_tension_fields(False)
        return tensor

    def _surface_tension_to_body_force(self) -> None:
        if self.lazy_fields:
            self._ready().surface_tension_body_force(self._phi, self.SURFACE_TENSION_COEFF, self._normal, self._sf)
            self._stale = True
        else:
            self._surface_tension_fields(True)

    # ---- kernels of the reference, same names ----------------------------------------------------------------
    def init_phase_field(self) -> None:
        """multiphase_3d.py:55-78: a dry dripper, phi = -1 everywhere."""
        self._phi.fill_(-1.0)

    def compute_chemical_potential(self) -> None:
        """multiphase_3d.py:80-109."""
        kappa = 3.0 * self.SURFACE_TENSION_COEFF * self.INTERFACE_WIDTH / 8.0
        self._ghosts(self._phi)
        self._ready().chemical_potential(self._phi, self._lap, self._mu, kappa)
        self._ghosts(self._mu)

    def _surface_tension_fields(self, apply: bool) -> None:
        self._stale = False
        e = self._ready()
        if not e.zghost:
            e.surface_tension(self._phi, self._mu, self._grad_phi, self._grad_mu, self._normal, self._curv, self._sf,
                              self.SURFACE_TENSION_COEFF, apply=apply)
            return
        self._ghosts(self._phi)
        e.surface_tension_gradients(self._phi, self._mu, self._grad_phi, self._grad_mu, self._normal)
        self._ghosts(self._normal)
        e.surface_tension_curvature_force(self._phi, self._grad_phi, self._normal, self._curv, self._sf, self.SURFACE_TENSION_COEFF, apply=apply)

    def compute_gradients(self) -> None:
        """multiphase_3d.py:111-132.  The device pass also refreshes curvature and surface_force (one fused chain)."""
        self._surface_tension_fields(False)

    compute_curvature = compute_gradients                   # :134-149, same fused pass
    compute_surface_tension_force = compute_gradients       # :313-332, same fused pass

    def apply_surface_tension(self) -> None:
        """multiphase_3d.py:354-363."""
        self._ready().apply_surface_tension(self._fresh(self._sf))

    def accumulate_surface_tension_pre_collision(self) -> None:
        """multiphase_3d.py:409-418."""
        self._surface_tension_to_body_force()

    def update_density_from_phase(self) -> None:
        """multiphase_3d.py:365-381."""
        cfg = self.lbm.config
        self._ready().density_from_phase(self._phi, cfg.RHO_WATER, cfg.RHO_AIR)

    def copy_phase_field(self) -> None:
        self._phi.copy_(self._phi_new)

    def step(self, step_count: int = 0, precollision_applied: bool = False) -> None:
        """multiphase_3d.py:389-407."""
        cfg = self.lbm.config
        if (not precollision_applied) and step_count > 10:
            self._surface_tension_to_body_force()
        elif self.lazy_fields:
            self._ready(); self._stale = True          # the three field kernels of :396-398 are deferred until somebody reads them
        else:
            self._surface_tension_fields(False)
        self._ghosts(self._phi)
        self.lbm.engine.phase_field_step(self._phi, self._phi_new, self._mu, self.MOBILITY, cfg.DT, cfg.RHO_WATER, cfg.RHO_AIR)

    # ---- initial state (multiphase_3d.py:420-579) --------------------------------------------------------------
    def _set_dry_initial_state(self) -> None:
        fluid = self._ready().solid == 0
        self._phi[fluid] = -1.0
        self._phi_new[fluid] = -1.0

    def validate_initial_phase_consistency(self) -> None:
        phi = self._phi
        bad = int(((phi < -1.1) | (phi > 1.1)).sum().item())
        if bad > 0:
            raise ValueError(f"{bad} phase-field values outside [-1, 1]")

    def standardize_initial_state(self, force_dry_state: bool = True) -> None:
        if force_dry_state:
            self._set_dry_initial_state()
        self.update_density_from_phase()
        self.compute_chemical_potential()
        self.compute_gradients()
        self.validate_initial_phase_consistency()

    def get_interface_statistics(self) -> Dict[str, Any]:
        """multiphase_3d.py:290-311."""
        cfg = self.lbm.config
        phi = self._phi
        interface = phi.abs() < 0.9
        gmag = torch.linalg.vector_norm(self._fresh(self._grad_phi), dim=0)
        thick = (1.0 / (gmag + 1e-10))[interface]
        return {"interface_volume": float(interface.sum().item()) * cfg.SCALE_LENGTH ** 3,
                "water_fraction": float((phi > 0).sum().item()) / phi.numel(),
                "interface_thickness": (float(thick.mean().item()) if thick.numel() else float("nan")) * cfg.SCALE_LENGTH,
                "max_curvature": float(self._curv.abs().max().item()),
                "surface_tension_magnitude": float(torch.linalg.vector_norm(self._sf, dim=0).max().item())}


# ----------------------------------------------------------------------------------------------------
class _Scalar0D:
    """ti.field(dtype, shape=()) surface: value[None] get / set."""

    def __init__(self, dtype):
        self._dt, self._v = dtype, dtype(0)

    def __getitem__(self, _):
        return self._v

    def __setitem__(self, _, value):
        self._v = self._dt(value)


class PrecisePouringSystem:
    """src/physics/precise_pouring.py:12-404.  The nozzle state lives on the host (a handful of scalars, as in the
    reference); the two kernels run over the nozzle's bounding box on the device.  `solver` (or `bind`) names the solver
    whose engine executes them: the reference passes the fields to each call, and the fields of this framework belong to
    an engine."""

    def __init__(self, solver: Any = None, config: Any = None):
        self.lbm = solver
        # main.py:482 constructs `PrecisePouringSystem()` bare (the reference reads its global `config` module): the package
        # default then, and the solver that owns the fields arrives with bind() or through the first field argument
        cfg = config if config is not None else (solver.config if solver is not None else cfgmod.DEFAULT)
        self.config = cfg
        self.POUR_DIAMETER_CM = 0.5
        self.POUR_DIAMETER_GRID = self.POUR_DIAMETER_CM / cfg.GRID_SIZE_CM
        self.POUR_HEIGHT_CM = cfg.POUR_HEIGHT_CM
        self.POUR_VELOCITY = cfg.INLET_VELOCITY
        v60_top_z = int(5.0 + int(cfg.CUP_HEIGHT / cfg.SCALE_LENGTH))
        self.POUR_HEIGHT = max(8, min(int(v60_top_z + 2), cfg.NZ - 6))
        f32, i32 = np.float32, np.int32
        self.pouring_active, self.pour_pattern = _Scalar0D(i32), _Scalar0D(i32)
        self.pour_center_x, self.pour_center_y, self.pour_flow_rate, self.pour_time = (_Scalar0D(f32) for _ in range(4))
        self.spiral_radius, self.spiral_speed, self.spiral_center_x, self.spiral_center_y = (_Scalar0D(f32) for _ in range(4))

    def bind(self, solver) -> None:
        self.lbm = solver

    def start_pouring(self, center_x=None, center_y=None, flow_rate=1.0, pattern="center") -> None:
        """precise_pouring.py:49-74."""
        cfg = self.config
        center_x = cfg.NX // 2 if center_x is None else center_x
        center_y = cfg.NY // 2 if center_y is None else center_y
        self.pour_center_x[None] = center_x; self.pour_center_y[None] = center_y
        self.pour_flow_rate[None] = flow_rate
        self.pouring_active[None] = 1
        self.pour_time[None] = 0.0
        if pattern == "center":
            self.pour_pattern[None] = 0
        elif pattern == "spiral":
            self.pour_pattern[None] = 1
            self.spiral_center_x[None] = center_x; self.spiral_center_y[None] = center_y
            self.spiral_radius[None] = 5.0; self.spiral_speed[None] = 1.0

    def stop_pouring(self) -> None:
        self.pouring_active[None] = 0

    def _get_current_pour_position(self):
        """precise_pouring.py:81-97, f32 like the kernel-scope original."""
        f = np.float32
        x, y = self.pour_center_x[None], self.pour_center_y[None]
        if self.pour_pattern[None] == 1:
            t = self.pour_time[None] * self.spiral_speed[None]
            r = self.spiral_radius[None] * (f(1.0) + f(0.1) * t)
            x = self.spiral_center_x[None] + r * np.cos(t)
            y = self.spiral_center_y[None] + r * np.sin(t)
            d = self.POUR_DIAMETER_GRID
            x = max(f(d), min(f(self.config.NX - d), x))
            y = max(f(d), min(f(self.config.NY - d), y))
        return f(x), f(y)

    def _pour_struct(self, dt: float) -> "L.LbmPour":
        x, y = self._get_current_pour_position()
        return L.LbmPour(pour_x=float(x), pour_y=float(y), radius=float(np.float32(self.POUR_DIAMETER_GRID / 2.0)),
                         pour_z=int(self.POUR_HEIGHT), velocity=float(np.float32(self.POUR_VELOCITY)),
                         flow_rate=float(self.pour_flow_rate[None]), dt=float(np.float32(dt)))

    @staticmethod
    def _tensor(field_or_tensor):
        return field_or_tensor._get() if hasattr(field_or_tensor, "_get") else field_or_tensor

    def _engine(self, *fields):
        if self.lbm is None:               # constructed bare like the reference's: the first field argument names its solver
            for fld in fields:
                if getattr(fld, "owner", None) is not None:
                    self.lbm = fld.owner
                    break
        if self.lbm is None:
            raise ValueError("PrecisePouringSystem is not bound to a solver (pass solver=, call bind(), or hand it the solver's fields)")
        self.lbm._sync_flags()
        return self.lbm.engine

    def apply_pouring_force(self, lbm_body_force, solid, dt: float) -> None:
        """precise_pouring.py:131-163.  `solid` is accepted for signature parity; the kernel reads the engine's packed flags
        (same mask).  main.py:778 passes four arguments, which raises TypeError there as well and is swallowed by its
        try/except."""
        if self.pouring_active[None] != 1:
            return
        self.pour_time[None] = self.pour_time[None] + np.float32(dt)
        self._engine(lbm_body_force, solid).pouring_force(self._pour_struct(dt), self._tensor(lbm_body_force))

    def apply_gradual_phase_change(self, multiphase_phi, solid, dt: float) -> None:
        """precise_pouring.py:165-196."""
        if self.pouring_active[None] != 1:
            return
        self._engine(solid, multiphase_phi).pouring_phase_change(self._pour_struct(dt), self._tensor(multiphase_phi))

    def create_water_impact_force(self, particle_system, max_force: float, dt: float) -> None:
        raise NotImplementedError("create_water_impact_force (precise_pouring.py:196-233) is not on main.py's step path and is not built")

    def adjust_flow_rate(self, new_rate) -> None:
        self.pour_flow_rate[None] = max(0.1, min(3.0, new_rate))

    def switch_to_spiral_pour(self, radius=10.0, speed=1.0) -> None:
        if self.pouring_active[None]:
            self.pour_pattern[None] = 1
            self.spiral_radius[None] = radius; self.spiral_speed[None] = speed

    def move_pour_center(self, new_x, new_y) -> None:
        self.pour_center_x[None] = max(5, min(self.config.NX - 5, new_x))
        self.pour_center_y[None] = max(5, min(self.config.NY - 5, new_y))

    def get_pouring_info(self) -> Dict[str, Any]:
        """precise_pouring.py:235-274."""
        if self.pouring_active[None] != 1:
            return {"active": False, "position": (0, 0), "diameter_grid": 0, "diameter_cm": 0, "velocity": 0, "flow_rate": 0,
                    "pour_time": 0, "pattern": 0}
        x, y = self._get_current_pour_position()
        return {"active": True, "position": (float(x), float(y)), "diameter_grid": float(self.POUR_DIAMETER_GRID),
                "diameter_cm": float(self.POUR_DIAMETER_CM), "velocity": float(self.POUR_VELOCITY),
                "flow_rate": float(self.pour_flow_rate[None]), "pour_time": float(self.pour_time[None]),
                "pattern": int(self.pour_pattern[None])}

    def get_current_flow_rate(self) -> float:
        if self.pouring_active[None] == 0:
            return 0.0
        cfg = self.config
        return float(self.pour_flow_rate[None] * cfg.INLET_VELOCITY * (cfg.INLET_AREA / cfg.SCALE_LENGTH ** 2))

    def get_current_flow_rate_ml_s(self) -> float:
        return float(self.config.POUR_RATE_ML_S * max(0.0, self.pour_flow_rate[None]))

    def _check_pouring_conditions(self) -> Dict[str, Any]:
        """precise_pouring.py:314-349."""
        cfg = self.config
        r, pz, h = self.POUR_DIAMETER_GRID / 2.0, self.POUR_HEIGHT, 4.0
        cx, cy = cfg.NX // 2, cfg.NY // 2
        ii = np.arange(max(0, int(cx - r)), min(cfg.NX, int(cx + r + 1)))
        jj = np.arange(max(0, int(cy - r)), min(cfg.NY, int(cy + r + 1)))
        kk = np.arange(max(0, int(pz - h)), min(cfg.NZ, int(pz + 1)))
        d = np.sqrt((ii[:, None] - cx) ** 2.0 + (jj[None, :] - cy) ** 2.0)
        inside = int((d <= r).sum()) * int(((kk <= pz) & (kk >= pz - h)).sum())
        total = ii.size * jj.size * kk.size
        return {"center_position": (cx, cy), "pour_radius": r, "z_range": [pz - h, pz], "affected_cells": inside,
                "total_checked": total, "effectiveness": inside / max(1, total)}

    def get_pouring_diagnostics(self) -> Dict[str, Any]:
        return {"configuration": {"diameter_cm": self.POUR_DIAMETER_CM, "diameter_grid": self.POUR_DIAMETER_GRID,
                                  "height": self.POUR_HEIGHT, "velocity": self.POUR_VELOCITY, "grid_size_cm": self.config.GRID_SIZE_CM},
                "current_state": self.get_pouring_info(), "conditions_check": self._check_pouring_conditions()}
