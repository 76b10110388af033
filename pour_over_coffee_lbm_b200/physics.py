"""Facades of the physics modules whose work is fused into (or feeds) the D3Q19 step.

Names and call signatures follow the reference so main.py-style orchestration keeps working:
  BoundaryConditionManager   src/physics/boundary_conditions.py:326-556
  LESTurbulenceModel         src/physics/les_turbulence.py:40-389
  FilterPaperSystem          src/physics/filter_paper.py:47-926          (geometry + drag only)
  PressureGradientDrive      src/physics/pressure_gradient_drive.py:14-397 (force mode B / mixed)
  CoffeeParticleSystem       src/physics/coffee_particles.py:14-1245      (two-way coupling only)
All device work goes through the C ABI; nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch

from .engine import ParticleState, particles_couple, _ptr
from .fields import ScalarField, VectorField


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


# ----------------------------------------------------------------------------------------------------
class PressureGradientDrive:
    def __init__(self, lbm_solver: Any):
        self.lbm = lbm_solver
        self.MAX_PRESSURE_FORCE = 0.12           # pressure_gradient_drive.py:30 (host-mutable, main.py:913-925)
        self.force_drive_active = False
        self.mixed_drive_active = False
        self.density_drive_active = False

    def activate_force_drive(self, active: bool = True):
        self.force_drive_active = bool(active)
        if active: self.mixed_drive_active = False

    def activate_mixed_drive(self, active: bool = True):
        self.mixed_drive_active = bool(active)
        if active: self.force_drive_active = False

    def activate_density_drive(self, active: bool = True):
        # method A writes rho, which the next macroscopic pass overwrites: inert in LBMSolver
        self.density_drive_active = bool(active)

    def apply_force_drive(self):
        if self.force_drive_active:
            self.lbm.engine.add_pressure_gradient_force(self.MAX_PRESSURE_FORCE, 1.0)

    def apply(self, step: int = 0):
        """pressure_gradient_drive.py:257-272"""
        if self.force_drive_active:
            self.lbm.engine.add_pressure_gradient_force(self.MAX_PRESSURE_FORCE, 1.0)
        elif self.mixed_drive_active:
            self.lbm.engine.add_pressure_gradient_force(self.MAX_PRESSURE_FORCE, 0.5)


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
        return t.t()

    position = property(lambda self: self._pv(self.state.pos))
    velocity = property(lambda self: self._pv(self.state.vel))
    drag_force = property(lambda self: self._pv(self.state.drag))
    drag_force_new = property(lambda self: self._pv(self.state.drag_new))
    drag_force_old = property(lambda self: self._pv(self.state.drag_old))
    fluid_velocity_at_particle = property(lambda self: self._pv(self.state.u_fluid))
    radius = property(lambda self: self.state.radius)
    mass = property(lambda self: self.state.mass)
    active = property(lambda self: self.state.active)
    particle_reynolds = property(lambda self: self.state.reynolds)
    drag_coefficient = property(lambda self: self.state.cd)
    cell_index = property(lambda self: self._pv(self.state.cell))

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

    def compute_two_way_coupling_forces(self, fluid_u=None, relax: float = -1.0):
        """coffee_particles.py:1107-1154; relax >= 0 also applies the under-relaxation in the same kernel."""
        particles_couple(self._solver.engine, self.state, self.reaction_force_tensor, relax=relax,
                         water_density=self.water_density, water_viscosity=self.water_viscosity)

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
