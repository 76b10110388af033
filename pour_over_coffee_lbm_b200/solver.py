"""Solver facades: the reference's API surface over the B200 engine.

* `LBMSolver`          mirrors src/core/legacy/lbm_solver.py:40 (re-exported by src/core/lbm_solver.py:13-18):
                       same field names (f, f_new, rho, u, ux, uy, uz, u_sq, phase, solid, body_force, les_mask,
                       opposite_dir, boundary_manager ...) and methods (step, init_fields, clear_body_force,
                       step_with_particles, step_with_two_way_coupling, add_particle_reaction_forces ...).
* `UnifiedLBMSolver`   mirrors src/core/lbm_unified.py:34 (step() -> backend.execute_collision_streaming).
Both satisfy `LBMSolverProtocol` (src/core/lbm_protocol.py:12-193).

The defaults reproduce the reference: compat="reference", open faces, V60-capable flag field,
body force + phase + lagged-FD LES + filter damping -- all inside ONE kernel launch per step.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch

from . import config as cfgmod
from .config import LBMConfig
from .engine import D3Q19Engine, ParticleState, particles_couple
from .fields import ComponentField, ConstField, PopulationField, ScalarField, VectorField
from .physics import BoundaryConditionManager, LESTurbulenceModel


class LBMSolver:
    def __init__(self, nx: Optional[int] = None, ny: Optional[int] = None, nz: Optional[int] = None, *,
                 config: Optional[LBMConfig] = None, compat: str = "reference", periodic=(False, False, False),
                 geometry: bool = True, force: bool = True, phase: bool = True, les: Optional[bool] = None,
                 porous: Optional[bool] = None, strict: bool = True, device: int = 0, **engine_kw):
        cfg = config or LBMConfig(NX=nx or cfgmod.DEFAULT.NX, NY=ny or cfgmod.DEFAULT.NY,
                                  NZ=engine_kw.get("nz_global") or nz or cfgmod.DEFAULT.NZ)
        self.config = cfg
        nx, ny, nz = nx or cfg.NX, ny or cfg.NY, nz or cfg.NZ
        self.use_les = cfg.use_les if les is None else bool(les)          # legacy/lbm_solver.py:94
        walls = bool(geometry) or not all(periodic)
        porous = walls if porous is None else porous
        self.engine = D3Q19Engine(nx, ny, nz, compat=compat, periodic=periodic, walls=walls, force=force, phase=phase,
                                  les=self.use_les, porous=porous and walls, strict=strict, config=cfg, device=device,
                                  **engine_kw)
        e = self.engine
        zg = e.zghost
        self._flags_dirty = False
        dirty = self._mark_flags_dirty
        # ---- fields (LBMSolverProtocol + de-facto surface, main.py:378-442) -----------------------
        self.f = PopulationField(e)
        self.f_new = self.f                    # double buffering is internal (pointer swap); same view
        self.rho = ScalarField(lambda: e.rho, zg)
        self.u = VectorField(lambda: e.u, zg)
        self.ux = ComponentField(lambda: e.u, 0, zg)
        self.uy = ComponentField(lambda: e.u, 1, zg)
        self.uz = ComponentField(lambda: e.u, 2, zg)
        self.phase = ScalarField(lambda: e.phase, zg) if e.phase is not None else None
        self.body_force = VectorField(lambda: e.body_force, zg) if e.body_force is not None else None
        self.solid = ScalarField(lambda: e.solid, zg, on_write=dirty) if walls else None
        self.les_mask = ScalarField(lambda: e.les_mask, zg, on_write=dirty) if walls else None
        self.filter_zone = ScalarField(lambda: e.filter_zone, zg, on_write=dirty) if walls else None
        self.cx, self.cy, self.cz = ConstField(cfgmod.CX_3D), ConstField(cfgmod.CY_3D), ConstField(cfgmod.CZ_3D)
        self.w = ConstField(cfgmod.WEIGHTS_3D)
        self.e = ConstField(np.stack([cfgmod.CX_3D, cfgmod.CY_3D, cfgmod.CZ_3D], axis=1))
        self.opposite_dir = ConstField(cfgmod.OPPOSITE_3D)
        for fld in (self.f, self.rho, self.u, self.ux, self.uy, self.uz, self.phase, self.body_force, self.solid, self.les_mask, self.filter_zone):
            if fld is not None:
                fld.owner = self               # a module handed only fields (PrecisePouringSystem, main.py:778-780) finds the engine
        self.les_model = LESTurbulenceModel(self) if self.use_les else None
        self.boundary_manager = BoundaryConditionManager()
        self.memory_adapter = self
        self.step_count = 0
        self.layout_type = "SoA[q][z][y][x]"

    # ---- flags -------------------------------------------------------------------------------------
    def _mark_flags_dirty(self):
        self._flags_dirty = True

    def _sync_flags(self):
        """solid / les_mask / filter_zone were written through the field surface: repack the flag byte.
        After the first step the populations are converted g -> f (old mask) -> g (new mask) so the
        trajectory matches the reference, which streamed with the mask current at collision time."""
        if not self._flags_dirty:
            return
        self._flags_dirty = False
        e = self.engine
        if e.zghost and e.nranks > 1:
            # the field surface writes owned planes only: the neighbours' boundary planes of the masks belong in the ghost planes
            # before NEAR flags, neighbour masks and wall links are derived from them
            from . import slab
            for t in (e.solid, e.filter_zone, e.les_mask):
                slab.exchange_planes(t, e.rank, e.nranks, e.periodic[2])
        if e.steps_done > 0 and e.compat_name == "reference":
            e.set_geometry_preserving_f(lambda: None)
        else:
            e.pack_flags()

    # ---- u_sq is derived on demand (legacy/lbm_solver.py:333) ---------------------------------------
    @property
    def u_sq(self):
        u = self.engine.u
        sq = (u[0] * u[0] + u[1] * u[1]) + u[2] * u[2]
        return ScalarField(lambda: sq, self.engine.zghost)

    u_sqr = u_sq

    # ---- initialisation ---------------------------------------------------------------------------
    def init_fields(self) -> None:
        """legacy/lbm_solver.py:1067-1112"""
        e = self.engine
        if e.phase is not None: e.phase.zero_()
        if e.body_force is not None: e.body_force.zero_()
        self._sync_flags()
        e.init_equilibrium(1.0, (0.0, 0.0, 0.0))

    def initialize_fields(self, initial_density: float = 1.0, initial_velocity=(0.0, 0.0, 0.0)) -> None:
        """LBMSolverProtocol.initialize_fields, lbm_protocol.py:89"""
        self._sync_flags()
        self.engine.init_equilibrium(float(initial_density), tuple(initial_velocity))

    def reset_solver(self) -> None:
        self.init_fields()
        self.step_count = 0

    def set_geometry(self, geometry_function: Any) -> None:
        """LBMSolverProtocol.set_geometry: callable(i,j,k arrays) -> bool solid, or an array-like mask."""
        if callable(geometry_function):
            i, j, k = np.meshgrid(np.arange(self.engine.nx), np.arange(self.engine.ny),
                                  np.arange(self.engine.z0, self.engine.z0 + self.engine.nz), indexing="ij")
            mask = np.asarray(geometry_function(i, j, k))
        else:
            mask = np.asarray(geometry_function)
        self.solid.from_numpy(mask.astype(np.uint8))

    # ---- the step ---------------------------------------------------------------------------------
    def step(self) -> None:
        """legacy/lbm_solver.py:817-867: LES pre-pass + macroscopic + collide/stream + swap + filter
        damping are ONE fused kernel; the boundary manager then writes the open-face densities."""
        self._sync_flags()
        self.engine.step(1, write_macro_every=1)
        self.boundary_manager.apply_all_boundaries(self)
        self.step_count += 1

    def run(self, nsteps: int, write_macro_every: int = 0) -> None:
        """Many steps without returning to Python in between (no reference equivalent; used by benchmarks)."""
        self._sync_flags()
        if self.engine.ref_les:
            write_macro_every = 1
        self.engine.step(nsteps, write_macro_every=write_macro_every)
        self.step_count += nsteps

    _collision_streaming_step = lambda self: (self._sync_flags(), self.engine.step(1, write_macro_every=1))[-1]

    def collision_step(self) -> None:        # the fused kernel does both halves
        self.step()

    def streaming_step(self) -> None:
        pass

    def swap_fields(self) -> None:           # pointer swap happens inside lbm_step
        pass

    def compute_macroscopic_quantities(self) -> None:
        self._sync_flags()
        self.engine.macroscopic()

    _compute_macroscopic_quantities = compute_macroscopic_quantities

    def apply_boundary_conditions(self) -> None:
        self.boundary_manager.apply_all_boundaries(self)

    def clear_body_force(self) -> None:
        self.engine.clear_body_force()

    def add_force_term(self, force_field) -> None:
        t = force_field.to_torch() if hasattr(force_field, "to_torch") else torch.as_tensor(np.asarray(force_field))
        self.body_force.view().add_(t.to(self.engine.device, torch.float32))

    def enable_les_turbulence(self, smagorinsky_constant: float = 0.1) -> None:
        if not self.use_les:
            raise RuntimeError("construct the solver with les=True (the LES variant is a different kernel instantiation)")
        self.engine.set_params(cs_smag=float(smagorinsky_constant))

    # ---- particles (legacy/lbm_solver.py:1478-1509) -------------------------------------------------
    def add_particle_reaction_forces(self, particle_system) -> None:
        self.engine.add_reaction_force(particle_system.reaction_force_tensor)

    def step_with_two_way_coupling(self, particle_system, dt: float = 1.0, relaxation_factor: float = 0.8) -> None:
        self.clear_body_force()
        if particle_system:
            particle_system.compute_two_way_coupling_forces(self.u, relax=relaxation_factor)
            self.add_particle_reaction_forces(particle_system)
        self.step()

    def step_with_particles(self, particle_system) -> None:
        self.step()
        if particle_system is not None and hasattr(particle_system, "update"):
            particle_system.update(self)

    def get_coupling_diagnostics(self, particle_system=None) -> Dict[str, Any]:
        bf = self.engine.body_force
        fluid = (self.engine.solid == 0) if self.engine.solid is not None else torch.ones_like(bf[0], dtype=torch.bool)
        mag = torch.sqrt(bf[0] ** 2 + bf[1] ** 2 + bf[2] ** 2)
        d = {"lbm_step_count": self.step_count,
             "body_force_magnitude": float((mag * fluid).sum() / max(1, int(fluid.sum())))}
        if particle_system is not None and hasattr(particle_system, "get_coupling_diagnostics"):
            d["particle_coupling"] = particle_system.get_coupling_diagnostics()
        return d

    # ---- read-outs ---------------------------------------------------------------------------------
    def get_velocity_magnitude(self) -> np.ndarray:
        u = self.u.to_numpy()
        return np.sqrt(u[..., 0] ** 2 + u[..., 1] ** 2 + u[..., 2] ** 2)

    def get_velocity_vector_field(self):
        return self.u

    # velocity-layout helpers of the legacy class (legacy/lbm_solver.py:1114-1232): one device layout here, both views live
    def get_velocity_vector(self):
        return self.u

    def get_velocity_components(self):
        return self.ux, self.uy, self.uz

    def set_velocity_vector(self, u_field) -> None:
        self.u.from_numpy(u_field.to_numpy() if hasattr(u_field, "to_numpy") else np.asarray(u_field, np.float32))

    def has_soa_velocity_layout(self) -> bool:
        return True

    def sync_soa_to_vector_velocity(self) -> None:      # ux/uy/uz and u are views of the same tensor
        pass

    sync_vector_to_soa_velocity = sync_soa_to_vector_velocity

    def get_solver_type(self) -> str:
        return "b200"

    # No step_ultra_optimized / step_with_cfl_control / collide / stream here: main.py:803-824 probes those names with hasattr and
    # the reference's LBMSolver answers "absent" (tests/golden/reference_main_trace.json, "probes"), which sends main.py down
    # step_with_particles(particle_system) -- the path the recorded run took and tests/test_main_trace.py replays.

    get_velocity_field_for_thermal_coupling = get_velocity_vector_field

    def field_statistics(self) -> torch.Tensor:
        """lbm_field_statistics: [max|u|, min rho, max rho, sum rho, kinetic energy, NaN count, Inf count, fluid cells] of the
        owned slab's fluid cells, f64 device tensor, one fused pass (replaces visualizer.get_statistics,
        visualizer.py:169-183, and NumericalStabilityMonitor.check_field_stability, numerical_stability.py:52-110)."""
        return self.engine.field_statistics()

    def step_statistics(self) -> torch.Tensor:
        """[max |u|, mean rho over the fluid cells] as a 2-element device tensor (main.py:907-912), from the fused pass."""
        s = self.field_statistics()
        return torch.stack([s[0], s[3] / torch.clamp(s[7] - s[5] - s[6], min=1.0)]).float()

    def get_kinetic_energy(self) -> float:
        return float(self.field_statistics()[4])

    def get_mass_conservation_error(self) -> float:
        e = self.engine
        zs = slice(e.zghost, e.zghost + e.nz)
        return float(abs(e.rho[zs].double().mean() - 1.0))

    def check_stability(self) -> bool:
        """numerical_stability.py:52-110 reduced to its verdict: no NaN / Inf in rho, u and |u| below 0.3 lu."""
        s = self.field_statistics().tolist()
        return bool(s[5] == 0 and s[6] == 0 and s[0] < 0.3)

    def get_diagnostics(self) -> dict:
        s = self.step_statistics().tolist()
        return {"step": self.step_count, "max_velocity": s[0], "mean_density": s[1],
                "kinetic_energy": self.get_kinetic_energy(), "kernel_launches": self.engine.launch_count()}

    def get_memory_usage(self) -> dict:
        e = self.engine
        tensors = [*e.g, e.rho, *e.u_buf, e.body_force, e.phase, e.flags, e.solid, e.filter_zone, e.les_mask]
        total = sum(t.numel() * t.element_size() for t in tensors if t is not None)
        return {"total_bytes": total, "total_gb": total / 1e9, "populations_gb": 2 * e.g[0].numel() * 4 / 1e9}

    def optimize_memory_layout(self) -> None:
        pass        # the layout is fixed: SoA, x fastest, 128-bit aligned rows

    def export_vtk(self, filename: str) -> None:
        rho, u = self.rho.to_numpy(), self.u.to_numpy()
        nx, ny, nz = rho.shape
        with open(filename, "w") as fh:
            fh.write("# vtk DataFile Version 3.0\nlbm_b200\nASCII\nDATASET STRUCTURED_POINTS\n")
            fh.write(f"DIMENSIONS {nx} {ny} {nz}\nORIGIN 0 0 0\nSPACING 1 1 1\nPOINT_DATA {nx * ny * nz}\n")
            fh.write("SCALARS rho float 1\nLOOKUP_TABLE default\n")
            np.savetxt(fh, rho.transpose(2, 1, 0).ravel(), fmt="%.7g")
            fh.write("VECTORS u float\n")
            np.savetxt(fh, u.transpose(2, 1, 0, 3).reshape(-1, 3), fmt="%.7g")


class UnifiedLBMSolver:
    """src/core/lbm_unified.py:34-260.  The reference auto-selects Apple/CUDA/CPU Taichi backends;
    here there is exactly one backend (B200) and no fallback chain."""

    def __init__(self, preferred_backend: Optional[str] = None, **solver_kw):
        from .backend import B200Backend
        if preferred_backend not in (None, "auto", "b200", "cuda"):
            raise RuntimeError(f"backend '{preferred_backend}' does not exist in this build: B200 (sm_100a) only")
        self.backend = B200Backend()
        solver_kw.setdefault("compat", "reference")
        self._solver = LBMSolver(**solver_kw)
        self.backend.bind(self._solver)
        s = self._solver
        self.memory_adapter = s            # object with .f .f_new .rho .u .solid .phase (lbm_unified.py:190-198)
        self.f, self.f_new, self.rho, self.u = s.f, s.f_new, s.rho, s.u
        self.solid, self.phase = s.solid, s.phase
        self.body_force, self.boundary_manager = s.body_force, s.boundary_manager
        self.tau = s.config.TAU_WATER
        self.dt = s.config.DT
        self.Re = s.config.RE_CHAR

    def step(self):
        params = {"tau": self.tau, "dt": self.dt, "Reynolds": self.Re}
        self.backend.execute_collision_streaming(self.memory_adapter, params)

    def initialize_fields(self):
        self._solver.initialize_fields(self._solver.config.RHO_0)

    init_fields = initialize_fields

    def __getattr__(self, name):          # everything else is the LBMSolver surface
        return getattr(self.__dict__["_solver"], name)

    def get_solver_info(self) -> Dict[str, Any]:
        e = self._solver.engine
        return {"name": "Unified LBM Solver", "version": "B200", "backend": self.backend.get_backend_info(),
                "memory_adapter": "SoA[q][z][y][x]", "grid_size": (e.nx, e.ny, e.nz),
                "memory_usage_gb": self._solver.get_memory_usage()["total_gb"],
                "physics_params": {"tau": self.tau, "dt": self.dt, "Reynolds": self.Re}}

    def run_diagnostic(self) -> Dict[str, Any]:
        return {"backend_status": self.backend.get_backend_info(),
                "memory_status": self._solver.get_memory_usage(),
                "field_status": {"total_fields": 19 + 4, "grid_points": self._solver.engine.cells()}}


def create_unified_solver(preferred_backend: Optional[str] = None, **kw) -> UnifiedLBMSolver:
    return UnifiedLBMSolver(preferred_backend, **kw)
