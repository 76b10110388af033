"""GPU parity of geometry/flag building, particle coupling and the small force kernels vs the oracle."""
import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R
from oracle import ref_cpu as RC

pytestmark = pytest.mark.gpu


def _engine(*a, **k):
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine
    return D3Q19Engine(*a, **k)


def _torch(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- solid / filter-zone flags: bit-exact ------------------------------------------
@pytest.mark.parametrize("n", [64, 224, 256])
def test_v60_solid_and_filter_zone_bit_exact(n):
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    solid_ref = RC.v60_solid(cfg) if n > 64 else R.v60_solid(cfg)
    zone_ref = R.filter_zones(cfg)
    eng = _engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, macro_fields=False)
    eng.build_v60_geometry()
    solid = H.from_dev_scalar(eng.solid); zone = H.from_dev_scalar(eng.filter_zone)
    assert solid.dtype == np.uint8
    assert int((solid != solid_ref).sum()) == 0
    assert int((zone != zone_ref).sum()) == 0
    les = H.from_dev_scalar(eng.les_mask)
    assert np.array_equal(les, np.where(zone_ref == 1, 0, 1))
    if n == 224:      # fluid fraction of the default V60 mask (~35.3 %)
        assert abs((solid == 0).mean() - 0.3535) < 5e-4


def test_flag_byte_packing_and_near_bit():
    from pour_over_coffee_lbm_b200 import _lib as L
    n = 40
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    solid = R.v60_solid(cfg); zone = R.filter_zones(cfg)
    eng = _engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, macro_fields=False)
    eng.build_v60_geometry()
    flags = H.from_dev_scalar(eng.flags)
    assert np.array_equal((flags & L.FLAG_SOLID) != 0, solid != 0)
    assert np.array_equal((flags & L.FLAG_FILTER) != 0, zone == 1)
    assert np.array_equal((flags & L.FLAG_LES) != 0, zone != 1)
    # NEAR = any of the 18 neighbours solid or outside the (open) box
    pad = np.pad(solid, 1, constant_values=1)
    near = np.zeros(solid.shape, bool)
    for q in range(1, 19):
        ex, ey, ez = int(R.CX[q]), int(R.CY[q]), int(R.CZ[q])
        near |= pad[1 - ex:1 - ex + n, 1 - ey:1 - ey + n, 1 - ez:1 - ez + n] != 0
    assert np.array_equal((flags & L.FLAG_NEAR) != 0, near)


# ---- f <-> g conversion is exact data movement -----------------------------------------------------
def test_export_import_f_roundtrip_and_oracle():
    import torch
    n = 24
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    solid = R.v60_solid(cfg)
    eng = _engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, macro_fields=False)
    eng.solid.copy_(_torch(H.to_dev_scalar(solid))); eng.pack_flags()
    rng = np.random.default_rng(0)
    g = rng.random((19, n, n, n), dtype=np.float32)
    eng.g[eng.cur].copy_(_torch(H.to_dev_pop(g)))
    f = H.from_dev_pop(eng.export_f())
    f_ref = R.stream_from_post_collision(g, solid)
    fluid = solid == 0
    assert np.array_equal(f[:, fluid], f_ref[:, fluid])
    # import(export(g)) reproduces every population that can ever be read again
    eng.import_f(torch.from_numpy(H.to_dev_pop(f)))
    f2 = H.from_dev_pop(eng.export_f())
    assert np.array_equal(f2[:, fluid], f[:, fluid])


# ---- particles -----------------------------------------------------------------------------------------
# tests/test_trilinear_interpolation.py:24-38,143-159 of the reference: v=(x,y,z) on 16^3, ten positions
# must interpolate to themselves.
TRILINEAR_POSITIONS = [
    (5.0, 5.0, 5.0), (5.5, 5.5, 5.5), (5.25, 6.75, 7.1), (0.1, 0.1, 0.1), (14.9, 14.9, 14.9),
    (2.3, 8.7, 11.2), (7.8, 3.4, 9.6), (12.1, 13.5, 4.8), (1.7, 5.9, 14.3), (8.4, 11.2, 6.7)]


def _particle_solver(n):
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    from pour_over_coffee_lbm_b200.physics import CoffeeParticleSystem
    s = LBMSolver(nx=n, ny=n, nz=n, compat="reference", strict=True)
    ps = CoffeeParticleSystem(max_particles=4096, solver=s)
    return s, ps


def test_trilinear_known_answers():
    n = 16
    s, ps = _particle_solver(n)
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    s.u.from_numpy(np.stack([i, j, k], axis=-1).astype(np.float32))
    pos = np.array(TRILINEAR_POSITIONS, np.float32)
    ps.set_particles(pos)
    ps.compute_two_way_coupling_forces(s.u)
    got = ps.fluid_velocity_at_particle[: len(pos)].cpu().numpy()
    assert np.abs(got - pos).max() < 5e-6        # f32 (reference reports 4.8e-7 relative on this table)


def test_particle_cell_indices_bit_exact_and_forces():
    n, P = 48, 3000
    s, ps = _particle_solver(n)
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    rng = np.random.default_rng(42)
    u = H.smooth_velocity(n, 0.05, 3)
    s.u.from_numpy(u)
    pos = rng.uniform(-1.0, n + 1.0, size=(P, 3)).astype(np.float32)      # includes out-of-range positions (clamped)
    pos[:64] = np.round(pos[:64])                                         # exact-integer positions (truncation edge)
    pos[64:128] = pos[0]                                                   # many particles in ONE cell (warp aggregation)
    vel = (0.01 * rng.standard_normal((P, 3))).astype(np.float32)
    radius = np.clip(rng.normal(3.25e-4, 0.3 * 3.25e-4, P), 0.5 * 3.25e-4, 1.5 * 3.25e-4).astype(np.float32)
    mass = (np.float32(1200.0) * np.float32(4.0 / 3.0 * np.pi) * radius ** 3).astype(np.float32)
    ps.set_particles(pos, vel, radius, mass)
    ps.state.active[P - 10:P] = 0                                          # inactive tail
    active = np.ones(P, np.int32); active[P - 10:] = 0
    ps.state.drag_old[:, :P] = _torch(np.full((3, P), 1e-9, np.float32))
    ps.compute_two_way_coupling_forces(s.u, relax=0.8)
    drag_new, react, u_fl, re_p, cd, cell = R.two_way_coupling(cfg, u, pos, vel, radius, mass, active)
    act = active != 0
    # particle cell indices: bit-exact
    assert np.array_equal(ps.cell_index[:P].cpu().numpy()[act], cell[act])
    # gather: same operation order -> bit-exact
    assert np.array_equal(ps.fluid_velocity_at_particle[:P].cpu().numpy()[act], u_fl[act])
    # Reynolds exact; C_D / drag carry powf (<= 2 ulp)
    assert np.array_equal(ps.particle_reynolds[:P].cpu().numpy()[act], re_p[act])
    np.testing.assert_allclose(ps.drag_coefficient[:P].cpu().numpy()[act], cd[act], rtol=5e-7)
    np.testing.assert_allclose(ps.drag_force_new[:P].cpu().numpy()[act], drag_new[act], rtol=1e-6, atol=1e-20)
    # scatter: atomics are unordered -> tolerance scaled by the largest nodal force
    got = ps.reaction_force_field.to_numpy()
    assert np.abs(got - react).max() <= 1e-5 * np.abs(react).max()
    # momentum conservation of the scatter: sum of nodal forces = - sum of particle drags
    np.testing.assert_allclose(got.reshape(-1, 3).sum(0, dtype=np.float64), -drag_new[act].sum(0, dtype=np.float64), rtol=1e-4)
    # under-relaxation F = a F_new + (1-a) F_old
    drag, new_old = R.under_relax(drag_new, np.full((P, 3), 1e-9, np.float32), active, 0.8)
    np.testing.assert_allclose(ps.drag_force[:P].cpu().numpy()[act], drag[act], rtol=1e-6, atol=1e-20)
    # add_particle_reaction_forces: body_force += reaction on fluid cells only
    s.clear_body_force(); s.add_particle_reaction_forces(ps)
    np.testing.assert_array_equal(s.body_force.to_numpy(), got)


def test_particle_integrator_bit_exact():
    """lbm_particles_advance == CoffeeParticleSystem.update_particle_physics restated in the oracle
    (coffee_particles.py:641-831): free flight, acceleration and displacement caps, cone / bottom / top constraints
    with damping, invalid coordinates (deactivation), invalid velocities, massless and inactive particles; three
    consecutive steps so constrained particles move again."""
    import torch
    from pour_over_coffee_lbm_b200.config import LBMConfig
    from pour_over_coffee_lbm_b200.engine import ParticleState, particles_advance
    n = 64
    cfg = LBMConfig(NX=n, NY=n, NZ=n)
    rcfg = R.RefConfig(NX=n, NY=n, NZ=n)
    eng = _engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, config=cfg)
    P = 1500
    rng = np.random.default_rng(8)
    bounds = dict(center_x=n // 2, center_y=n // 2, bottom_z=n // 4, bottom_radius_lu=rcfg.BOTTOM_RADIUS / rcfg.SCALE_LENGTH,
                  top_radius_lu=rcfg.TOP_RADIUS / rcfg.SCALE_LENGTH)
    pos = np.stack([rng.uniform(-2, n + 2, P), rng.uniform(-2, n + 2, P), rng.uniform(-2, n + 2, P)], 1).astype(np.float32)
    vel = (rng.standard_normal((P, 3)) * rng.choice([0.05, 5.0, 40.0, 400.0], (P, 1))).astype(np.float32)
    force = (rng.standard_normal((P, 3)) * rng.choice([1e-9, 1e-6, 1e-2], (P, 1))).astype(np.float32)
    mass = rng.choice([0.0, 1e-11, 1.7e-7, 3e-7], P).astype(np.float32)
    active = (rng.random(P) < 0.9).astype(np.int32)
    pos[:5] = np.nan; vel[5:10, 1] = np.nan; pos[10:15, 2] = np.inf
    ps = ParticleState(P, eng.device)
    ps.pos.copy_(_torch(pos.T)); ps.vel.copy_(_torch(vel.T)); ps.mass.copy_(_torch(mass)); ps.active.copy_(_torch(active))
    f_dev = _torch(force.T).contiguous()
    counters = torch.zeros(2, dtype=torch.int32, device="cuda")
    tot = [0, 0]
    for step, dt in enumerate((5e-3, 1.0, 1e-12)):      # the last two exercise the dt clamp
        ce, bv = R.update_particle_physics(rcfg, pos, vel, force, mass, active, dt, **bounds)
        tot[0] += ce; tot[1] += bv
        particles_advance(eng, ps, dt, force=f_dev, counters=counters, **bounds)
        got_pos = ps.pos.cpu().numpy().T; got_vel = ps.vel.cpu().numpy().T
        assert np.array_equal(ps.active.cpu().numpy(), active)
        act = active == 1
        assert np.array_equal(got_pos[act], pos[act]) and np.array_equal(got_vel[act], vel[act], equal_nan=True)   # massless particles keep a NaN velocity
        assert np.array_equal(f_dev.cpu().numpy().T, force, equal_nan=True)
        force[:] = (rng.standard_normal((P, 3)) * 1e-6).astype(np.float32); f_dev.copy_(_torch(force.T))
    assert counters.cpu().tolist() == tot
    assert tot[1] > 0 and tot[0] > 0


def test_empty_and_single_particle():
    n = 16
    s, ps = _particle_solver(n)
    ps.compute_two_way_coupling_forces(s.u, relax=0.8)              # zero active particles
    assert float(ps.reaction_force_tensor.abs().max()) == 0.0
    s.u.fill([0.01, 0.0, 0.0])
    ps.set_particles(np.array([[8.25, 8.5, 8.75]], np.float32))
    ps.compute_two_way_coupling_forces(s.u, relax=0.8)
    r = ps.reaction_force_field.to_numpy()
    assert (r[..., 0] <= 0).all() and r[..., 0].min() < 0          # reaction opposes the drag (+x)
    assert np.count_nonzero(r[..., 0]) == 8


# ---- neighbours that feed body_force -------------------------------------------------------------------
def test_pressure_gradient_and_forchheimer_force_bit_exact():
    n = 32
    st = H.reference_v60_state(n, seed=5)
    rng = np.random.default_rng(1)
    st.rho = (1.0 + 0.05 * rng.standard_normal(st.rho.shape)).astype(np.float32)
    st.u = (0.02 * rng.standard_normal(st.u.shape)).astype(np.float32)
    from pour_over_coffee_lbm_b200.config import LBMConfig
    eng = _engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, force=True, phase=True, porous=True,
                  config=LBMConfig(NX=n, NY=n, NZ=n))
    eng.solid.copy_(_torch(H.to_dev_scalar(st.solid))); eng.filter_zone.copy_(_torch(H.to_dev_scalar(st.filter_zone)))
    eng.pack_flags()
    eng.rho.copy_(_torch(H.to_dev_scalar(st.rho))); eng.u.copy_(_torch(H.to_dev_vec(st.u)))
    # PressureGradientDrive force mode and mixed (0.5x) mode
    for scale in (1.0, 0.5):
        st.body_force[:] = 0; eng.clear_body_force()
        pf = R.pressure_gradient_force(st, 0.12)
        R.accumulate_pressure_force(st, pf, scale)
        eng.add_pressure_gradient_force(0.12, scale)
        assert np.array_equal(H.from_dev_vec(eng.body_force), st.body_force)
        # written instead of accumulated (no clear pass): same values on fluid cells, solid cells untouched
        keep = eng.body_force.clone()
        eng.body_force.fill_(7.0)
        eng.set_pressure_gradient_force(0.12, scale)
        got = H.from_dev_vec(eng.body_force); fluid = st.solid == 0
        assert np.array_equal(got[fluid], st.body_force[fluid]) and np.all(got[~fluid] == 7.0)
        eng.body_force.copy_(keep)
    # FilterPaperSystem.compute_forchheimer_resistance on top
    R.compute_forchheimer_resistance(st)
    eng.add_forchheimer_force()
    got = H.from_dev_vec(eng.body_force)
    assert np.array_equal(got, st.body_force)
    zone = (st.filter_zone == 1) & (st.solid == 0)
    assert np.abs(got[zone]).max() > 0          # tests/test_forchheimer.py:31-74: non-zero in the zone


@pytest.mark.parametrize("compat,n", [("reference", 48), ("physical", 48), ("reference", 30)], ids=["dense_scan_vec4", "quad_list", "dense_scan_ragged"])
def test_fused_field_statistics_match_numpy_and_are_deterministic(compat, n):
    """lbm_field_statistics (one pass, device result) against NumPy on a V60 mask with injected NaN / Inf: the dense scans (four cells
    per thread / one cell per thread on a ragged nx) and the scan over the packed quad list of the four-cell walls kernel."""
    import torch
    st = H.reference_v60_state(n, seed=7)
    rng = np.random.default_rng(2)
    rho = (1.0 + 0.05 * rng.standard_normal(st.rho.shape)).astype(np.float32)
    u = (0.03 * rng.standard_normal(st.u.shape)).astype(np.float32)
    fluid = st.solid == 0
    idx = np.argwhere(fluid)
    rho[tuple(idx[3])] = np.nan; rho[tuple(idx[40])] = np.inf; u[tuple(idx[77])][1] = np.nan; u[tuple(idx[90])][2] = -np.inf
    from pour_over_coffee_lbm_b200.config import LBMConfig
    eng = _engine(n, n, n, compat=compat, periodic=(False, False, False), walls=True, config=LBMConfig(NX=n, NY=n, NZ=n))
    eng.solid.copy_(_torch(H.to_dev_scalar(st.solid))); eng.pack_flags()
    eng.rho.copy_(_torch(H.to_dev_scalar(rho))); eng.u.copy_(_torch(H.to_dev_vec(u)))
    a = eng.field_statistics().clone(); b = eng.field_statistics().clone()
    assert torch.equal(a, b)                                           # deterministic, bit for bit
    got = a.cpu().numpy()
    r = rho[fluid]; uf = u[fluid]
    um = np.sqrt((uf[:, 0] * uf[:, 0] + uf[:, 1] * uf[:, 1]) + uf[:, 2] * uf[:, 2])
    rfin = np.isfinite(r); ufin = np.isfinite(um)
    assert got[0] == um[ufin].max() and got[1] == r[rfin].min() and got[2] == r[rfin].max()
    assert np.isclose(got[3], r[rfin].astype(np.float64).sum(), rtol=1e-12)
    ke = 0.5 * r.astype(np.float64) * (uf.astype(np.float64) ** 2).sum(1)
    assert np.isclose(got[4], ke[rfin & ufin].sum(), rtol=1e-12)
    assert got[5] == np.isnan(r).sum() + np.isnan(um).sum() and got[6] == np.isinf(r).sum() + np.isinf(um).sum()
    assert got[7] == fluid.sum()


# ---- facades ----------------------------------------------------------------------------------------------
def test_solver_facade_surface_and_protocol():
    from pour_over_coffee_lbm_b200.solver import LBMSolver, UnifiedLBMSolver
    from pour_over_coffee_lbm_b200.physics import FilterPaperSystem, PressureGradientDrive
    n = 32
    s = LBMSolver(nx=n, ny=n, nz=n)
    for name in ("f", "f_new", "rho", "u", "solid", "phase", "ux", "uy", "uz", "body_force", "boundary_manager", "les_mask",
                 "opposite_dir"):
        assert hasattr(s, name)
    for name in ("step", "collision_step", "streaming_step", "compute_macroscopic_quantities", "apply_boundary_conditions",
                 "initialize_fields", "set_geometry", "get_diagnostics", "check_stability", "get_kinetic_energy",
                 "get_mass_conservation_error", "get_memory_usage", "optimize_memory_layout", "enable_les_turbulence",
                 "add_force_term", "export_vtk", "init_fields", "clear_body_force", "swap_fields",
                 "step_with_two_way_coupling", "get_velocity_magnitude"):
        assert callable(getattr(s, name))
    s.init_fields()
    assert s.f.shape == (19, n, n, n) and s.u.shape == (n, n, n, 3) and s.rho.shape == (n, n, n)
    # init state: f = w_q, rho = 1 (legacy/lbm_solver.py:1067-1112)
    f = s.f.to_numpy()
    assert np.array_equal(f[0], np.full((n, n, n), np.float32(1 / 3)))
    fp = FilterPaperSystem(s); fp.initialize_filter_geometry()
    assert np.array_equal(s.solid.to_numpy(), R.v60_solid(R.RefConfig(NX=n, NY=n, NZ=n)))
    frac = fp.get_filter_statistics()["filter_fraction"]
    assert 0 < frac < 0.5                                   # tests/test_filter_paper.py:38-65
    drive = PressureGradientDrive(s); drive.activate_force_drive(True)
    # phase stays 0 as in the reference's own test (with phase > 0.001 the default GRAVITY_LU = 44.1 enters
    # u = (m + F/2)/rho unclamped -- quirk Q4 -- and |u| exceeds any stability bound by construction)
    for _ in range(3):                                      # tests/test_lbm_solver_unit.py:50-60: (apply; step) x3, no NaN
        s.clear_body_force(); drive.apply(); s.step()
    assert s.check_stability()
    rho = s.rho.to_numpy()
    assert np.isfinite(rho).all() and (rho > 0).all()
    us = UnifiedLBMSolver(nx=n, ny=n, nz=n)
    us.initialize_fields(); us.step(); us.step()
    m = us.backend.get_performance_metrics()
    assert m["throughput_mlups"] > 0 and us.backend.validate_platform()


def test_main_py_orchestration_surface_on_the_device():
    """The calls main.py makes around the step (main.py:560-700, 735-935): filter geometry -> coffee bed creation ->
    pre-stabilisation loop (step + particle integrator with the filter's bounds) -> drive + coupled step, statistics."""
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    from pour_over_coffee_lbm_b200.physics import CoffeeParticleSystem, FilterPaperSystem, PressureGradientDrive
    n = 64
    s = LBMSolver(nx=n, ny=n, nz=n); s.init_fields()
    fp = FilterPaperSystem(s); fp.initialize_filter_geometry()
    b = fp.get_coffee_bed_boundary()
    assert b["center_x"] == n * 0.5 and b["bottom_z"] == 5.0 and b["top_radius_lu"] > b["bottom_radius_lu"] > 0
    assert abs(b["get_radius_at_height"](b["top_z"]) - b["top_radius_lu"]) < 1e-4
    ps = CoffeeParticleSystem(3000, solver=s)
    created = ps.initialize_coffee_bed_confined(fp, seed=3)
    assert 500 < created <= 2000
    st = ps.get_particle_statistics()
    assert st["count"] == created and 0.5 * 3.25e-4 <= st["min_radius"] <= st["max_radius"] <= 1.5 * 3.25e-4
    pos0 = st["positions"]
    # every grain starts inside the cone, above the filter surface
    r = np.hypot(pos0[:, 0] - b["center_x"], pos0[:, 1] - b["center_y"])
    assert (pos0[:, 2] >= b["bottom_z"] + 2.0).all() and (r <= [b["get_radius_at_height"](z) for z in pos0[:, 2]]).all()
    cfg = s.config
    for _ in range(5):                                                     # main.py:667-679
        s.step()
        ps.update_particle_physics(cfg.DT * cfg.SCALE_TIME * 0.1, b["center_x"], b["center_y"], b["bottom_z"], b["bottom_radius_lu"],
                                   b["top_radius_lu"])
    assert int((ps.active == 1).sum()) == created and ps.coordinate_errors == 0
    drive = PressureGradientDrive(s); drive.activate_force_drive(True)
    assert drive.get_status()["force_drive"] and not drive.get_status()["density_drive"]
    for _ in range(3):
        s.clear_body_force(); drive.apply(); s.step_with_two_way_coupling(ps, 1.0, 0.8); fp.step(ps)
    drive.compute_pressure_gradient()
    assert float(drive._pressure_force.abs().max()) <= 0.12 * 1.0001
    stats = drive.get_statistics()
    assert np.isfinite(list(stats.values())).all() and s.check_stability() and drive.check_enhanced_stability()
    ux, uy, uz = s.get_velocity_components()
    assert ux.shape == (n, n, n) and s.has_soa_velocity_layout() and s.get_solver_type() == "b200"
    drive.activate_density_drive(True); drive.apply()                       # method A: nudges rho (cosmetic)
    assert not drive.force_drive_active and np.isfinite(s.rho.to_numpy()).all()


def test_body_force_accumulation_fluid_only():
    """tests/test_lbm_body_force.py:77-100,218-239 of the reference: body_force += semantics, fluid cells only."""
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    from pour_over_coffee_lbm_b200.physics import CoffeeParticleSystem, FilterPaperSystem
    n = 32
    s = LBMSolver(nx=n, ny=n, nz=n)
    FilterPaperSystem(s).initialize_filter_geometry()
    ps = CoffeeParticleSystem(64, solver=s)
    ps.reaction_force_tensor.fill_(1e-3)
    s.clear_body_force()
    s.add_particle_reaction_forces(ps); s.add_particle_reaction_forces(ps)
    bf = s.body_force.to_numpy(); solid = s.solid.to_numpy()
    assert np.allclose(bf[solid == 0], 2e-3) and np.all(bf[solid != 0] == 0)


def test_sparse_clear_of_the_reaction_field_equals_the_dense_clear():
    """lbm_particles_couple_sparse (the previous call's deposits cleared cell by cell from ps.cell) against lbm_particles_couple
    (memset of the whole field) over several calls with moving particles, some of them switched off in between: same reaction
    field (the scatter atomics are unordered: 1e-6 of the field scale), identical per-particle outputs."""
    import torch
    from pour_over_coffee_lbm_b200.engine import ParticleState, particles_couple
    n, P = 48, 5000
    rng = np.random.default_rng(3)
    eng = _engine(n, n, n, compat="physical", periodic=(False, False, False), walls=True, force=True)
    eng.u.copy_(_torch((0.02 * rng.standard_normal((3, n, n, n))).astype(np.float32)))

    def particles():
        ps = ParticleState(P, eng.device)
        r = np.random.default_rng(9)
        ps.pos.copy_(_torch(r.uniform(2.0, n - 3.0, (3, P)).astype(np.float32)))
        ps.vel.copy_(_torch((1e-3 * r.standard_normal((3, P))).astype(np.float32)))
        ps.radius.fill_(3.25e-4); ps.mass.fill_(float(4.0 / 3.0 * 3.14159 * 3.25e-4 ** 3 * 1200.0)); ps.active.fill_(1)
        return ps
    a, b = particles(), particles()
    dense = torch.zeros_like(eng.u); sparse = torch.zeros_like(eng.u)
    for it in range(4):
        particles_couple(eng, a, dense, relax=0.8)
        particles_couple(eng, b, sparse, relax=0.8, sparse_clear=True)
        scale = float(dense.abs().max())
        assert scale > 0 and float((dense - sparse).abs().max()) <= 1e-6 * scale
        assert int((sparse != 0).sum()) <= 24 * P                           # nothing of the earlier positions is left behind
        for x, y in ((a.cell, b.cell), (a.drag, b.drag), (a.u_fluid, b.u_fluid)):
            assert torch.equal(x, y)
        step = _torch((0.9 * rng.standard_normal((3, P))).astype(np.float32))
        for ps in (a, b):
            ps.pos.add_(step).clamp_(1.0, n - 2.5)
            ps.active[it::7] = 0


def test_particle_cell_indices_bit_exact_at_one_million_particles():
    """BASELINE: "particle cell indices must match bit-exactly" at the particle count of configs[3]: 10^6 particles in the V60 bed
    of a 512-wide box (positions only -- no 512^3 populations needed: u is a thin synthetic slab), kernel against the oracle's
    f32 clamp + truncation (coffee_particles.py:1054-1056), including positions outside the box and exactly on cell faces."""
    import torch
    from pour_over_coffee_lbm_b200.engine import ParticleState, particles_couple
    nx, nz, P = 512, 16, 1_000_000
    rng = np.random.default_rng(42)
    eng = _engine(nx, nx, nz, compat="physical", periodic=(False, False, False), walls=True, force=True)
    pos = np.stack([rng.uniform(-3.0, nx + 3.0, P), rng.uniform(-3.0, nx + 3.0, P), rng.uniform(-2.0, nz + 2.0, P)]).astype(np.float32)
    pos[:, ::1000] = np.round(pos[:, ::1000])                                # exactly on cell faces
    ps = ParticleState(P, eng.device)
    ps.pos.copy_(_torch(pos)); ps.radius.fill_(3.25e-4); ps.mass.fill_(1e-7); ps.active.fill_(1)
    particles_couple(eng, ps, torch.zeros_like(eng.u), relax=0.8)
    cfg = R.RefConfig(NX=nx, NY=nx, NZ=nz)
    i, j, k, _ = R.particle_cell_and_weights(cfg, pos.T.copy())
    assert np.array_equal(ps.cell.cpu().numpy(), np.stack([i, j, k]))
