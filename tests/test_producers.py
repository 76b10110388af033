"""Per-step producers next to the LBM step: MultiphaseFlow3D's surface-tension chain and phase-field
step, PrecisePouringSystem's nozzle force and gradual phase change.

tests/golden/reference_run_multiphase.npz was recorded by running the UNMODIFIED reference modules
(src/core/multiphase_3d.py, src/physics/precise_pouring.py) under the pure-Python Taichi stand-in
(tests/golden/make_reference_goldens.py).  CPU: oracle/producers_ref.py reproduces it bit for bit.  GPU: the CUDA
kernels, driven through the facade classes and the C ABI, reproduce it bit for bit -- except the nozzle's Gaussian, which
goes through expf on the device (<= 2 ulp, CUDA math library) and NumPy's exp in the recording.
"""
import os

import numpy as np
import pytest

import helpers as H
from oracle import producers_ref as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run_multiphase.npz")
GOLD_FP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run_filter_particles.npz")


@pytest.fixture(scope="module")
def z():
    return np.load(GOLD)


def _consts(z):
    return tuple(float(z[k]) for k in ("sigma", "mobility", "dt", "rho_water", "rho_air"))


# ---- CPU: oracle == recorded reference run ------------------------------------------------------------------------------
def test_oracle_surface_tension_chain_reproduces_the_reference_run(z):
    n = int(z["n"]); sig = _consts(z)[0]
    m = P.MultiphaseState(n); m.phi = z["phi"].copy()
    lap = P.compute_chemical_potential(m, sig, float(z["interface_width"]))
    assert np.array_equal(m.mu, z["mu"]) and np.array_equal(lap, z["laplacian_phi"])
    bf = z["body_force"].copy()
    P.accumulate_surface_tension_pre_collision(m, z["rho"], z["solid"], bf, sig)
    for name, got in (("grad_phi", m.grad_phi), ("grad_mu", m.grad_mu), ("normal", m.normal), ("curvature", m.curvature),
                      ("surface_force", m.surface_force), ("body_force", bf)):
        assert np.array_equal(got, z["st_" + name]), name
    # the scenario exercises every branch: flat patches without a normal, saturated cells, the rho guard
    c = (slice(1, -1),) * 3
    assert (np.linalg.norm(z["st_normal"][c], axis=-1) == 0).sum() > 0
    assert (np.abs(z["phi"][c]) >= 0.9).sum() > 100 and (np.abs(z["st_surface_force"]).sum(-1) > 0).sum() > 100
    assert (z["st_body_force"] != z["body_force"]).any(-1).sum() > 100


def test_oracle_phase_field_step_reproduces_the_reference_run(z):
    n = int(z["n"]); sig, mob, dt, rw, ra = _consts(z)
    m = P.MultiphaseState(n); m.phi = z["phi"].copy(); m.phi_new = z["phi_new_in"].copy(); m.mu = z["mu"].copy()
    rho = z["rho"].copy(); bf = z["st_body_force"].copy(); phase = np.zeros_like(rho)
    P.multiphase_step(m, z["u"], rho, phase, z["solid"], bf, sig, mob, dt, rw, ra, 20, True)
    for name, got in (("phi", m.phi), ("phi_new", m.phi_new), ("rho", rho), ("phase", phase), ("body_force", bf)):
        assert np.array_equal(got, z["s1_" + name]), name
    # the outer layer of phi comes from phi_new's (never written) outer layer: copy_phase_field copies everything
    assert np.array_equal(m.phi[0], z["phi_new_in"][0]) and not np.array_equal(z["phi"][0], z["phi_new_in"][0])
    P.multiphase_step(m, z["u"], rho, phase, z["solid"], bf, sig, mob, dt, rw, ra, 21, False)
    for name, got in (("phi", m.phi), ("rho", rho), ("phase", phase), ("body_force", bf), ("surface_force", m.surface_force),
                      ("curvature", m.curvature)):
        assert np.array_equal(got, z["s2_" + name]), name
    assert (z["s2_body_force"] != z["s1_body_force"]).any()          # step_count > 10 and not precollision_applied: force re-applied


def test_oracle_pouring_reproduces_the_reference_run(z):
    n = int(z["n"])
    bf = z["s2_body_force"].copy(); phi = z["s2_phi"].copy(); solid = z["solid"]
    p = P.PourState(n, float(z["pour_diameter"]), int(z["pour_height"]), float(z["pour_velocity"]))
    P.apply_pouring_force(p, bf, solid, 1.0)                         # not started: nothing happens
    assert np.array_equal(bf, z["s2_body_force"])
    p.start_pouring(pattern="center", flow_rate=0.3)
    P.apply_pouring_force(p, bf, solid, 0.1); P.apply_gradual_phase_change(p, phi, solid, 0.1)
    assert np.array_equal(bf, z["p1_body_force"]) and np.array_equal(phi, z["p1_phi"])
    assert (z["p1_body_force"] != z["s2_body_force"]).any(-1).sum() > 50
    p.start_pouring(pattern="spiral", flow_rate=1.0)
    for dt in z["p2_dts"]:
        P.apply_pouring_force(p, bf, solid, float(dt)); P.apply_gradual_phase_change(p, phi, solid, float(dt))
    assert np.array_equal(bf, z["p2_body_force"]) and np.array_equal(phi, z["p2_phi"])
    assert float(p.pour_time) == float(z["p2_pour_time"])
    assert np.abs(z["p2_body_force"] - z["p1_body_force"]).max() == pytest.approx(10.0, rel=0.02)    # the 10 lu/ts^2 cap was hit
    p.active = 0
    P.apply_pouring_force(p, bf, solid, 1.0)
    assert np.array_equal(bf, z["p3_body_force"])


def test_facade_constants_match_the_reference_run(z):
    """SURFACE_TENSION_LU, RHO_AIR, INLET_VELOCITY, the nozzle diameter and height the reference derived at 16^3."""
    from pour_over_coffee_lbm_b200.config import LBMConfig
    n = int(z["n"])
    c = LBMConfig(NX=n, NY=n, NZ=n)
    assert c.SURFACE_TENSION_LU == float(z["sigma"]) and c.RHO_AIR == float(z["rho_air"]) and c.RHO_WATER == float(z["rho_water"])
    assert c.INLET_VELOCITY == float(z["pour_velocity"]) and c.DT == float(z["dt"])
    assert 0.5 / c.GRID_SIZE_CM == float(z["cfg_pour_diameter_grid"])
    assert max(8, min(int(int(5.0 + int(c.CUP_HEIGHT / c.SCALE_LENGTH)) + 2), c.NZ - 6)) == int(z["cfg_pour_height"])


def test_oracle_filter_particle_interception_reproduces_the_reference_run():
    """FilterPaperSystem.block_particles_at_filter + update_dynamic_resistance, recorded with ti.random() pinned to 0.5
    (Taichi's stream is unseeded), i.e. without the horizontal kick: noise = 0 here."""
    z = np.load(GOLD_FP)
    vel = z["p_vel"].copy(); acc = z["accumulated_in"].copy(); blk = z["blockage_in"].copy()
    for t in range(2):
        P.block_particles_at_filter(z["filter_zone"], z["p_pos"], vel, z["p_active"], acc, float(z["scale_length"]), noise=0.0)
        assert np.array_equal(vel, z[f"b{t}_vel"]) and np.array_equal(acc, z[f"b{t}_accumulated"])
    assert (z["b0_vel"][:, 2] != z["p_vel"][:, 2]).sum() > 20 and np.array_equal(z["b0_vel"][:, :2], z["p_vel"][:, :2])
    assert (z["b0_accumulated"] != z["accumulated_in"]).sum() > 10
    for t in range(2):
        P.update_dynamic_resistance(z["filter_zone"], blk, acc)
        assert np.array_equal(blk, z[f"r{t}_blockage"]) and np.array_equal(acc, z[f"r{t}_accumulated"])
    # CoffeeParticleSystem.apply_fluid_forces: drag + buoyancy + gravity with every guard of the reference
    vel = z["ff_vel"].copy(); act = z["ff_active"].copy(); force = z["ff_force_in"].copy()
    err = P.apply_fluid_forces(z["ff_u"], z["ff_pos"], vel, z["ff_radius"], z["ff_mass"], act, force, float(z["water_density"]),
                               float(z["water_viscosity"]), float(z["particle_gravity"]))
    assert np.array_equal(force, z["ff_force"]) and np.array_equal(vel, z["ff_vel_out"]) and np.array_equal(act, z["ff_active_out"])
    assert err == int(z["ff_errors"]) > 10 and (z["ff_force"] != z["ff_force_in"]).any(-1).sum() > 50
    # the kick of the product path: bounded by noise / 2, a pure function of (seed, particle)
    v1 = z["p_vel"].copy(); v2 = z["p_vel"].copy(); v3 = z["p_vel"].copy()
    a = z["accumulated_in"].copy()
    P.block_particles_at_filter(z["filter_zone"], z["p_pos"], v1, z["p_active"], a.copy(), float(z["scale_length"]), 0.01, seed=7)
    P.block_particles_at_filter(z["filter_zone"], z["p_pos"], v2, z["p_active"], a.copy(), float(z["scale_length"]), 0.01, seed=7)
    P.block_particles_at_filter(z["filter_zone"], z["p_pos"], v3, z["p_active"], a.copy(), float(z["scale_length"]), 0.01, seed=8)
    hit = z["b0_vel"][:, 2] != z["p_vel"][:, 2]
    assert np.array_equal(v1, v2) and not np.array_equal(v1, v3) and np.array_equal(v1[:, 2], z["b0_vel"][:, 2])
    assert np.abs(v1[:, :2] - z["p_vel"][:, :2]).max() <= 0.005 + 1e-9 and np.array_equal(v1[~hit], z["p_vel"][~hit])
    assert 0.2 < np.mean([float(P.uniform01(3, p, 0)) for p in range(2000)]) - 0.0 < 0.8


def test_pouring_facade_host_state_follows_the_oracle(z):
    """The host half of PrecisePouringSystem (no device needed): pour_time accumulation, centre / spiral nozzle position,
    the struct handed to lbm_pouring_force, flow-rate limits, diagnostics."""
    from pour_over_coffee_lbm_b200.config import LBMConfig
    from pour_over_coffee_lbm_b200.physics import PrecisePouringSystem
    n = int(z["n"])
    pp = PrecisePouringSystem(config=LBMConfig(NX=n, NY=n, NZ=n))
    pp.POUR_DIAMETER_GRID = float(z["pour_diameter"]); pp.POUR_HEIGHT = int(z["pour_height"])
    ref = P.PourState(n, float(z["pour_diameter"]), int(z["pour_height"]), float(z["pour_velocity"]))
    assert pp.get_pouring_info()["active"] is False and pp.get_current_flow_rate() == 0.0
    pp.start_pouring(pattern="spiral", flow_rate=1.0); ref.start_pouring(pattern="spiral", flow_rate=1.0)
    for dt in (0.5, 1e-3, 0.25, 2.0):
        pp.pour_time[None] = pp.pour_time[None] + np.float32(dt)          # what apply_pouring_force does before the launch
        ref.pour_time = np.float32(ref.pour_time + np.float32(dt))
        x, y = pp._get_current_pour_position(); rx, ry = ref.position()
        assert x == rx and y == ry and x.dtype == np.float32
        st = pp._pour_struct(dt)
        assert st.pour_x == float(rx) and st.pour_y == float(ry) and st.pour_z == int(z["pour_height"])
        assert st.radius == float(np.float32(float(z["pour_diameter"]) / 2.0)) and st.dt == float(np.float32(dt))
    assert float(pp.pour_time[None]) == float(ref.pour_time)
    info = pp.get_pouring_info()
    assert info["active"] and info["pattern"] == 1 and info["position"] == (float(x), float(y))
    pp.adjust_flow_rate(7.0); assert float(pp.pour_flow_rate[None]) == 3.0
    pp.adjust_flow_rate(0.0); assert float(pp.pour_flow_rate[None]) == np.float32(0.1)
    pp.move_pour_center(-3, 100); assert (float(pp.pour_center_x[None]), float(pp.pour_center_y[None])) == (5.0, n - 5.0)
    chk = pp._check_pouring_conditions()
    assert chk["affected_cells"] > 0 and 0 < chk["effectiveness"] <= 1 and chk["z_range"] == [pp.POUR_HEIGHT - 4.0, pp.POUR_HEIGHT]
    assert pp.get_current_flow_rate_ml_s() == pytest.approx(0.4) and pp.get_pouring_diagnostics()["configuration"]["height"] == pp.POUR_HEIGHT
    with pytest.raises(ValueError):
        pp.apply_pouring_force(None, None, 0.1)                              # not bound to a solver: refuses, never a CPU path
    pp.stop_pouring()
    pp.apply_pouring_force(None, None, 0.1)                                  # inactive: returns before touching anything


# ---- GPU: CUDA kernels through the facades == recorded reference run ---------------------------------------------------
def _torch(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _solver(z, **kw):
    from pour_over_coffee_lbm_b200.physics import FilterPaperSystem, MultiphaseFlow3D
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    n = int(z["n"])
    s = LBMSolver(nx=n, ny=n, nz=n, compat="reference", strict=True, **kw)
    s.init_fields()
    FilterPaperSystem(s).initialize_filter_geometry()
    assert np.array_equal(H.from_dev_scalar(s.engine.solid), z["solid"])
    mp = MultiphaseFlow3D(s)
    assert mp.SURFACE_TENSION_COEFF == float(z["sigma"]) and mp.MOBILITY == float(z["mobility"])
    mp.phi.from_numpy(z["phi"]); mp.phi_new.from_numpy(z["phi_new_in"])
    s.rho.from_numpy(z["rho"]); s.u.from_numpy(z["u"]); s.body_force.from_numpy(z["body_force"])
    return s, mp


@pytest.mark.gpu
def test_gpu_multiphase_chain_reproduces_the_reference_run(z):
    s, mp = _solver(z)
    mp.compute_chemical_potential()
    assert np.array_equal(mp.mu.to_numpy(), z["mu"]) and np.array_equal(mp.laplacian_phi.to_numpy(), z["laplacian_phi"])
    launches = s.engine.launch_count()
    mp.accumulate_surface_tension_pre_collision()
    assert s.engine.launch_count() - launches == 2                  # 4 Taichi kernels of the reference in 2 launches
    for name, field in (("grad_phi", mp.grad_phi), ("grad_mu", mp.grad_mu), ("normal", mp.normal), ("curvature", mp.curvature),
                        ("surface_force", mp.surface_force), ("body_force", s.body_force)):
        assert np.array_equal(field.to_numpy(), z["st_" + name]), name
    mp.step(20, precollision_applied=True)
    for name, field in (("phi", mp.phi), ("phi_new", mp.phi_new), ("rho", s.rho), ("phase", s.phase), ("body_force", s.body_force)):
        assert np.array_equal(field.to_numpy(), z["s1_" + name]), name
    mp.step(21, precollision_applied=False)
    for name, field in (("phi", mp.phi), ("rho", s.rho), ("phase", s.phase), ("body_force", s.body_force),
                        ("surface_force", mp.surface_force), ("curvature", mp.curvature)):
        assert np.array_equal(field.to_numpy(), z["s2_" + name]), name
    # the stand-alone entry points
    s.body_force.from_numpy(z["s1_body_force"]); s.rho.from_numpy(z["s1_rho"])
    mp.apply_surface_tension()
    assert np.array_equal(s.body_force.to_numpy(), z["s2_body_force"])
    mp.update_density_from_phase()
    assert np.array_equal(s.rho.to_numpy(), z["s2_rho"]) and np.array_equal(s.phase.to_numpy(), z["s2_phase"])
    st = mp.get_interface_statistics()
    assert st["max_curvature"] == pytest.approx(float(np.abs(z["s2_curvature"]).max()), rel=1e-6)


@pytest.mark.gpu
def test_gpu_pouring_reproduces_the_reference_run(z):
    from pour_over_coffee_lbm_b200.physics import PrecisePouringSystem
    s, mp = _solver(z)
    pp = PrecisePouringSystem(s)
    assert pp.POUR_DIAMETER_GRID == float(z["cfg_pour_diameter_grid"]) and pp.POUR_HEIGHT == int(z["cfg_pour_height"])
    assert pp.POUR_VELOCITY == float(z["pour_velocity"])
    pp.POUR_DIAMETER_GRID = float(z["pour_diameter"]); pp.POUR_HEIGHT = int(z["pour_height"])
    s.body_force.from_numpy(z["s2_body_force"]); mp.phi.from_numpy(z["s2_phi"])
    pp.apply_pouring_force(s.body_force, s.solid, 1.0)              # not started
    assert np.array_equal(s.body_force.to_numpy(), z["s2_body_force"])

    def close(got, want, base, what):
        # untouched cells are bit-identical; nozzle cells carry expf's <= 2 ulp on the increment
        touched = want != base
        assert np.array_equal(got[~touched], want[~touched]), what
        inc = np.abs(want - base)[touched]
        assert np.all(np.abs(got - want)[touched] <= 8 * np.finfo(np.float32).eps * np.maximum(inc, np.abs(want[touched]))), what

    pp.start_pouring(pattern="center", flow_rate=0.3)
    pp.apply_pouring_force(s.body_force, s.solid, 0.1); pp.apply_gradual_phase_change(mp.phi, s.solid, 0.1)
    close(s.body_force.to_numpy(), z["p1_body_force"], z["s2_body_force"], "centre force")
    close(mp.phi.to_numpy(), z["p1_phi"], z["s2_phi"], "centre phase change")
    s.body_force.from_numpy(z["p1_body_force"]); mp.phi.from_numpy(z["p1_phi"])
    pp.start_pouring(pattern="spiral", flow_rate=1.0)
    for dt in z["p2_dts"]:
        pp.apply_pouring_force(s.body_force, s.solid, float(dt)); pp.apply_gradual_phase_change(mp.phi, s.solid, float(dt))
    assert float(pp.pour_time[None]) == float(z["p2_pour_time"])
    bf = s.body_force.to_numpy(); phi = mp.phi.to_numpy()
    assert np.allclose(bf, z["p2_body_force"], rtol=1e-6, atol=1e-7) and np.array_equal(bf[z["p2_body_force"] == z["p1_body_force"]],
                                                                                         z["p1_body_force"][z["p2_body_force"] == z["p1_body_force"]])
    assert np.allclose(phi, z["p2_phi"], rtol=1e-6, atol=1e-7)
    info = pp.get_pouring_info()
    assert info["active"] and info["pattern"] == 1
    pp.stop_pouring()
    before = s.body_force.to_numpy()
    pp.apply_pouring_force(s.body_force, s.solid, 1.0)
    assert np.array_equal(s.body_force.to_numpy(), before)


@pytest.mark.gpu
def test_gpu_main_py_step_order_with_producers_runs(z):
    """main.py:770-839 on the device: clear -> pouring force + phase change -> pressure drive -> surface tension ->
    step -> multiphase step, a few iterations on the V60 box; fields stay finite and the phase field stays in [-1, 1]."""
    import torch
    from pour_over_coffee_lbm_b200.physics import PrecisePouringSystem, PressureGradientDrive
    s, mp = _solver(z, gravity_lu=1e-5)          # the default GRAVITY_LU of a 16^3 box (618 lu/ts^2) saturates everything
    s.init_fields()
    mp.standardize_initial_state(force_dry_state=True)
    pp = PrecisePouringSystem(s); pp.POUR_DIAMETER_GRID = 5.0; pp.POUR_HEIGHT = 10
    pd = PressureGradientDrive(s); pd.activate_force_drive(True)
    pp.start_pouring(pattern="center"); pp.adjust_flow_rate(0.3)
    for it in range(1, 14):
        s.clear_body_force()
        pp.apply_pouring_force(s.body_force, s.solid, 1.0)
        pp.apply_gradual_phase_change(mp.phi, s.solid, 1.0)
        pd.apply(it)
        if it > 10:
            mp.accumulate_surface_tension_pre_collision()
        s.step()
        mp.step(it, precollision_applied=True)
    phi = mp._phi
    assert torch.isfinite(phi).all() and float(phi.abs().max()) <= 1.0 + 1e-3
    assert float((phi > -1.0).sum()) > 0                            # water arrived under the nozzle
    assert torch.isfinite(s.engine.rho).all() and torch.isfinite(s.engine.u).all()
    assert float(s.engine.body_force[2].min()) < 0.0                # the nozzle pushed down this step


@pytest.mark.gpu
def test_gpu_filter_particle_interception_reproduces_the_reference_run():
    """lbm_particles_block_at_filter / lbm_filter_dynamic_resistance through FilterPaperSystem: bounce, accumulation (atomics of
    one constant: order-free) bit-exact with noise = 0; the blockage update within expf's 2 ulp; the kick equals the oracle's
    counter-based draw bit for bit."""
    import torch
    from pour_over_coffee_lbm_b200.physics import CoffeeParticleSystem, FilterPaperSystem
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    z = np.load(GOLD_FP)
    n = int(z["n"]); npart = z["p_pos"].shape[0]
    s = LBMSolver(nx=n, ny=n, nz=n, compat="reference", strict=True); s.init_fields()
    fp = FilterPaperSystem(s); fp.initialize_filter_geometry()
    assert np.array_equal(H.from_dev_scalar(s.engine.filter_zone), z["filter_zone"])
    assert float(np.float32(s.config.SCALE_LENGTH)) == float(np.float32(z["scale_length"]))
    ps = CoffeeParticleSystem(npart, solver=s)

    def load():
        ps.set_particles(z["p_pos"], z["p_vel"], z["p_radius"])
        ps.state.active.copy_(torch.from_numpy(z["p_active"]).cuda())
        fp._ensure_accumulated()
        fp.accumulated_particles.from_numpy(z["accumulated_in"]); fp.filter_blockage.from_numpy(z["blockage_in"])
    load()
    for t in range(2):
        fp.block_particles_at_filter(particle_system=ps, noise=0.0)
        assert np.array_equal(ps.velocity.cpu().numpy(), z[f"b{t}_vel"])
        assert np.array_equal(fp.accumulated_particles.to_numpy(), z[f"b{t}_accumulated"])
    for t in range(2):
        fp.update_dynamic_resistance()
        assert np.array_equal(fp.accumulated_particles.to_numpy(), z[f"r{t}_accumulated"])
        assert np.allclose(fp.filter_blockage.to_numpy(), z[f"r{t}_blockage"], rtol=1e-6, atol=1e-8)
    load()
    fp.block_particles_at_filter(particle_system=ps, noise=0.01, seed=7)
    want = z["p_vel"].copy(); acc = z["accumulated_in"].copy()
    P.block_particles_at_filter(z["filter_zone"], z["p_pos"], want, z["p_active"], acc, float(z["scale_length"]), 0.01, seed=7)
    assert np.array_equal(ps.velocity.cpu().numpy(), want) and np.array_equal(fp.accumulated_particles.to_numpy(), acc)


@pytest.mark.gpu
def test_gpu_apply_fluid_forces_reproduces_the_reference_run():
    """lbm_particles_fluid_forces through CoffeeParticleSystem.apply_fluid_forces: force, reset velocities, deactivated
    particles and the error counter equal the recorded run of the reference's kernel bit for bit."""
    import torch
    from pour_over_coffee_lbm_b200.physics import CoffeeParticleSystem
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    z = np.load(GOLD_FP)
    n = int(z["n"]); npart = z["ff_pos"].shape[0]
    s = LBMSolver(nx=n, ny=n, nz=n, compat="reference", strict=True); s.init_fields()
    s.u.from_numpy(z["ff_u"])
    ps = CoffeeParticleSystem(npart, solver=s)
    assert ps.water_density == float(z["water_density"]) and ps.water_viscosity == float(z["water_viscosity"]) and ps.gravity == float(z["particle_gravity"])
    ps.set_particles(z["ff_pos"], z["ff_vel"], z["ff_radius"], z["ff_mass"])
    ps.state.active.copy_(torch.from_numpy(z["ff_active"]).cuda())
    ps.force_tensor = torch.from_numpy(np.ascontiguousarray(z["ff_force_in"].T)).cuda()
    ps.error_counters = torch.zeros(2, dtype=torch.int32, device="cuda")
    ps.apply_fluid_forces(s.u, s.u, s.u, s.rho, s.rho, 0.01)
    assert np.array_equal(ps.force.cpu().numpy(), z["ff_force"])
    assert np.array_equal(ps.velocity.cpu().numpy(), z["ff_vel_out"]) and np.array_equal(ps.active.cpu().numpy(), z["ff_active_out"])
    assert ps.coordinate_errors == int(z["ff_errors"])


@pytest.mark.gpu
@pytest.mark.parametrize("nx", [24, 18], ids=["vec4", "ragged_vec1"])
def test_gpu_multiphase_kernels_on_a_non_cubic_box_match_the_oracle(nx):
    """nx = 24 runs the 4-cells-per-thread kernels (edge lanes of every row live), nx = 18 the one-cell kernels; random fields
    everywhere, the never-written outer layers included; bit-exact against oracle/producers_ref.py."""
    import torch
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine
    ny, nz = 10, 7
    rng = np.random.default_rng(3)
    sh = (nx, ny, nz)
    phi = np.clip(rng.normal(0.0, 0.8, sh), -1, 1).astype(np.float32); phi_new0 = rng.normal(0, 0.1, sh).astype(np.float32)
    mu0 = rng.normal(0, 0.05, sh).astype(np.float32); u = rng.normal(0, 0.05, sh + (3,)).astype(np.float32)
    rho = (1 + 0.1 * rng.standard_normal(sh)).astype(np.float32); rho[3, 4, 2] = 0.0
    bf0 = (1e-3 * rng.standard_normal(sh + (3,))).astype(np.float32); sf0 = (1e-3 * rng.standard_normal(sh + (3,))).astype(np.float32)
    solid = (rng.random(sh) < 0.4).astype(np.uint8)
    eng = D3Q19Engine(nx, ny, nz, compat="reference", periodic=(False, False, False), walls=True, force=True, phase=True)
    eng.solid.copy_(_torch(H.to_dev_scalar(solid))); eng.pack_flags()
    eng.rho.copy_(_torch(H.to_dev_scalar(rho))); eng.u.copy_(_torch(H.to_dev_vec(u))); eng.body_force.copy_(_torch(H.to_dev_vec(bf0)))
    d_phi, d_new, d_mu = _torch(H.to_dev_scalar(phi)), _torch(H.to_dev_scalar(phi_new0)), _torch(H.to_dev_scalar(mu0))
    d_sf = _torch(H.to_dev_vec(sf0)); d_curv = torch.zeros_like(d_phi)
    d_g, d_gm, d_n = torch.zeros_like(d_sf), torch.zeros_like(d_sf), torch.zeros_like(d_sf)
    eng.surface_tension(d_phi, d_mu, d_g, d_gm, d_n, d_curv, d_sf, 0.05, apply=True)
    eng.phase_field_step(d_phi, d_new, d_mu, 0.001, 1.0, 1.0, 0.00125)
    m = P.MultiphaseState(sh); m.phi = phi.copy(); m.phi_new = phi_new0.copy(); m.mu = mu0.copy(); m.surface_force = sf0.copy()
    bf = bf0.copy(); r = rho.copy(); ph = np.zeros_like(r)
    P.accumulate_surface_tension_pre_collision(m, r, solid, bf, 0.05)
    P.update_phase_field_cahn_hilliard(m, u, 0.001, 1.0); P.apply_phase_separation(m, 1.0); m.phi[...] = m.phi_new
    P.update_density_from_phase(m, r, ph, 1.0, 0.00125)
    assert np.array_equal(H.from_dev_scalar(d_phi), m.phi) and np.array_equal(H.from_dev_scalar(d_new), m.phi_new)
    assert np.array_equal(H.from_dev_scalar(eng.rho), r) and np.array_equal(H.from_dev_scalar(eng.phase), ph)
    assert np.array_equal(H.from_dev_vec(eng.body_force), bf) and np.array_equal(H.from_dev_vec(d_sf), m.surface_force)
    assert np.array_equal(H.from_dev_scalar(d_curv), m.curvature) and np.array_equal(H.from_dev_vec(d_n), m.normal)
    assert np.array_equal(H.from_dev_vec(d_g), m.grad_phi) and np.array_equal(H.from_dev_vec(d_gm), m.grad_mu)


@pytest.mark.gpu
def test_gpu_lazy_fields_same_body_force_in_one_launch(z):
    """MultiphaseFlow3D(lazy_fields=True): lbm_surface_tension_body_force (one launch over the interface band, no intermediate
    fields) gives the recorded body_force bit for bit; the diagnostic fields appear, exact, when they are read."""
    from pour_over_coffee_lbm_b200.physics import MultiphaseFlow3D
    s, _ = _solver(z)
    mp = MultiphaseFlow3D(s, lazy_fields=True)
    mp.phi.from_numpy(z["phi"]); mp.phi_new.from_numpy(z["phi_new_in"])
    mp.compute_chemical_potential()
    launches = s.engine.launch_count()
    mp.accumulate_surface_tension_pre_collision()
    assert s.engine.launch_count() - launches == 1
    assert np.array_equal(s.body_force.to_numpy(), z["st_body_force"])
    assert np.array_equal(mp.curvature.to_numpy(), z["st_curvature"]) and np.array_equal(mp.surface_force.to_numpy(), z["st_surface_force"])
    assert np.array_equal(mp.normal.to_numpy(), z["st_normal"]) and np.array_equal(s.body_force.to_numpy(), z["st_body_force"])
    launches = s.engine.launch_count()
    mp.step(20, precollision_applied=True)
    assert s.engine.launch_count() - launches == 2                  # only the phase-field update: the field kernels were deferred
    for name, field in (("phi", mp.phi), ("phi_new", mp.phi_new), ("rho", s.rho), ("phase", s.phase), ("body_force", s.body_force)):
        assert np.array_equal(field.to_numpy(), z["s1_" + name]), name
    mp.step(21, precollision_applied=False)
    for name, field in (("phi", mp.phi), ("rho", s.rho), ("phase", s.phase), ("body_force", s.body_force)):
        assert np.array_equal(field.to_numpy(), z["s2_" + name]), name


@pytest.mark.gpu
def test_gpu_one_launch_surface_tension_equals_the_chain_at_scale():
    """Size-independent property at 128^3 with the V60 mask: body_force from lbm_surface_tension_body_force (one launch, interface
    band only) is bit-identical to the four-kernel chain's, for a wavy noisy interface that crosses the cone."""
    import torch
    from pour_over_coffee_lbm_b200.physics import FilterPaperSystem
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    n = 128
    s = LBMSolver(nx=n, ny=n, nz=n, compat="reference", strict=True); s.init_fields()
    FilterPaperSystem(s).initialize_filter_geometry()
    e = s.engine
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    zc = torch.arange(n, device="cuda", dtype=torch.float32)[:, None, None]
    xc = torch.arange(n, device="cuda", dtype=torch.float32)[None, None, :]
    phi = torch.tanh((0.5 * n + 4.0 * torch.sin(0.2 * xc) - zc) / 2.0).expand(n, n, n).contiguous()
    phi = (phi + 0.02 * torch.randn((n, n, n), device="cuda", generator=g)).clamp_(-1.0, 1.0)
    e.rho.copy_(1.0 + 0.05 * torch.randn((n, n, n), device="cuda", generator=g))
    bf0 = 1e-4 * torch.randn((3, n, n, n), device="cuda", generator=g)
    sc = lambda: torch.zeros_like(e.rho)
    vc = lambda: torch.zeros_like(e.body_force)
    gphi, nrm, sf, curv = vc(), vc(), vc(), sc()
    e.body_force.copy_(bf0)
    e.surface_tension(phi, None, gphi, None, nrm, curv, sf, 0.05, apply=True)
    chain = e.body_force.clone()
    e.body_force.copy_(bf0)
    e.surface_tension_body_force(phi, 0.05)
    assert torch.equal(e.body_force, chain)
    changed = int((chain != bf0).any(0).sum())
    assert 2_000 < changed < 0.2 * n ** 3                           # the band is thin: most of the box is never touched
    assert float(curv.abs().max()) > 0 and torch.isfinite(chain).all()
