#!/usr/bin/env python
"""Pins the oracle to the REFERENCE'S OWN SOURCE CODE (run in the authoring container, where /root/reference exists).

The reference is Python + Taichi and Taichi cannot be installed here, so its kernels are executed by a small pure-Python
stand-in for the `taichi` package (tests/golden/taichi_shim: fields = NumPy arrays, IEEE f32 scalar arithmetic in source
order, see its docstring).  This script imports the reference's unmodified modules from /root/reference on a small grid
(config/core.py is loaded with NX = NY = NZ patched -- the only change, done in memory), drives LBMSolver.step() and the
neighbour kernels on seeded inputs, and writes the inputs and what the reference's code computed to
tests/golden/reference_run_*.npz.  tests/test_oracle_vs_reference_run.py (CPU) checks oracle/ against these fixtures;
the GPU parity tests check the CUDA kernels against the same files.  Nothing under tests/ reads /root/reference.
"""
import importlib.util
import io
import contextlib
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def load_reference(n: int):
    """Import the reference with the Taichi stand-in and an n^3 grid.  Returns the `config` module."""
    sys.path.insert(0, os.path.join(HERE, "taichi_shim"))
    sys.path.insert(1, REF)
    src = open(os.path.join(REF, "config", "core.py")).read()
    for name in ("NX", "NY", "NZ"):
        src, k = re.subn(rf"^{name} = 224\s*$", f"{name} = {n}", src, flags=re.M)
        assert k == 1, name

    core = types.ModuleType("config.core"); core.__file__ = os.path.join(REF, "config", "core.py")
    exec(compile(src, core.__file__, "exec"), core.__dict__)
    sys.modules["config.core"] = core
    with contextlib.redirect_stdout(io.StringIO()):
        import config
    assert config.NX == n and sys.modules["config.core"] is core
    return config


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def run_step_scenario(config, n, steps, seed, gravity, phase_mode):
    """LBMSolver + FilterPaperSystem wired like main.py (main.py:560-640), seeded state, `steps` calls of step()."""
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
    import helpers as H
    config.GRAVITY_LU = gravity          # an input (config/physics.py value saturates every clamp); read at kernel time
    with quiet():
        from src.core.legacy.lbm_solver import LBMSolver
        from src.physics.filter_paper import FilterPaperSystem
        s = LBMSolver()
        s.init_fields()
        fp = FilterPaperSystem(s)
        fp.initialize_filter_geometry()
        s.boundary_manager.set_filter_system(fp)
    # the same seeded inputs tests/helpers.py:reference_v60_state builds for the oracle
    if phase_mode == "air_random":      # tau_air everywhere (the stable regime of the legacy solver), gravity*phase active
        st = H.reference_v60_state(n, seed=seed, gravity=gravity, body=1e-5, phase_mode="none")
        st.phase[:] = np.random.default_rng(seed + 1000).uniform(0.0, 0.5, size=st.phase.shape).astype(np.float32)
    else:
        st = H.reference_v60_state(n, seed=seed, gravity=gravity, body=1e-5, phase_mode=phase_mode)
    inp = dict(f=st.f.copy(), phase=st.phase.copy(), body_force=st.body_force.copy())
    s.f.from_numpy(inp["f"]); s.f_new.from_numpy(inp["f"]); s.phase.from_numpy(inp["phase"]); s.body_force.from_numpy(inp["body_force"])
    geom = dict(solid=s.solid.to_numpy().astype(np.uint8), filter_zone=fp.filter_zone.to_numpy().astype(np.int32),
                les_mask=s.les_mask.to_numpy().astype(np.int32))
    with quiet():
        for _ in range(steps):
            s.step()
    out = dict(rho=s.rho.to_numpy(), u=s.u.to_numpy(), f_out=s.f.to_numpy())
    if hasattr(s, "les_model") and s.les_model is not None and hasattr(s.les_model, "nu_sgs"):
        out["nu_sgs"] = s.les_model.nu_sgs.to_numpy()
    return inp, geom, out, st


def run_open_box_scenario(config, n, steps, seed, gravity):
    """The first 30 steps of main.py: LBMSolver in its init_fields geometry (no V60 mask, no filter system) -- every face
    is an open face: populations that would enter from outside keep their stale w_q (quirk Q6), the boundary manager's
    top / bottom / outlet strategies write rho on the fluid faces (quirk Q5).  A solid obstacle is added so bounce-back
    next to open faces is covered as well."""
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
    import helpers as H
    from oracle import d3q19_ref as R
    config.GRAVITY_LU = gravity
    with quiet():
        from src.core.legacy.lbm_solver import LBMSolver
        s = LBMSolver()
        s.init_fields()
    rng = np.random.default_rng(seed)
    solid = np.zeros((n, n, n), np.uint8)
    solid[5:9, 4:8, 0:3] = 1; solid[n - 1, 3:7, 6:10] = 1; solid[6:10, 6:10, 7:11] = 1
    phase = rng.uniform(0.0, 1.0, (n, n, n)).astype(np.float32)
    bf = (1e-5 * rng.standard_normal((n, n, n, 3))).astype(np.float32)
    u0 = H.smooth_velocity(n, 0.02, seed); rho0 = H.smooth_density(n, 0.01, seed)
    f0 = np.stack([R.equilibrium_ref(rho0, u0[..., 0], u0[..., 1], u0[..., 2], q, "config") for q in range(19)]).astype(np.float32)
    # a state reachable from init_fields (f = f_new = w_q): populations that would have entered through a face were never
    # written by any step, so those slots still hold w_q in both buffers
    for q in range(19):
        for ax, e in enumerate((int(R.CX[q]), int(R.CY[q]), int(R.CZ[q]))):
            if e != 0:
                sl = [slice(None)] * 3; sl[ax] = 0 if e > 0 else -1
                f0[q][tuple(sl)] = R.W[q]
    s.solid.from_numpy(solid); s.phase.from_numpy(phase); s.body_force.from_numpy(bf); s.f.from_numpy(f0); s.f_new.from_numpy(f0)
    with quiet():
        for _ in range(steps):
            s.step()
    inp = dict(f=f0, phase=phase, body_force=bf, solid=solid, les_mask=s.les_mask.to_numpy().astype(np.int32))
    out = dict(rho=s.rho.to_numpy(), u=s.u.to_numpy(), f_out=s.f.to_numpy())
    return inp, out


def oracle_open_box(inp, n, steps, gravity):
    from oracle import d3q19_ref as R
    st = R.init_fields(R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=gravity))
    st.solid = inp["solid"].copy(); st.phase = inp["phase"].copy(); st.body_force = inp["body_force"].copy()
    st.f = inp["f"].copy(); st.f_new = inp["f"].copy(); st.les_mask = inp["les_mask"].copy()
    for _ in range(steps):
        R.step(st)
    return st


def run_neighbour_scenario(config, n, seed):
    """Body-force producers and the particle kernels of the reference on one seeded state:
    PressureGradientDrive (force and mixed mode), FilterPaperSystem.compute_forchheimer_resistance,
    CoffeeParticleSystem.compute_two_way_coupling_forces / apply_under_relaxation / update_particle_physics."""
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
    import helpers as H
    with quiet():
        from src.core.legacy.lbm_solver import LBMSolver
        from src.physics.filter_paper import FilterPaperSystem
        from src.physics.pressure_gradient_drive import PressureGradientDrive
        from src.physics.coffee_particles import CoffeeParticleSystem
        s = LBMSolver(); s.init_fields()
        fp = FilterPaperSystem(s); fp.initialize_filter_geometry()
        pg = PressureGradientDrive(s)
    rng = np.random.default_rng(seed)
    shp = (n, n, n)
    rho = (1.0 + 0.05 * rng.standard_normal(shp)).astype(np.float32)
    u = (0.02 * rng.standard_normal(shp + (3,))).astype(np.float32)
    s.rho.from_numpy(rho); s.u.from_numpy(u)
    res = dict(n=n, rho=rho, u=u, solid=s.solid.to_numpy().astype(np.uint8), filter_zone=fp.filter_zone.to_numpy().astype(np.int32))
    with quiet():
        s.body_force.fill(0.0); pg.activate_force_drive(True); pg.apply(0)
        res["bf_force_drive"] = s.body_force.to_numpy()
        s.body_force.fill(0.0); pg.activate_mixed_drive(True); pg.apply(0)
        res["bf_mixed_drive"] = s.body_force.to_numpy()
        fp.compute_forchheimer_resistance()
        res["bf_mixed_plus_forchheimer"] = s.body_force.to_numpy()
    # ---- particles ----
    P = 400
    with quiet():
        ps = CoffeeParticleSystem(P)
    pos = rng.uniform(-1.0, n + 1.0, (P, 3)).astype(np.float32)
    pos[:40] = rng.uniform(3.0, n - 4.0, (40, 3)).astype(np.float32)       # some surely in the bulk
    vel = (rng.standard_normal((P, 3)) * rng.choice([1e-3, 0.05, 5.0, 40.0], (P, 1))).astype(np.float32)
    radius = np.clip(rng.normal(3.25e-4, 1e-4, P), 1.6e-4, 4.9e-4).astype(np.float32)
    mass = ((np.float32(4.0 / 3.0) * np.float32(3.14159)) * (radius * radius * radius) * np.float32(config.COFFEE_BEAN_DENSITY)).astype(np.float32)
    mass[::17] = 0.0
    active = (rng.random(P) < 0.9).astype(np.int32)
    ps.position.from_numpy(pos); ps.velocity.from_numpy(vel); ps.radius.from_numpy(radius); ps.mass.from_numpy(mass); ps.active.from_numpy(active)
    old = (1e-9 * rng.standard_normal((P, 3))).astype(np.float32)
    ps.drag_force_old.from_numpy(old)
    with quiet():
        ps.compute_two_way_coupling_forces(s.u)
        res.update(p_pos=pos, p_vel=vel, p_radius=radius, p_mass=mass, p_active=active, p_drag_old_in=old,
                   p_drag_new=ps.drag_force_new.to_numpy(), p_reaction=ps.reaction_force_field.to_numpy(),
                   p_u_fluid=ps.fluid_velocity_at_particle.to_numpy(), p_reynolds=ps.particle_reynolds.to_numpy(),
                   p_cd=ps.drag_coefficient.to_numpy())
        ps.apply_under_relaxation(0.8)
        res.update(p_drag=ps.drag_force.to_numpy(), p_drag_old_out=ps.drag_force_old.to_numpy())
        # integrator: three calls (dt clamp on the 2nd / 3rd), forces set before each
        b = fp.get_coffee_bed_boundary()
        res["bounds"] = np.array([b["center_x"], b["center_y"], b["bottom_z"], b["bottom_radius_lu"], b["top_radius_lu"]], np.float64)
        force = (rng.standard_normal((P, 3)) * rng.choice([1e-9, 1e-6, 1e-2], (P, 1))).astype(np.float32)
        ps.force.from_numpy(force)
        res["p_force_in"] = force
        dts = (5e-3, 1.0, 1e-12)
        for t, dt in enumerate(dts):
            ps.update_particle_physics(dt, b["center_x"], b["center_y"], b["bottom_z"], b["bottom_radius_lu"], b["top_radius_lu"])
            res[f"adv{t}_pos"] = ps.position.to_numpy(); res[f"adv{t}_vel"] = ps.velocity.to_numpy(); res[f"adv{t}_active"] = ps.active.to_numpy()
        res["adv_dts"] = np.array(dts); res["adv_counters"] = np.array([ps.coordinate_errors[None], ps.boundary_violations[None]])
    return res


def check_neighbours_against_oracle(res):
    from oracle import d3q19_ref as R
    n = int(res["n"])
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    st = R.init_fields(cfg); R.attach_filter_system(st)
    st.rho = res["rho"].copy(); st.u = res["u"].copy()
    ok = {}
    ok["solid"] = np.array_equal(st.solid, res["solid"]); ok["zone"] = np.array_equal(st.filter_zone, res["filter_zone"])
    st.body_force[:] = 0; R.accumulate_pressure_force(st, R.pressure_gradient_force(st, 0.12), 1.0)
    ok["force_drive"] = np.array_equal(st.body_force, res["bf_force_drive"])
    st.body_force[:] = 0; R.accumulate_pressure_force(st, R.pressure_gradient_force(st, 0.12), 0.5)
    ok["mixed_drive"] = np.array_equal(st.body_force, res["bf_mixed_drive"])
    R.compute_forchheimer_resistance(st)
    ok["forchheimer"] = np.array_equal(st.body_force, res["bf_mixed_plus_forchheimer"])
    dn, react, ufl, re_p, cd, cell = R.two_way_coupling(cfg, st.u, res["p_pos"], res["p_vel"], res["p_radius"], res["p_mass"], res["p_active"])
    act = res["p_active"] != 0
    ok["drag_new"] = np.array_equal(dn[act], res["p_drag_new"][act]); ok["u_fluid"] = np.array_equal(ufl[act], res["p_u_fluid"][act])
    ok["reynolds"] = np.array_equal(re_p[act], res["p_reynolds"][act])
    ok["cd_close"] = bool(np.allclose(cd[act], res["p_cd"][act], rtol=3e-7, atol=0))
    ok["reaction_close"] = bool(np.allclose(react, res["p_reaction"], rtol=1e-5, atol=1e-12))
    drag, new_old = R.under_relax(res["p_drag_new"], res["p_drag_old_in"], res["p_active"], 0.8)
    ok["under_relax"] = np.array_equal(drag[act], res["p_drag"][act]) and np.array_equal(new_old[act], res["p_drag_old_out"][act])
    pos = res["p_pos"].copy(); vel = res["p_vel"].copy(); force = res["p_force_in"].copy(); active = res["p_active"].copy()
    cx, cy, bz, br, tr = [float(v) for v in res["bounds"]]
    tot = [0, 0]; adv = True
    for t, dt in enumerate(res["adv_dts"]):
        ce, bv = R.update_particle_physics(cfg, pos, vel, force, res["p_mass"], active, float(dt), cx, cy, bz, br, tr)
        tot[0] += ce; tot[1] += bv
        a = active == 1
        adv &= np.array_equal(active, res[f"adv{t}_active"]) and np.array_equal(pos[a], res[f"adv{t}_pos"][a]) and \
            np.array_equal(vel[a], res[f"adv{t}_vel"][a], equal_nan=True)
    ok["integrator"] = bool(adv); ok["integrator_counters"] = tot == [int(v) for v in res["adv_counters"]]
    return ok


def run_multiphase_scenario(config, n, seed):
    """The per-step producers next to the LBM step in main.py (main.py:770-800, 839): MultiphaseFlow3D's surface-tension
    chain and phase-field step, PrecisePouringSystem's nozzle force and gradual phase change, on one seeded state with
    the V60 mask.  (main.py itself calls apply_pouring_force with four arguments, which raises and is swallowed by its
    try/except, main.py:778-783; the method is driven here with its declared signature.)"""
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
    import helpers as H
    with quiet():
        from src.core.legacy.lbm_solver import LBMSolver
        from src.physics.filter_paper import FilterPaperSystem
        from src.core.multiphase_3d import MultiphaseFlow3D
        from src.physics.precise_pouring import PrecisePouringSystem
        s = LBMSolver(); s.init_fields()
        fp = FilterPaperSystem(s); fp.initialize_filter_geometry()
        mp = MultiphaseFlow3D(s)
        pp = PrecisePouringSystem()
    rng = np.random.default_rng(seed)
    x = np.arange(n, dtype=np.float32)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    # a wavy free surface through the cone, some noise, saturated cells on both sides, a flat patch (zero gradient)
    phi = np.tanh((6.5 - Z + 0.8 * np.sin(0.7 * X) + 0.5 * np.cos(0.9 * Y)) / 1.5).astype(np.float32)
    phi += (0.03 * rng.standard_normal(phi.shape)).astype(np.float32)
    phi = np.clip(phi, -1.0, 1.0).astype(np.float32)
    phi[2:6, 2:6, 2:5] = 0.25
    u = H.smooth_velocity(n, 0.05, seed) + (0.01 * rng.standard_normal((n, n, n, 3))).astype(np.float32)
    u = u.astype(np.float32); u[3, 4, 5] = 0.0
    rho = (1.0 + 0.05 * rng.standard_normal((n, n, n))).astype(np.float32)
    rho[7, 7, 7] = 0.0; rho[8, 7, 6] = 5e-11                    # the rho > 1e-10 guard of apply_surface_tension
    bf = (1e-4 * rng.standard_normal((n, n, n, 3))).astype(np.float32)
    phi_new0 = (0.1 * rng.standard_normal((n, n, n))).astype(np.float32)   # the outer layer of phi_new is never written by the update
    mp.phi.from_numpy(phi); mp.phi_new.from_numpy(phi_new0); s.u.from_numpy(u); s.rho.from_numpy(rho); s.body_force.from_numpy(bf)
    with quiet():
        mp.compute_chemical_potential()                         # a non-zero mu for the diffusion term (an input here)
    res = dict(n=n, solid=s.solid.to_numpy().astype(np.uint8), phi=phi, phi_new_in=phi_new0, mu=mp.mu.to_numpy(),
               laplacian_phi=mp.laplacian_phi.to_numpy(), interface_width=float(mp.INTERFACE_WIDTH), u=u, rho=rho, body_force=bf,
               cfg_pour_diameter_grid=float(pp.POUR_DIAMETER_GRID), cfg_pour_height=int(pp.POUR_HEIGHT),
               sigma=float(mp.SURFACE_TENSION_COEFF), mobility=float(mp.MOBILITY), dt=float(config.DT),
               rho_water=float(config.RHO_WATER), rho_air=float(config.RHO_AIR))
    with quiet():
        mp.accumulate_surface_tension_pre_collision()
    res.update(st_grad_phi=mp.grad_phi.to_numpy(), st_grad_mu=mp.grad_mu.to_numpy(), st_normal=mp.normal.to_numpy(),
               st_curvature=mp.curvature.to_numpy(), st_surface_force=mp.surface_force.to_numpy(), st_body_force=s.body_force.to_numpy())
    with quiet():
        mp.step(20, precollision_applied=True)
    res.update(s1_phi=mp.phi.to_numpy(), s1_phi_new=mp.phi_new.to_numpy(), s1_rho=s.rho.to_numpy(), s1_phase=s.phase.to_numpy(),
               s1_body_force=s.body_force.to_numpy())
    with quiet():
        mp.step(21, precollision_applied=False)
    res.update(s2_phi=mp.phi.to_numpy(), s2_rho=s.rho.to_numpy(), s2_phase=s.phase.to_numpy(), s2_body_force=s.body_force.to_numpy(),
               s2_surface_force=mp.surface_force.to_numpy(), s2_curvature=mp.curvature.to_numpy())
    # ---- pouring: a nozzle that covers a few cells of the 16^3 box (diameter and height are plain attributes) ----
    pp.POUR_DIAMETER_GRID = 5.0
    pp.POUR_HEIGHT = 10
    res.update(pour_diameter=float(pp.POUR_DIAMETER_GRID), pour_height=int(pp.POUR_HEIGHT), pour_velocity=float(pp.POUR_VELOCITY))
    calls = []
    with quiet():
        pp.start_pouring(pattern="center", flow_rate=0.3)
        pp.apply_pouring_force(s.body_force, s.solid, 0.1); calls.append(("center", 0.3, 0.1))
        pp.apply_gradual_phase_change(mp.phi, s.solid, 0.1)
        res.update(p1_body_force=s.body_force.to_numpy(), p1_phi=mp.phi.to_numpy())
        pp.start_pouring(pattern="spiral", flow_rate=1.0)
        for dt in (0.5, 1e-3, 1e-9):            # plain; acceleration capped at 10; dt below the 1e-8 guard
            pp.apply_pouring_force(s.body_force, s.solid, dt)
            pp.apply_gradual_phase_change(mp.phi, s.solid, dt)
        res.update(p2_body_force=s.body_force.to_numpy(), p2_phi=mp.phi.to_numpy(), p2_pour_time=float(pp.pour_time[None]),
                   p2_dts=np.array([0.5, 1e-3, 1e-9]))
        pp.stop_pouring()
        pp.apply_pouring_force(s.body_force, s.solid, 1.0)
        res.update(p3_body_force=s.body_force.to_numpy())
    return res


def check_multiphase_against_oracle(res):
    from oracle import producers_ref as P
    n = int(res["n"])
    m = P.MultiphaseState(n)
    m.phi = res["phi"].copy(); m.phi_new = res["phi_new_in"].copy()
    u = res["u"]; rho = res["rho"].copy(); bf = res["body_force"].copy(); solid = res["solid"]; phase = np.zeros_like(rho)
    sig, mob, dt, rw, ra = (float(res[k]) for k in ("sigma", "mobility", "dt", "rho_water", "rho_air"))
    ok = {}
    lap = P.compute_chemical_potential(m, sig, float(res["interface_width"]))
    ok["chemical_potential"] = np.array_equal(m.mu, res["mu"]) and np.array_equal(lap, res["laplacian_phi"])
    P.accumulate_surface_tension_pre_collision(m, rho, solid, bf, sig)
    for k, a in (("grad_phi", m.grad_phi), ("grad_mu", m.grad_mu), ("normal", m.normal), ("curvature", m.curvature),
                 ("surface_force", m.surface_force), ("body_force", bf)):
        ok["st_" + k] = np.array_equal(a, res["st_" + k])
    P.multiphase_step(m, u, rho, phase, solid, bf, sig, mob, dt, rw, ra, 20, True)
    ok["s1"] = all(np.array_equal(a, res[k]) for k, a in (("s1_phi", m.phi), ("s1_phi_new", m.phi_new), ("s1_rho", rho), ("s1_phase", phase),
                                                         ("s1_body_force", bf)))
    P.multiphase_step(m, u, rho, phase, solid, bf, sig, mob, dt, rw, ra, 21, False)
    ok["s2"] = all(np.array_equal(a, res[k]) for k, a in (("s2_phi", m.phi), ("s2_rho", rho), ("s2_phase", phase), ("s2_body_force", bf),
                                                         ("s2_surface_force", m.surface_force), ("s2_curvature", m.curvature)))
    p = P.PourState(n, float(res["pour_diameter"]), int(res["pour_height"]), float(res["pour_velocity"]))
    p.start_pouring(pattern="center", flow_rate=0.3)
    P.apply_pouring_force(p, bf, solid, 0.1); P.apply_gradual_phase_change(p, m.phi, solid, 0.1)
    ok["pour_center"] = np.array_equal(bf, res["p1_body_force"]) and np.array_equal(m.phi, res["p1_phi"])
    p.start_pouring(pattern="spiral", flow_rate=1.0)
    for d in res["p2_dts"]:
        P.apply_pouring_force(p, bf, solid, float(d)); P.apply_gradual_phase_change(p, m.phi, solid, float(d))
    ok["pour_spiral"] = np.array_equal(bf, res["p2_body_force"]) and np.array_equal(m.phi, res["p2_phi"]) and \
        float(p.pour_time) == float(res["p2_pour_time"])
    p.active = 0
    P.apply_pouring_force(p, bf, solid, 1.0)
    ok["pour_stopped"] = np.array_equal(bf, res["p3_body_force"])
    ok["nozzle_cells"] = int((res["p1_body_force"] != res["s2_body_force"]).any(-1).sum()) > 10
    return ok


def run_filter_particle_scenario(config, n, seed):
    """FilterPaperSystem.block_particles_at_filter + update_dynamic_resistance (filter_paper.py:616-746; the particle half of
    FilterPaperSystem.step :748-790).  ti.random() -- an unseeded stream in Taichi -- is pinned to 0.5 for the recording, so the
    horizontal kick is zero and everything recorded is deterministic.  The method divides positions by SCALE_LENGTH, so the
    scenario places the particles at (lattice coordinate x SCALE_LENGTH) to reach the filter zone at all."""
    import taichi
    taichi.random = lambda dt=None: np.float32(0.5)
    with quiet():
        from src.core.legacy.lbm_solver import LBMSolver
        from src.physics.filter_paper import FilterPaperSystem
        from src.physics.coffee_particles import CoffeeParticleSystem
        s = LBMSolver(); s.init_fields()
        fp = FilterPaperSystem(s); fp.initialize_filter_geometry()
        P = 300
        ps = CoffeeParticleSystem(P)
    rng = np.random.default_rng(seed)
    zone = fp.filter_zone.to_numpy().astype(np.int32)
    cells = np.argwhere(zone == 1)
    pick = cells[rng.integers(0, len(cells), 200)].astype(np.float64)
    pick[:, 2] += rng.integers(-3, 4, 200)                                   # within / beyond the 5-plane search window
    rest = rng.uniform(-2.0, n + 2.0, (P - 200, 3))
    lat = np.concatenate([pick + rng.uniform(0.05, 0.95, pick.shape), rest])
    pos = (lat * config.SCALE_LENGTH).astype(np.float32)
    pos[-5:] = rng.uniform(0.0, n, (5, 3)).astype(np.float32)                # lattice-unit positions, as main.py feeds them: far outside
    vel = (0.05 * rng.standard_normal((P, 3))).astype(np.float32)
    active = (rng.random(P) < 0.9).astype(np.int32)
    radius = np.full(P, 3.25e-4, np.float32)
    ps.position.from_numpy(pos); ps.velocity.from_numpy(vel); ps.radius.from_numpy(radius); ps.active.from_numpy(active)
    ps.particle_count[None] = P
    acc0 = np.where(zone == 1, rng.uniform(0.0, 30.0, zone.shape), 0.0).astype(np.float32)
    blk0 = np.where(zone == 1, rng.uniform(0.0, 0.5, zone.shape), 0.0).astype(np.float32)
    fp.accumulated_particles.from_numpy(acc0); fp.filter_blockage.from_numpy(blk0)
    res = dict(n=n, scale_length=float(config.SCALE_LENGTH), filter_zone=zone, p_pos=pos, p_vel=vel, p_active=active, p_radius=radius,
               accumulated_in=acc0, blockage_in=blk0)
    with quiet():
        for t in range(2):
            fp.block_particles_at_filter(ps.position, ps.velocity, ps.radius, ps.active, ps.particle_count)
            res[f"b{t}_vel"] = ps.velocity.to_numpy(); res[f"b{t}_accumulated"] = fp.accumulated_particles.to_numpy()
        for t in range(2):
            fp.update_dynamic_resistance()
            res[f"r{t}_blockage"] = fp.filter_blockage.to_numpy(); res[f"r{t}_accumulated"] = fp.accumulated_particles.to_numpy()
    # ---- CoffeeParticleSystem.apply_fluid_forces (coffee_particles.py:547-639): the producer of the integrator's force ----
    P2 = 400
    with quiet():
        ps2 = CoffeeParticleSystem(P2)
    u = (rng.standard_normal((n, n, n, 3)) * rng.choice([1e-7, 0.02, 0.5, 30.0, 200.0], (n, n, n, 1), p=[0.1, 0.5, 0.3, 0.05, 0.05])).astype(np.float32)
    pos2 = rng.uniform(-1.0, n + 1.0, (P2, 3)).astype(np.float32)
    pos2[:150] = rng.uniform(0.5, n - 1.5, (150, 3)).astype(np.float32)
    vel2 = (rng.standard_normal((P2, 3)) * rng.choice([1e-3, 0.05, 3.0, 40.0], (P2, 1))).astype(np.float32)
    rad2 = np.clip(rng.normal(3.25e-4, 1e-4, P2), 1.6e-4, 4.9e-4).astype(np.float32)
    rad2[::23] = 0.02; rad2[5::29] = 1e-6                                       # validate_radius rejects both
    mass2 = ((np.float32(4.0 / 3.0) * np.float32(3.14159)) * (rad2 * rad2 * rad2) * np.float32(config.COFFEE_BEAN_DENSITY)).astype(np.float32)
    mass2[::17] = 0.0; mass2[3::31] = 1e-12                                     # mass guard; tiny mass -> the force caps
    act2 = (rng.random(P2) < 0.9).astype(np.int32)
    force0 = (1e-7 * rng.standard_normal((P2, 3))).astype(np.float32)
    s.u.from_numpy(u)
    ps2.position.from_numpy(pos2); ps2.velocity.from_numpy(vel2); ps2.radius.from_numpy(rad2); ps2.mass.from_numpy(mass2)
    ps2.active.from_numpy(act2); ps2.force.from_numpy(force0)
    with quiet():
        ps2.apply_fluid_forces(s.u, s.u, s.u, s.rho, s.rho, 0.01)
    res.update(ff_u=u, ff_pos=pos2, ff_vel=vel2, ff_radius=rad2, ff_mass=mass2, ff_active=act2, ff_force_in=force0,
               ff_force=ps2.force.to_numpy(), ff_vel_out=ps2.velocity.to_numpy(), ff_active_out=ps2.active.to_numpy(),
               ff_errors=int(ps2.coordinate_errors[None]), water_density=float(ps2.water_density), water_viscosity=float(ps2.water_viscosity),
               particle_gravity=float(ps2.gravity))
    return res


def check_filter_particles_against_oracle(res):
    from oracle import producers_ref as P
    vel = res["p_vel"].copy(); acc = res["accumulated_in"].copy(); blk = res["blockage_in"].copy()
    ok = {}
    for t in range(2):
        P.block_particles_at_filter(res["filter_zone"], res["p_pos"], vel, res["p_active"], acc, float(res["scale_length"]), noise=0.0)
        ok[f"block{t}"] = np.array_equal(vel, res[f"b{t}_vel"]) and np.array_equal(acc, res[f"b{t}_accumulated"])
    for t in range(2):
        P.update_dynamic_resistance(res["filter_zone"], blk, acc)
        ok[f"resist{t}"] = np.array_equal(blk, res[f"r{t}_blockage"]) and np.array_equal(acc, res[f"r{t}_accumulated"])
    vel = res["ff_vel"].copy(); act = res["ff_active"].copy(); force = res["ff_force_in"].copy()
    err = P.apply_fluid_forces(res["ff_u"], res["ff_pos"], vel, res["ff_radius"], res["ff_mass"], act, force, float(res["water_density"]),
                               float(res["water_viscosity"]), float(res["particle_gravity"]))
    ok["fluid_forces"] = np.array_equal(force, res["ff_force"]) and np.array_equal(vel, res["ff_vel_out"]) and \
        np.array_equal(act, res["ff_active_out"]) and err == int(res["ff_errors"])
    ok["fluid_forces_cover"] = int((res["ff_force"] != res["ff_force_in"]).any(-1).sum()) > 50 and err > 10
    ok["bounced"] = int((res["b0_vel"][:, 2] != res["p_vel"][:, 2]).sum()) > 20
    ok["second_pass_quiet"] = np.array_equal(res["b0_vel"], res["b1_vel"])      # after the bounce v_z > 0: nothing more happens
    return ok


def run_coupled_scenario(config, n, seed, steps=6):
    """BASELINE configs[3] as the reference wires it: LBMSolver.step_with_two_way_coupling(particle_system, dt, relax)
    (legacy/lbm_solver.py:1485-1509: clear body force -> coupling on the current u -> under-relaxation -> reaction into the
    body force -> step) followed by CoffeeParticleSystem.update_particle_physics with the filter's bounds (main.py:672-679),
    `steps` times on the V60 box, air-phase relaxation."""
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
    import helpers as H
    gravity = 2e-5
    config.GRAVITY_LU = gravity
    with quiet():
        from src.core.legacy.lbm_solver import LBMSolver
        from src.physics.filter_paper import FilterPaperSystem
        from src.physics.coffee_particles import CoffeeParticleSystem
        s = LBMSolver(); s.init_fields()
        fp = FilterPaperSystem(s); fp.initialize_filter_geometry()
        s.boundary_manager.set_filter_system(fp)
        P = 250
        ps = CoffeeParticleSystem(P)
    st = H.reference_v60_state(n, seed=seed, gravity=gravity, body=1e-5, phase_mode="none")
    st.phase[:] = np.random.default_rng(seed + 1000).uniform(0.0, 0.5, size=st.phase.shape).astype(np.float32)
    rng = np.random.default_rng(seed + 7)
    fluid_cells = np.argwhere(st.solid == 0)
    pos = (fluid_cells[rng.integers(0, len(fluid_cells), P)] + rng.uniform(0.05, 0.95, (P, 3))).astype(np.float32)
    pos[-20:] = rng.uniform(-1.0, n + 1.0, (20, 3)).astype(np.float32)
    vel = (0.02 * rng.standard_normal((P, 3))).astype(np.float32)
    radius = np.clip(rng.normal(3.25e-4, 1e-4, P), 1.6e-4, 4.9e-4).astype(np.float32)
    mass = ((np.float32(4.0 / 3.0) * np.float32(3.14159)) * (radius * radius * radius) * np.float32(config.COFFEE_BEAN_DENSITY)).astype(np.float32)
    active = (rng.random(P) < 0.92).astype(np.int32)
    s.f.from_numpy(st.f); s.f_new.from_numpy(st.f); s.phase.from_numpy(st.phase)
    ps.position.from_numpy(pos); ps.velocity.from_numpy(vel); ps.radius.from_numpy(radius); ps.mass.from_numpy(mass); ps.active.from_numpy(active)
    b = fp.get_coffee_bed_boundary()
    bounds = np.array([b["center_x"], b["center_y"], b["bottom_z"], b["bottom_radius_lu"], b["top_radius_lu"]], np.float64)
    dt_p = 5e-3
    res = dict(n=n, steps=steps, gravity=gravity, seed=seed, f=st.f.copy(), phase=st.phase.copy(), solid=s.solid.to_numpy().astype(np.uint8),
               p_pos=pos, p_vel=vel, p_radius=radius, p_mass=mass, p_active=active, bounds=bounds, dt_particles=dt_p, relax=0.8)
    with quiet():
        for _ in range(steps):
            s.step_with_two_way_coupling(ps, 1.0, 0.8)
            ps.update_particle_physics(dt_p, b["center_x"], b["center_y"], b["bottom_z"], b["bottom_radius_lu"], b["top_radius_lu"])
    res.update(rho=s.rho.to_numpy(), u=s.u.to_numpy(), f_out=s.f.to_numpy(), body_force=s.body_force.to_numpy(),
               p_pos_out=ps.position.to_numpy(), p_vel_out=ps.velocity.to_numpy(), p_active_out=ps.active.to_numpy(),
               p_drag=ps.drag_force.to_numpy(), p_drag_old=ps.drag_force_old.to_numpy(), p_reynolds=ps.particle_reynolds.to_numpy(),
               p_reaction=ps.reaction_force_field.to_numpy())
    return res


def oracle_coupled(res):
    """The same sequence from oracle/ functions; returns (State, pos, vel, active, drag, drag_old, reaction)."""
    from oracle import d3q19_ref as R
    n = int(res["n"])
    cfg = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=float(res["gravity"]))
    st = R.init_fields(cfg); R.attach_filter_system(st)
    st.f = res["f"].copy(); st.f_new = res["f"].copy(); st.phase = res["phase"].copy()
    pos, vel, active = res["p_pos"].copy(), res["p_vel"].copy(), res["p_active"].copy()
    drag_old = np.zeros_like(pos); drag = np.zeros_like(pos); react = None
    force = np.zeros_like(pos)
    cx, cy, bz, br, tr = [float(v) for v in res["bounds"]]
    for _ in range(int(res["steps"])):
        st.body_force[:] = 0
        dn, react, ufl, re_p, cd, cell = R.two_way_coupling(cfg, st.u, pos, vel, res["p_radius"], res["p_mass"], active, sequential=True)
        new_drag, new_old = R.under_relax(dn, drag_old, active, float(res["relax"]))
        a = active != 0
        drag[a] = new_drag[a]; drag_old[a] = new_old[a]
        R.add_particle_reaction_forces(st, react)
        R.step(st)
        R.update_particle_physics(cfg, pos, vel, force, res["p_mass"], active, float(res["dt_particles"]), cx, cy, bz, br, tr)
    return st, pos, vel, active, drag, drag_old, react, re_p


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    config = load_reference(n)
    import time
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import d3q19_ref as R
    all_ok = True
    if len(sys.argv) > 2 and sys.argv[2] == "long":
        # BASELINE's "rho and u after 1000 steps" criterion against the reference's own code: ~50 min of emulation at 16^3
        steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
        t = time.time()
        inp, geom, out, st = run_step_scenario(config, n, steps, seed=36, gravity=2e-5, phase_mode="air_random")
        from oracle import ref_cpu as RC
        cs = RC.CState(st); cs.step(steps)
        fluid = geom["solid"] == 0
        ok = np.array_equal(out["rho"][fluid], cs.rho[fluid]) and np.array_equal(out["u"][fluid], cs.u[fluid]) and \
            np.array_equal(out["f_out"][:, fluid], cs.f[:, fluid])
        print(f"[reference run] air_random_{steps}: n={n} reference {time.time() - t:.0f} s  C oracle bit-exact: {ok}  max|u| {np.abs(out['u'][fluid]).max():.3e}")
        np.savez_compressed(os.path.join(HERE, f"reference_run_long_air_{steps}.npz"), n=n, steps=steps, gravity=2e-5, seed=36,
                            phase_mode="air_random", **inp, **geom, **out)
        sys.exit(0 if ok else 1)
    if len(sys.argv) > 2 and sys.argv[2] == "long2":
        # a second long run in the legacy solver's stable (air) regime on another grid: `20 long2 1000` = 20^3 (1591 fluid cells), 50 x the
        # gravity of the first one (rho spreads over [0.996, 1.007]), another seed; ~100 min of emulation
        steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
        seed, gravity = 41, 1e-3
        t = time.time()
        inp, geom, out, st = run_step_scenario(config, n, steps, seed=seed, gravity=gravity, phase_mode="air_random")
        from oracle import ref_cpu as RC
        cs = RC.CState(st); cs.step(steps)
        fluid = geom["solid"] == 0
        ok = np.array_equal(out["rho"][fluid], cs.rho[fluid]) and np.array_equal(out["u"][fluid], cs.u[fluid]) and \
            np.array_equal(out["f_out"][:, fluid], cs.f[:, fluid])
        print(f"[reference run] air_random_{steps} n={n} g={gravity}: reference {time.time() - t:.0f} s  C oracle bit-exact: {ok}  "
              f"max|u| {np.abs(out['u'][fluid]).max():.3e}  rho [{out['rho'][fluid].min():.5f}, {out['rho'][fluid].max():.5f}]")
        np.savez_compressed(os.path.join(HERE, f"reference_run_long_air_{steps}_n{n}.npz"), n=n, steps=steps, gravity=gravity, seed=seed,
                            phase_mode="air_random", **inp, **geom, **out)
        sys.exit(0 if ok else 1)
    if len(sys.argv) > 2 and sys.argv[2] == "coupled_long":
        # the coupled sequence over ten times as many steps (particles cross cells, the cone constraint and the deactivation of stray
        # particles act repeatedly): the oracle must follow the reference's run through all of them
        steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
        t = time.time()
        res = run_coupled_scenario(config, n, seed=57, steps=steps)
        st, pos, vel, active, drag, drag_old, react, re_p = oracle_coupled(res)
        fluid = res["solid"] == 0; a = res["p_active_out"] == 1
        ok = dict(rho=np.array_equal(st.rho[fluid], res["rho"][fluid]), u=np.array_equal(st.u[fluid], res["u"][fluid]),
                  f=np.array_equal(st.f[:, fluid], res["f_out"][:, fluid]), active=np.array_equal(active, res["p_active_out"]),
                  pos=np.array_equal(pos[a], res["p_pos_out"][a]), vel=np.array_equal(vel[a], res["p_vel_out"][a]),
                  drag=bool(np.allclose(drag[a], res["p_drag"][a], rtol=1e-6, atol=1e-20)),
                  reaction=bool(np.allclose(react, res["p_reaction"], rtol=1e-5, atol=1e-14)))
        print(f"[reference run] coupled step x{steps} ({time.time() - t:.0f} s) vs oracle:", ok, " active in / out:", int(res["p_active"].sum()),
              int(res["p_active_out"].sum()), " max|reaction|", float(np.abs(res["p_reaction"]).max()))
        np.savez_compressed(os.path.join(HERE, f"reference_run_coupled_{steps}.npz"), **res)
        sys.exit(0 if all(ok.values()) else 1)
    if len(sys.argv) > 2 and sys.argv[2] == "coupled":
        t = time.time()
        res = run_coupled_scenario(config, n, seed=51)
        st, pos, vel, active, drag, drag_old, react, re_p = oracle_coupled(res)
        fluid = res["solid"] == 0; a = res["p_active_out"] == 1
        ok = dict(rho=np.array_equal(st.rho[fluid], res["rho"][fluid]), u=np.array_equal(st.u[fluid], res["u"][fluid]),
                  f=np.array_equal(st.f[:, fluid], res["f_out"][:, fluid]), active=np.array_equal(active, res["p_active_out"]),
                  pos=np.array_equal(pos[a], res["p_pos_out"][a]), vel=np.array_equal(vel[a], res["p_vel_out"][a]),
                  drag=bool(np.allclose(drag[a], res["p_drag"][a], rtol=1e-6, atol=1e-20)),
                  reaction=bool(np.allclose(react, res["p_reaction"], rtol=1e-5, atol=1e-14)))
        print(f"[reference run] coupled step x{int(res['steps'])} ({time.time() - t:.0f} s) vs oracle:", ok,
              " moving particles:", int((res["p_reynolds"] > 0).sum()), " max|reaction|", float(np.abs(res["p_reaction"]).max()))
        np.savez_compressed(os.path.join(HERE, "reference_run_coupled.npz"), **res)
        sys.exit(0 if all(ok.values()) else 1)
    if len(sys.argv) > 2 and sys.argv[2] == "density_drive":
        # PressureGradientDrive method A (pressure_gradient_drive.py:54-122): the target profile and three nudges of rho
        with quiet():
            from src.core.legacy.lbm_solver import LBMSolver
            from src.physics.filter_paper import FilterPaperSystem
            from src.physics.pressure_gradient_drive import PressureGradientDrive
            s = LBMSolver(); s.init_fields()
            fp = FilterPaperSystem(s); fp.initialize_filter_geometry()
            pg = PressureGradientDrive(s)
        rng = np.random.default_rng(53)
        rho = (1.0 + 0.6 * rng.standard_normal((n, n, n))).astype(np.float32)      # well outside [0.5, 2] in places
        rho[::5, ::3, ::2] = (1.0 + 1e-3 * rng.standard_normal(rho[::5, ::3, ::2].shape)).astype(np.float32)   # and inside the 0.001 band
        s.rho.from_numpy(rho)
        res = dict(n=n, rho=rho, solid=s.solid.to_numpy().astype(np.uint8), target=pg.target_density.to_numpy())
        with quiet():
            pg.activate_density_drive(True)
            for t in range(3):
                pg.apply(t)
                res[f"rho_after_{t + 1}"] = s.rho.to_numpy()
        changed = int((res["rho_after_3"] != rho).sum())
        print("[reference run] density drive: cells changed", changed, "of", rho.size, " target profile", res["target"][0, 0, :])
        np.savez_compressed(os.path.join(HERE, "reference_run_density_drive.npz"), **res)
        sys.exit(0 if changed > 0 else 1)
    if len(sys.argv) > 2 and sys.argv[2] == "producers":
        res = run_multiphase_scenario(config, n, seed=43)
        ok = check_multiphase_against_oracle(res)
        print("[reference run] multiphase / pouring producers vs oracle:", ok)
        np.savez_compressed(os.path.join(HERE, "reference_run_multiphase.npz"), **res)
        res2 = run_filter_particle_scenario(config, n, seed=47)
        ok2 = check_filter_particles_against_oracle(res2)
        print("[reference run] filter / particle interception vs oracle:", ok2)
        np.savez_compressed(os.path.join(HERE, "reference_run_filter_particles.npz"), **res2)
        sys.exit(0 if all(ok.values()) and all(ok2.values()) else 1)
    default_gravity = float(config.GRAVITY_LU)
    scenarios = [("split_phase_small_gravity", 4, 31, 2e-5, "split"), ("water_default_gravity", 3, 32, default_gravity, "water"),
                 ("air_phase", 6, 33, 1e-4, "none")]
    for name, steps, seed, gravity, phase_mode in scenarios:
        t = time.time()
        inp, geom, out, st = run_step_scenario(config, n, steps, seed=seed, gravity=gravity, phase_mode=phase_mode)
        geo_ok = np.array_equal(geom["solid"], st.solid) and np.array_equal(geom["filter_zone"], st.filter_zone) and np.array_equal(geom["les_mask"], st.les_mask)
        for _ in range(steps):
            R.step(st)
        fluid = geom["solid"] == 0
        ok = geo_ok and np.array_equal(out["rho"][fluid], st.rho[fluid]) and np.array_equal(out["u"][fluid], st.u[fluid]) and \
            np.array_equal(out["f_out"][:, fluid], st.f[:, fluid])
        nu_active = int((out.get("nu_sgs", np.zeros(1)) > 0).sum())
        print(f"[reference run] {name}: n={n} steps={steps} gravity={gravity:g}  reference {time.time() - t:.1f} s  oracle bit-exact: {ok}  (LES active cells: {nu_active})")
        all_ok &= bool(ok)
        np.savez_compressed(os.path.join(HERE, f"reference_run_step_{name}.npz"), n=n, steps=steps, gravity=gravity, seed=seed,
                            phase_mode=phase_mode, **inp, **geom, **out)
    steps, gravity = 5, 3e-5
    t = time.time()
    inp, out = run_open_box_scenario(config, n, steps, seed=35, gravity=gravity)
    st = oracle_open_box(inp, n, steps, gravity)
    fluid = inp["solid"] == 0
    ok = np.array_equal(out["rho"][fluid], st.rho[fluid]) and np.array_equal(out["u"][fluid], st.u[fluid]) and \
        np.array_equal(out["f_out"][:, fluid], st.f[:, fluid])
    print(f"[reference run] open_box_no_filter: n={n} steps={steps}  reference {time.time() - t:.1f} s  oracle bit-exact: {ok}")
    all_ok &= bool(ok)
    np.savez_compressed(os.path.join(HERE, "reference_run_openbox.npz"), n=n, steps=steps, gravity=gravity, **inp, **out)
    res = run_neighbour_scenario(config, n, seed=41)
    ok = check_neighbours_against_oracle(res)
    print("[reference run] neighbours / particles vs oracle:", ok)
    all_ok &= all(ok.values())
    np.savez_compressed(os.path.join(HERE, "reference_run_neighbours.npz"), **res)
    res = run_multiphase_scenario(config, n, seed=43)
    ok = check_multiphase_against_oracle(res)
    print("[reference run] multiphase / pouring producers vs oracle:", ok)
    all_ok &= all(ok.values())
    np.savez_compressed(os.path.join(HERE, "reference_run_multiphase.npz"), **res)
    res = run_filter_particle_scenario(config, n, seed=47)
    ok = check_filter_particles_against_oracle(res)
    print("[reference run] filter / particle interception vs oracle:", ok)
    all_ok &= all(ok.values())
    np.savez_compressed(os.path.join(HERE, "reference_run_filter_particles.npz"), **res)
    res = run_coupled_scenario(config, n, seed=51)
    st, pos, vel, active, drag, drag_old, react, re_p = oracle_coupled(res)
    fluid = res["solid"] == 0; a = res["p_active_out"] == 1
    ok = np.array_equal(st.rho[fluid], res["rho"][fluid]) and np.array_equal(st.u[fluid], res["u"][fluid]) and \
        np.array_equal(st.f[:, fluid], res["f_out"][:, fluid]) and np.array_equal(pos[a], res["p_pos_out"][a])
    print("[reference run] coupled sequence vs oracle bit-exact:", ok)
    all_ok &= bool(ok)
    np.savez_compressed(os.path.join(HERE, "reference_run_coupled.npz"), **res)
    print("ALL OK" if all_ok else "MISMATCH")
    sys.exit(0 if all_ok else 1)
