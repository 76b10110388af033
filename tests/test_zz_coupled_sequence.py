"""BASELINE configs[3] as the reference wires it -- LBMSolver.step_with_two_way_coupling(particle_system, dt, relax)
(legacy/lbm_solver.py:1485-1509) followed by CoffeeParticleSystem.update_particle_physics with the filter's bounds
(main.py:672-679), six times on the V60 box -- against a recorded run of the reference's own code
(tests/golden/reference_run_coupled.npz, made by tests/golden/make_reference_goldens.py `coupled`).

CPU: the oracle functions chained in that order reproduce the recording bit for bit (scatter summed in the serial order of
the reference's loop).  GPU: the facade classes over the C ABI; the scatter's atomics are unordered on a GPU, so the flow
fields carry rounding noise of the reaction force: rho, u within BASELINE's 1e-5 relative, particles bit-exact.
"""
import glob
import os

import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run_coupled.npz")
# the oracle is also held against a ten times longer recording of the same sequence (reference_run_coupled_60.npz, another seed); the device
# test stays on the six-step one: its scatter atomics are unordered, and sixty steps of two-way feedback would turn that rounding noise
# into a tolerance question rather than a parity check
GOLDS = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run_coupled*.npz")))


@pytest.mark.parametrize("path", GOLDS, ids=[os.path.basename(p)[14:-4] for p in GOLDS])
def test_oracle_coupled_sequence_reproduces_the_reference_run(path):
    z = np.load(path)
    n = int(z["n"])
    cfg = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=float(z["gravity"]))
    st = R.init_fields(cfg); R.attach_filter_system(st)
    assert np.array_equal(st.solid, z["solid"])
    st.f = z["f"].copy(); st.f_new = z["f"].copy(); st.phase = z["phase"].copy()
    pos, vel, active = z["p_pos"].copy(), z["p_vel"].copy(), z["p_active"].copy()
    drag_old = np.zeros_like(pos); drag = np.zeros_like(pos); force = np.zeros_like(pos)
    cx, cy, bz, br, tr = [float(v) for v in z["bounds"]]
    for _ in range(int(z["steps"])):
        st.body_force[:] = 0                                                       # clear_body_force
        dn, react, ufl, re_p, cd, cell = R.two_way_coupling(cfg, st.u, pos, vel, z["p_radius"], z["p_mass"], active, sequential=True)
        new_drag, new_old = R.under_relax(dn, drag_old, active, float(z["relax"]))
        a = active != 0
        drag[a] = new_drag[a]; drag_old[a] = new_old[a]
        R.add_particle_reaction_forces(st, react)
        R.step(st)
        R.update_particle_physics(cfg, pos, vel, force, z["p_mass"], active, float(z["dt_particles"]), cx, cy, bz, br, tr)
    fluid = z["solid"] == 0; a = z["p_active_out"] == 1
    assert np.array_equal(st.rho[fluid], z["rho"][fluid]) and np.array_equal(st.u[fluid], z["u"][fluid])
    assert np.array_equal(st.f[:, fluid], z["f_out"][:, fluid]) and np.array_equal(st.body_force, z["body_force"])
    assert np.array_equal(active, z["p_active_out"]) and np.array_equal(pos[a], z["p_pos_out"][a]) and np.array_equal(vel[a], z["p_vel_out"][a])
    assert np.allclose(drag[a], z["p_drag"][a], rtol=1e-6, atol=1e-20)              # powf in C_D
    assert np.array_equal(react, z["p_reaction"]) or np.allclose(react, z["p_reaction"], rtol=1e-6, atol=1e-14)
    assert float(np.abs(z["p_reaction"]).max()) > 1e-6 and int((z["p_reynolds"] > 0).sum()) > 100     # the coupling is not a no-op


@pytest.mark.gpu
def test_gpu_coupled_sequence_reproduces_the_reference_run():
    import torch
    from pour_over_coffee_lbm_b200.config import LBMConfig
    from pour_over_coffee_lbm_b200.physics import CoffeeParticleSystem, FilterPaperSystem
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    z = np.load(GOLD)
    n, gravity = int(z["n"]), float(z["gravity"])
    c = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=gravity)
    cfg = LBMConfig(NX=n, NY=n, NZ=n, TAU_FLUID=c.TAU_WATER, TAU_AIR=c.TAU_AIR, GRAVITY_LU=gravity)
    s = LBMSolver(nx=n, ny=n, nz=n, config=cfg, compat="reference", strict=True, gravity_lu=gravity)
    s.init_fields()
    fp = FilterPaperSystem(s); fp.initialize_filter_geometry()
    assert np.array_equal(s.solid.to_numpy(), z["solid"])
    s.f.from_numpy(z["f"]); s.phase.from_numpy(z["phase"])
    npart = z["p_pos"].shape[0]
    ps = CoffeeParticleSystem(npart, solver=s)
    ps.set_particles(z["p_pos"], z["p_vel"], z["p_radius"], z["p_mass"])
    ps.state.active.copy_(torch.from_numpy(z["p_active"]).cuda())
    cx, cy, bz, br, tr = [float(v) for v in z["bounds"]]
    for _ in range(int(z["steps"])):
        s.step_with_two_way_coupling(ps, 1.0, float(z["relax"]))
        ps.update_particle_physics(float(z["dt_particles"]), cx, cy, bz, br, tr)
    fluid = z["solid"] == 0; a = z["p_active_out"] == 1
    rho, u = s.rho.to_numpy(), s.u.to_numpy()
    assert H.rel_err(rho[fluid], z["rho"][fluid]) <= 1e-5 and H.rel_err(u[fluid], z["u"][fluid]) <= 1e-5      # BASELINE's bar
    assert H.rel_err(s.f.to_numpy()[:, fluid], z["f_out"][:, fluid]) <= 1e-5
    assert np.allclose(s.body_force.to_numpy(), z["body_force"], rtol=1e-4, atol=1e-11)
    assert np.array_equal(ps.active.cpu().numpy(), z["p_active_out"])
    assert np.array_equal(ps.position.cpu().numpy()[a], z["p_pos_out"][a]) and np.array_equal(ps.velocity.cpu().numpy()[a], z["p_vel_out"][a])
    assert np.allclose(ps.drag_force.cpu().numpy()[a], z["p_drag"][a], rtol=1e-5, atol=1e-18)


@pytest.mark.parametrize("path", GOLDS, ids=[os.path.basename(p)[14:-4] for p in GOLDS])
def test_emulated_product_kernels_run_the_coupled_loop_of_the_reference(path):
    """BASELINE configs[3] on product kernel SOURCE without a GPU: clear -> particles_couple_kernel (gather, drag, under-relaxation,
    scatter) -> add_reaction_kernel -> the legacy-compatible step kernel -> particles_advance_kernel, compiled by g++ and run thread by
    thread (tests/emu), `steps` times, against the reference's recorded loop (six and sixty steps).  A warp is emulated as lanes that run
    one after the other, so the scatter adds up in particle order like the reference's serial loop: positions, velocities and active
    flags bit for bit; rho, u, f to the rounding of C_D's powf (glibc vs NumPy, <= 2 ulp on the drag)."""
    import ctypes as C
    from pour_over_coffee_lbm_b200.config import LBMConfig
    import test_particles_emulated as TP
    import test_step_reference_emulated as TS
    emu_p = H.build_emu("emu_particles", ['lbm_particles.cu', 'lbm_common.cuh'])
    emu_a = H.build_emu("emu_aux", ["lbm_aux.cu", "lbm_phys.cuh", "lbm_common.cuh"])
    emu_s = H.build_emu("emu_step_reference", TS.DEPS)
    z = np.load(path)
    n, steps, gravity = int(z["n"]), int(z["steps"]), float(z["gravity"])
    c = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=gravity)
    cfg = LBMConfig(NX=n, NY=n, NZ=n, TAU_FLUID=c.TAU_WATER, TAU_AIR=c.TAU_AIR, GRAVITY_LU=gravity)
    k_lu, beta_lu = cfg.forchheimer_parameters(); c_darcy, c_forch = cfg.filter_constants()
    st0 = R.init_fields(c); R.attach_filter_system(st0)
    solid, zone, les_mask = st0.solid, st0.filter_zone, st0.les_mask
    assert np.array_equal(solid, z["solid"])
    nbr = TS.neighbour_masks(solid)
    flags = (solid.astype(np.uint8) * 1) | ((zone != 0).astype(np.uint8) * 2) | ((les_mask != 0).astype(np.uint8) * 4) | ((nbr != 0).astype(np.uint8) * 8)
    d_flags, d_nbr = H.to_dev_scalar(flags), H.to_dev_scalar(nbr)
    bufs = [H.to_dev_pop(TS.f_to_g(z["f"], solid)), None]; bufs[1] = bufs[0].copy()
    d_phase = H.to_dev_scalar(z["phase"])
    d_rho = np.ones((n, n, n), np.float32)
    u_bufs = [np.zeros((3, n, n, n), np.float32), np.zeros((3, n, n, n), np.float32)]
    d_force = np.zeros((3, n, n, n), np.float32); d_react = np.zeros((3, n, n, n), np.float32)
    d_blockage = np.zeros((n, n, n), np.float32)
    ps = TP.State(z["p_pos"], z["p_vel"], z["p_radius"], z["p_mass"], z["p_active"])
    s = ps.struct()
    cx, cy, bz, br, tr = [float(v) for v in z["bounds"]]
    b = TP.Bounds(cx, cy, bz, br, tr, float(np.float32(cfg.CUP_HEIGHT / cfg.SCALE_LENGTH)), float(n), float(n - 5))
    zero_force = np.zeros((3, ps.n), np.float32); counters = np.zeros(2, np.int32)
    rho_w = np.float32(c.WATER_DENSITY_90C); mu_w = np.float32(c.WATER_VISCOSITY_90C * c.WATER_DENSITY_90C)
    f32 = lambda v: C.c_float(float(v)); i32 = lambda v: C.c_int(int(v)); p = TP._p
    cur = 0
    for _ in range(steps):
        d_force[:] = 0                                                                       # LBMSolver.clear_body_force
        emu_p.emu_particles_couple(i32(n), i32(n), i32(n), p(u_bufs[cur]), p(d_react), C.byref(s), f32(rho_w), f32(mu_w), f32(float(z["relax"])))
        emu_a.emu_add_reaction(i32(n), i32(n), i32(n), p(d_react), p(d_flags), p(d_force))
        emu_s.emu_step_reference(i32(n), i32(n), i32(n), p(bufs[cur]), p(bufs[1 - cur]), p(d_rho), p(u_bufs[cur]), p(u_bufs[1 - cur]),
                                 p(d_force), p(d_phase), p(d_blockage), p(d_flags), p(d_nbr), i32(1), i32(1), f32(cfg.TAU_WATER), f32(cfg.TAU_AIR),
                                 f32(gravity), f32(cfg.LES_CS), f32(0.55), f32(1.90), f32(k_lu), f32(beta_lu), f32(c_darcy), f32(c_forch))
        cur = 1 - cur
        emu_p.emu_particles_advance(C.byref(s), p(zero_force), C.byref(b), f32(float(z["dt_particles"])), p(counters))
    fluid = solid == 0; a = z["p_active_out"] == 1
    assert np.array_equal(ps.active, z["p_active_out"])
    assert np.array_equal(ps.pos.T[a], z["p_pos_out"][a]) and np.array_equal(ps.vel.T[a], z["p_vel_out"][a])
    rho = np.transpose(d_rho, (2, 1, 0)); u = np.transpose(u_bufs[cur], (3, 2, 1, 0))
    f_out = TS.g_to_f(np.transpose(bufs[cur], (0, 3, 2, 1)), solid)
    assert np.allclose(rho[fluid], z["rho"][fluid], rtol=1e-6, atol=0) and np.allclose(u[fluid], z["u"][fluid], rtol=1e-5, atol=1e-9)
    assert np.allclose(f_out[:, fluid], z["f_out"][:, fluid], rtol=1e-6, atol=0)
    assert np.allclose(ps.drag.T[a], z["p_drag"][a], rtol=1e-6, atol=1e-20)
