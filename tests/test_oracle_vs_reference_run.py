"""The oracle against recorded runs of the REFERENCE'S OWN SOURCE CODE.

tests/golden/reference_run_*.npz were produced by tests/golden/make_reference_goldens.py: the unmodified reference
modules (LBMSolver, FilterPaperSystem, PressureGradientDrive, CoffeeParticleSystem) executed under the pure-Python Taichi
stand-in of tests/golden/taichi_shim on seeded 16^3 states.  Here the oracle restarts from the recorded inputs and must
reproduce the recorded outputs bit for bit.  Nothing in this file reads /root/reference.
"""
import glob
import os

import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEP_FILES = sorted(glob.glob(os.path.join(GOLD, "reference_run_step_*.npz")))


def oracle_state_from_fixture(z):
    n = int(z["n"])
    cfg = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=float(z["gravity"]))
    st = R.init_fields(cfg)
    R.attach_filter_system(st)
    st.f = z["f"].copy(); st.f_new = z["f"].copy(); st.phase = z["phase"].copy(); st.body_force = z["body_force"].copy()
    return st


def test_fixtures_present():
    assert len(STEP_FILES) == 3 and os.path.exists(os.path.join(GOLD, "reference_run_neighbours.npz"))
    assert os.path.exists(os.path.join(GOLD, "reference_run_openbox.npz"))
    assert os.path.exists(os.path.join(GOLD, "reference_run_long_air_1000.npz"))


@pytest.mark.parametrize("path", STEP_FILES, ids=[os.path.basename(p)[19:-4] for p in STEP_FILES])
def test_oracle_step_reproduces_the_reference_run(path):
    """LBMSolver.step() (LES pre-pass, moments, collide + push-stream + bounce-back, copy-swap, filter damping, face BCs)
    with the V60 geometry from FilterPaperSystem: rho, u and f after `steps` calls, plus the geometry masks."""
    z = np.load(path)
    st = oracle_state_from_fixture(z)
    assert np.array_equal(st.solid, z["solid"]) and np.array_equal(st.filter_zone, z["filter_zone"])
    assert np.array_equal(st.les_mask, z["les_mask"])
    for _ in range(int(z["steps"])):
        R.step(st)
    fluid = z["solid"] == 0
    assert fluid.sum() > 500
    assert np.array_equal(st.rho[fluid], z["rho"][fluid])
    assert np.array_equal(st.u[fluid], z["u"][fluid])
    assert np.array_equal(st.f[:, fluid], z["f_out"][:, fluid])
    assert np.abs(z["u"][fluid]).max() > 1e-4           # the run moved
    # the C/OpenMP restatement (the timed CPU baseline) reproduces it as well
    from oracle import ref_cpu as RC
    cs = RC.CState(oracle_state_from_fixture(z))
    cs.step(int(z["steps"]))
    assert np.array_equal(cs.rho[fluid], z["rho"][fluid]) and np.array_equal(cs.u[fluid], z["u"][fluid])
    assert np.array_equal(cs.f[:, fluid], z["f_out"][:, fluid])


LONG_FILES = sorted(glob.glob(os.path.join(GOLD, "reference_run_long_air_*.npz")))


@pytest.mark.parametrize("path", LONG_FILES, ids=[os.path.basename(p)[19:-4] for p in LONG_FILES])
def test_oracle_reproduces_1000_steps_of_the_reference_run(path):
    """BASELINE's criterion "rho and u after 1000 steps" against the reference's own code: 1000 calls of LBMSolver.step()
    (tau_air relaxation -- the stable regime of the legacy solver --, random phase in [0, 0.5] so gravity acts, seeded state; recorded once
    under the Taichi stand-in): V60 16^3 at gravity 2e-5 (~50 min of emulation) and V60 20^3 at gravity 1e-3, another seed (~90 min).
    The C oracle (all 1000 steps) and the NumPy oracle (its own 1000 steps) land on the recorded rho, u, f bit for bit."""
    z = np.load(path)
    steps = int(z["steps"])
    assert steps == 1000
    fluid = z["solid"] == 0
    from oracle import ref_cpu as RC
    cs = RC.CState(oracle_state_from_fixture(z))
    cs.step(steps)
    assert np.array_equal(cs.rho[fluid], z["rho"][fluid]) and np.array_equal(cs.u[fluid], z["u"][fluid])
    assert np.array_equal(cs.f[:, fluid], z["f_out"][:, fluid])
    assert np.isfinite(z["u"]).all() and np.abs(z["u"][fluid]).max() > 1e-6 and abs(float(z["rho"][fluid].mean()) - 1.0) < 0.05
    if int(z["n"]) > 16:
        return                       # the NumPy oracle's own 1000 steps: on the 16^3 recording only (CPU time; C == NumPy is tested separately)
    st = oracle_state_from_fixture(z)
    for _ in range(steps):
        R.step(st)
    assert np.array_equal(st.rho[fluid], z["rho"][fluid]) and np.array_equal(st.u[fluid], z["u"][fluid])
    assert np.array_equal(st.f[:, fluid], z["f_out"][:, fluid])


def test_oracle_open_box_reproduces_the_reference_run():
    """No V60 mask, no filter system (the first 30 steps of main.py): open faces with stale w_q inflow (quirk Q6), face
    rho writes of the boundary manager (quirk Q5), obstacles touching the faces."""
    z = np.load(os.path.join(GOLD, "reference_run_openbox.npz"))
    n, steps = int(z["n"]), int(z["steps"])
    st = R.init_fields(R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=float(z["gravity"])))
    st.solid = z["solid"].copy(); st.phase = z["phase"].copy(); st.body_force = z["body_force"].copy()
    st.f = z["f"].copy(); st.f_new = z["f"].copy(); st.les_mask = z["les_mask"].copy()
    for _ in range(steps):
        R.step(st)
    fluid = z["solid"] == 0
    assert np.array_equal(st.rho[fluid], z["rho"][fluid]) and np.array_equal(st.u[fluid], z["u"][fluid])
    assert np.array_equal(st.f[:, fluid], z["f_out"][:, fluid])
    assert fluid[0].any() and fluid[:, :, -1].any()            # the faces really are fluid


def test_oracle_neighbour_kernels_reproduce_the_reference_run():
    """PressureGradientDrive (force / mixed mode), compute_forchheimer_resistance, two-way particle coupling,
    under-relaxation and the particle integrator (three calls, counters)."""
    z = np.load(os.path.join(GOLD, "reference_run_neighbours.npz"))
    n = int(z["n"])
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    st = R.init_fields(cfg); R.attach_filter_system(st)
    st.rho = z["rho"].copy(); st.u = z["u"].copy()
    assert np.array_equal(st.solid, z["solid"]) and np.array_equal(st.filter_zone, z["filter_zone"])
    st.body_force[:] = 0; R.accumulate_pressure_force(st, R.pressure_gradient_force(st, 0.12), 1.0)
    assert np.array_equal(st.body_force, z["bf_force_drive"])
    st.body_force[:] = 0; R.accumulate_pressure_force(st, R.pressure_gradient_force(st, 0.12), 0.5)
    assert np.array_equal(st.body_force, z["bf_mixed_drive"])
    R.compute_forchheimer_resistance(st)
    assert np.array_equal(st.body_force, z["bf_mixed_plus_forchheimer"])
    assert np.abs(z["bf_mixed_plus_forchheimer"] - z["bf_mixed_drive"]).max() > 0

    act = z["p_active"] != 0
    dn, react, ufl, re_p, cd, cell = R.two_way_coupling(cfg, st.u, z["p_pos"], z["p_vel"], z["p_radius"], z["p_mass"], z["p_active"])
    assert np.array_equal(dn[act], z["p_drag_new"][act]) and np.array_equal(ufl[act], z["p_u_fluid"][act])
    assert np.array_equal(re_p[act], z["p_reynolds"][act])
    assert np.allclose(cd[act], z["p_cd"][act], rtol=3e-7, atol=0)              # powf: libm vs NumPy, <= 2 ulp
    assert np.allclose(react, z["p_reaction"], rtol=1e-5, atol=1e-12)           # scatter order
    drag, new_old = R.under_relax(z["p_drag_new"], z["p_drag_old_in"], z["p_active"], 0.8)
    assert np.array_equal(drag[act], z["p_drag"][act]) and np.array_equal(new_old[act], z["p_drag_old_out"][act])

    pos = z["p_pos"].copy(); vel = z["p_vel"].copy(); force = z["p_force_in"].copy(); active = z["p_active"].copy()
    cx, cy, bz, br, tr = [float(v) for v in z["bounds"]]
    tot = [0, 0]
    for t, dt in enumerate(z["adv_dts"]):
        ce, bv = R.update_particle_physics(cfg, pos, vel, force, z["p_mass"], active, float(dt), cx, cy, bz, br, tr)
        tot[0] += ce; tot[1] += bv
        a = active == 1
        assert np.array_equal(active, z[f"adv{t}_active"])
        assert np.array_equal(pos[a], z[f"adv{t}_pos"][a]) and np.array_equal(vel[a], z[f"adv{t}_vel"][a], equal_nan=True)
    assert tot == [int(v) for v in z["adv_counters"]] and tot[1] > 0


def test_density_drive_reproduces_the_reference_run():
    """PressureGradientDrive method A (target profile + three nudges of rho, pressure_gradient_drive.py:54-122) recorded from the
    reference's own source (make_reference_goldens.py 16 density_drive): oracle restatement bit for bit."""
    z = np.load(os.path.join(GOLD, "reference_run_density_drive.npz"))
    n = int(z["n"])
    target = R.density_drive_target(n)
    assert np.array_equal(z["target"], np.broadcast_to(target[None, None, :], (n, n, n)))
    rho = z["rho"]
    for t in range(3):
        rho = R.density_drive(rho, z["solid"], target)
        assert np.array_equal(rho, z[f"rho_after_{t + 1}"])
    assert (rho != z["rho"]).sum() > 100
