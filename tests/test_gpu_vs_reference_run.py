"""The CUDA path against recorded runs of the REFERENCE'S OWN SOURCE CODE (tests/golden/reference_run_*.npz, produced by
tests/golden/make_reference_goldens.py: the unmodified reference modules executed under a pure-Python Taichi stand-in).

The kernels start from the recorded inputs, go through the C ABI, and must reproduce what the reference's code computed:
bit for bit with the strict build (compat = reference), flags and particle cell data included.
"""
import glob
import os

import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEP_FILES = sorted(glob.glob(os.path.join(GOLD, "reference_run_step_*.npz")))


def _torch(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _engine(*a, **k):
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine
    return D3Q19Engine(*a, **k)


def _cfg(n, gravity):
    from pour_over_coffee_lbm_b200.config import LBMConfig
    c = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=gravity)
    return LBMConfig(NX=n, NY=n, NZ=n, TAU_FLUID=c.TAU_WATER, TAU_AIR=c.TAU_AIR, GRAVITY_LU=gravity)


@pytest.mark.parametrize("vec", [1, 2, 4])
@pytest.mark.parametrize("path", STEP_FILES, ids=[os.path.basename(p)[19:-4] for p in STEP_FILES])
def test_step_kernel_reproduces_the_reference_run(path, vec):
    """lbm_build_v60_geometry + lbm_pack_flags + lbm_import_f + lbm_step (compat = reference, strict) from the recorded
    inputs: geometry masks, rho, u and f equal what LBMSolver.step() of the reference produced."""
    z = np.load(path)
    n, steps, gravity = int(z["n"]), int(z["steps"]), float(z["gravity"])
    eng = _engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, force=True, phase=True, les=True,
                  porous=True, strict=True, vec=vec, config=_cfg(n, gravity), gravity_lu=gravity)
    eng.build_v60_geometry()                                              # FilterPaperSystem.initialize_filter_geometry on the device
    assert np.array_equal(H.from_dev_scalar(eng.solid), z["solid"])
    assert np.array_equal(H.from_dev_scalar(eng.filter_zone), z["filter_zone"])
    assert np.array_equal(H.from_dev_scalar(eng.les_mask), z["les_mask"])
    eng.phase.copy_(_torch(H.to_dev_scalar(z["phase"]))); eng.body_force.copy_(_torch(H.to_dev_vec(z["body_force"])))
    eng.import_f(_torch(H.to_dev_pop(z["f"])))
    eng.step(steps)
    fluid = z["solid"] == 0
    assert np.array_equal(H.from_dev_scalar(eng.rho)[fluid], z["rho"][fluid])
    assert np.array_equal(H.from_dev_vec(eng.u)[fluid], z["u"][fluid])
    assert np.array_equal(H.from_dev_pop(eng.export_f())[:, fluid], z["f_out"][:, fluid])


LONG_FILES = sorted(glob.glob(os.path.join(GOLD, "reference_run_long_air_*.npz")))


@pytest.mark.parametrize("vec", [1, 2, 4])
@pytest.mark.parametrize("path", LONG_FILES, ids=[os.path.basename(p)[19:-4] for p in LONG_FILES])
def test_step_kernel_reproduces_1000_steps_of_the_reference_run(path, vec):
    """BASELINE: "rho and u must agree within 1e-5 relative (fp32) after 1000 steps" -- against 1000 recorded calls of the
    reference's own LBMSolver.step() (tests/golden/reference_run_long_air_1000.npz: V60 16^3; ..._n20.npz: V60 20^3, 50 x the gravity,
    another seed): the strict build agrees bit for bit."""
    z = np.load(path)
    n, steps, gravity = int(z["n"]), int(z["steps"]), float(z["gravity"])
    assert steps == 1000
    eng = _engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, force=True, phase=True, les=True,
                  porous=True, strict=True, vec=vec, config=_cfg(n, gravity), gravity_lu=gravity)
    eng.build_v60_geometry()
    assert np.array_equal(H.from_dev_scalar(eng.solid), z["solid"])
    eng.phase.copy_(_torch(H.to_dev_scalar(z["phase"]))); eng.body_force.copy_(_torch(H.to_dev_vec(z["body_force"])))
    eng.import_f(_torch(H.to_dev_pop(z["f"])))
    eng.step(steps)
    fluid = z["solid"] == 0
    assert np.array_equal(H.from_dev_scalar(eng.rho)[fluid], z["rho"][fluid])
    assert np.array_equal(H.from_dev_vec(eng.u)[fluid], z["u"][fluid])
    assert np.array_equal(H.from_dev_pop(eng.export_f())[:, fluid], z["f_out"][:, fluid])


@pytest.mark.parametrize("vec", [1, 2, 4])
def test_open_box_step_and_face_bc_reproduce_the_reference_run(vec):
    """No V60 mask, no filter system: open faces (stale w_q inflow), boundary-manager face writes (lbm_face_bc), obstacles
    touching the faces -- the first 30 steps of main.py."""
    z = np.load(os.path.join(GOLD, "reference_run_openbox.npz"))
    n, steps, gravity = int(z["n"]), int(z["steps"]), float(z["gravity"])
    eng = _engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, force=True, phase=True, les=True,
                  strict=True, vec=vec, config=_cfg(n, gravity), gravity_lu=gravity)
    eng.solid.copy_(_torch(H.to_dev_scalar(z["solid"]))); eng.les_mask.copy_(_torch(H.to_dev_scalar(z["les_mask"]))); eng.pack_flags()
    eng.phase.copy_(_torch(H.to_dev_scalar(z["phase"]))); eng.body_force.copy_(_torch(H.to_dev_vec(z["body_force"])))
    eng.import_f(_torch(H.to_dev_pop(z["f"])))
    for _ in range(steps):
        eng.step(1); eng.face_bc()
    fluid = z["solid"] == 0
    assert np.array_equal(H.from_dev_scalar(eng.rho)[fluid], z["rho"][fluid])
    assert np.array_equal(H.from_dev_vec(eng.u)[fluid], z["u"][fluid])
    assert np.array_equal(H.from_dev_pop(eng.export_f())[:, fluid], z["f_out"][:, fluid])


def test_neighbour_kernels_reproduce_the_reference_run():
    """Pressure-gradient drive (force / mixed mode), Forchheimer resistance, particle coupling (cell data, drag, Reynolds
    numbers bit-exact; C_D within powf's 2 ulp; scattered reaction within the atomics' ordering noise), under-relaxation
    and the particle integrator with its error counters."""
    # (bit-exact: interpolated fluid velocity, Reynolds number, cell indices, integrator; drag within powf's noise)
    import torch
    from pour_over_coffee_lbm_b200.engine import ParticleState, particles_couple, particles_advance
    z = np.load(os.path.join(GOLD, "reference_run_neighbours.npz"))
    n = int(z["n"])
    eng = _engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, force=True, phase=True, porous=True,
                  config=_cfg(n, R.RefConfig().GRAVITY_LU))
    eng.build_v60_geometry()
    assert np.array_equal(H.from_dev_scalar(eng.solid), z["solid"])
    eng.rho.copy_(_torch(H.to_dev_scalar(z["rho"]))); eng.u.copy_(_torch(H.to_dev_vec(z["u"])))
    eng.clear_body_force(); eng.add_pressure_gradient_force(0.12, 1.0)
    assert np.array_equal(H.from_dev_vec(eng.body_force), z["bf_force_drive"])
    eng.clear_body_force(); eng.add_pressure_gradient_force(0.12, 0.5)
    assert np.array_equal(H.from_dev_vec(eng.body_force), z["bf_mixed_drive"])
    eng.add_forchheimer_force()
    assert np.array_equal(H.from_dev_vec(eng.body_force), z["bf_mixed_plus_forchheimer"])

    P = z["p_pos"].shape[0]
    act = z["p_active"] != 0
    ps = ParticleState(P, eng.device)
    ps.pos.copy_(_torch(z["p_pos"].T)); ps.vel.copy_(_torch(z["p_vel"].T)); ps.radius.copy_(_torch(z["p_radius"]))
    ps.mass.copy_(_torch(z["p_mass"])); ps.active.copy_(_torch(z["p_active"])); ps.drag_old.copy_(_torch(z["p_drag_old_in"].T))
    react = torch.zeros_like(eng.u)
    particles_couple(eng, ps, react, relax=0.8)
    # C_D goes through powf (device libm vs NumPy: <= 2 ulp), and the drag inherits it
    assert np.allclose(ps.drag_new.cpu().numpy().T[act], z["p_drag_new"][act], rtol=1e-6, atol=0)
    assert np.array_equal(ps.u_fluid.cpu().numpy().T[act], z["p_u_fluid"][act])
    assert np.array_equal(ps.reynolds.cpu().numpy()[act], z["p_reynolds"][act])
    assert np.allclose(ps.cd.cpu().numpy()[act], z["p_cd"][act], rtol=3e-7, atol=0)
    assert np.allclose(H.from_dev_vec(react), z["p_reaction"], rtol=1e-5, atol=1e-12)
    assert np.allclose(ps.drag.cpu().numpy().T[act], z["p_drag"][act], rtol=1e-6, atol=1e-16)
    assert np.allclose(ps.drag_old.cpu().numpy().T[act], z["p_drag_old_out"][act], rtol=1e-6, atol=1e-16)
    assert np.array_equal(ps.cell.cpu().numpy().T[act], np.stack(R.particle_cell_and_weights(R.RefConfig(NX=n, NY=n, NZ=n), z["p_pos"])[:3], 1)[act])

    cx, cy, bz, br, tr = [float(v) for v in z["bounds"]]
    force = _torch(z["p_force_in"].T).contiguous()
    counters = torch.zeros(2, dtype=torch.int32, device="cuda")
    ps.pos.copy_(_torch(z["p_pos"].T)); ps.vel.copy_(_torch(z["p_vel"].T)); ps.active.copy_(_torch(z["p_active"]))
    for t, dt in enumerate(z["adv_dts"]):
        particles_advance(eng, ps, float(dt), cx, cy, bz, br, tr, force=force, counters=counters)
        a = z[f"adv{t}_active"] == 1
        assert np.array_equal(ps.active.cpu().numpy(), z[f"adv{t}_active"])
        assert np.array_equal(ps.pos.cpu().numpy().T[a], z[f"adv{t}_pos"][a])
        assert np.array_equal(ps.vel.cpu().numpy().T[a], z[f"adv{t}_vel"][a], equal_nan=True)
    assert counters.cpu().tolist() == [int(v) for v in z["adv_counters"]]


def test_density_drive_reproduces_the_reference_run():
    """lbm_density_drive through the PressureGradientDrive facade (apply() in density mode, three calls) against the recorded run of
    the reference's PressureGradientDrive method A: target profile and rho after every call, bit for bit."""
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    from pour_over_coffee_lbm_b200.physics import PressureGradientDrive
    z = np.load(os.path.join(GOLD, "reference_run_density_drive.npz"))
    n = int(z["n"])
    s = LBMSolver(config=_cfg(n, R.RefConfig().GRAVITY_LU)); s.init_fields()
    s.engine.build_v60_geometry()
    assert np.array_equal(H.from_dev_scalar(s.engine.solid), z["solid"])
    pg = PressureGradientDrive(s)
    assert np.array_equal(pg.target_density.to_numpy(), z["target"])
    s.rho.from_numpy(z["rho"])
    pg.activate_density_drive(True)
    for t in range(3):
        pg.apply(t)
        assert np.array_equal(s.rho.to_numpy(), z[f"rho_after_{t + 1}"])
