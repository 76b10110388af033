"""The particle kernels' SOURCE (csrc/lbm_particles.cu), compiled by g++ and executed thread by thread on the CPU
(tests/emu/emu_particles.cpp), against the recorded runs of the reference's CoffeeParticleSystem -- the fixtures the GPU tests
use.  See tests/test_producers_emulated.py for why.  Test infrastructure only."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


class Particles(C.Structure):          # include/lbm_b200.h: lbm_particles
    _fields_ = [(n, C.c_void_p) for n in ("pos", "vel", "radius", "mass", "active", "drag_new", "drag_old", "drag", "u_fluid", "reynolds", "cd",
                                          "cell")] + [("n", C.c_int)]


class Bounds(C.Structure):             # include/lbm_b200.h: lbm_particle_bounds
    _fields_ = [(n, C.c_float) for n in ("center_x", "center_y", "bottom_z", "bottom_radius_lu", "top_radius_lu", "cup_height_lu", "max_coordinate",
                                         "nz_minus_5")]


@pytest.fixture(scope="module")
def emu():
    return H.build_emu("emu_particles", ['lbm_particles.cu', 'lbm_common.cuh'])


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class State:
    def __init__(self, pos, vel, radius, mass, active):
        n = pos.shape[0]
        t = lambda a: np.ascontiguousarray(a.T.astype(np.float32))
        self.pos, self.vel = t(pos), t(vel)
        self.radius, self.mass = radius.astype(np.float32).copy(), mass.astype(np.float32).copy()
        self.active = active.astype(np.int32).copy()
        z3 = lambda: np.zeros((3, n), np.float32)
        self.drag_new, self.drag_old, self.drag, self.u_fluid = z3(), z3(), z3(), z3()
        self.reynolds, self.cd = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.cell = np.zeros((3, n), np.int32)
        self.n = n

    def struct(self):
        return Particles(*[_p(getattr(self, k)) for k in ("pos", "vel", "radius", "mass", "active", "drag_new", "drag_old", "drag", "u_fluid",
                                                            "reynolds", "cd", "cell")], self.n)


def test_emulated_coupling_under_relaxation_and_integrator_reproduce_the_reference_run(emu):
    z = np.load(os.path.join(GOLD, "reference_run_neighbours.npz"))
    n = int(z["n"]); cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    act = z["p_active"] != 0
    st = State(z["p_pos"], z["p_vel"], z["p_radius"], z["p_mass"], z["p_active"])
    st.drag_old[:] = z["p_drag_old_in"].T
    u = H.to_dev_vec(z["u"]); react = np.zeros_like(u)
    s = st.struct()
    rho_w = np.float32(cfg.WATER_DENSITY_90C); mu_w = np.float32(cfg.WATER_VISCOSITY_90C * cfg.WATER_DENSITY_90C)
    emu.emu_particles_couple(C.c_int(n), C.c_int(n), C.c_int(n), _p(u), _p(react), C.byref(s), C.c_float(rho_w), C.c_float(mu_w), C.c_float(0.8))
    assert np.array_equal(st.u_fluid.T[act], z["p_u_fluid"][act]) and np.array_equal(st.reynolds[act], z["p_reynolds"][act])
    assert np.allclose(st.cd[act], z["p_cd"][act], rtol=3e-7, atol=0)                    # powf: glibc vs NumPy
    assert np.allclose(st.drag_new.T[act], z["p_drag_new"][act], rtol=1e-6, atol=0)
    assert np.allclose(np.transpose(react, (3, 2, 1, 0)), z["p_reaction"], rtol=1e-5, atol=1e-12)        # scatter order
    assert np.allclose(st.drag.T[act], z["p_drag"][act], rtol=1e-6, atol=1e-16) and np.allclose(st.drag_old.T[act], z["p_drag_old_out"][act], rtol=1e-6, atol=1e-16)
    assert np.array_equal(st.cell.T[act], np.stack(R.particle_cell_and_weights(cfg, z["p_pos"])[:3], 1)[act])
    # stand-alone under-relaxation on the recorded drag_new
    st2 = State(z["p_pos"], z["p_vel"], z["p_radius"], z["p_mass"], z["p_active"])
    st2.drag_new[:] = z["p_drag_new"].T; st2.drag_old[:] = z["p_drag_old_in"].T
    s2 = st2.struct()
    emu.emu_particles_under_relax(C.byref(s2), C.c_float(0.8))
    assert np.array_equal(st2.drag.T[act], z["p_drag"][act]) and np.array_equal(st2.drag_old.T[act], z["p_drag_old_out"][act])
    # integrator: three calls with the recorded forces and dt's
    st3 = State(z["p_pos"], z["p_vel"], z["p_radius"], z["p_mass"], z["p_active"])
    s3 = st3.struct()
    force = np.ascontiguousarray(z["p_force_in"].T)
    cx, cy, bz, br, tr = [float(v) for v in z["bounds"]]
    cup = np.float32(cfg.CUP_HEIGHT / cfg.SCALE_LENGTH)
    b = Bounds(cx, cy, bz, br, tr, float(cup), float(max(n, n, n)), float(n - 5))
    counters = np.zeros(2, np.int32)
    for t, dt in enumerate(z["adv_dts"]):
        emu.emu_particles_advance(C.byref(s3), _p(force), C.byref(b), C.c_float(float(dt)), _p(counters))
        a = st3.active == 1
        assert np.array_equal(st3.active, z[f"adv{t}_active"])
        assert np.array_equal(st3.pos.T[a], z[f"adv{t}_pos"][a]) and np.array_equal(st3.vel.T[a], z[f"adv{t}_vel"][a], equal_nan=True)
    assert counters.tolist() == [int(v) for v in z["adv_counters"]]


def test_emulated_fluid_forces_reproduce_the_reference_run(emu):
    z = np.load(os.path.join(GOLD, "reference_run_filter_particles.npz"))
    n = int(z["n"])
    st = State(z["ff_pos"], z["ff_vel"], z["ff_radius"], z["ff_mass"], z["ff_active"])
    s = st.struct()
    force = np.ascontiguousarray(z["ff_force_in"].T)
    counters = np.zeros(2, np.int32)
    u = H.to_dev_vec(z["ff_u"])
    emu.emu_particles_fluid_forces(C.c_int(n), C.c_int(n), C.c_int(n), _p(u), C.byref(s), _p(force), C.c_double(float(z["water_density"])),
                                   C.c_double(float(z["water_viscosity"])), C.c_double(float(z["particle_gravity"])), _p(counters))
    assert np.array_equal(force.T, z["ff_force"]) and np.array_equal(st.vel.T, z["ff_vel_out"]) and np.array_equal(st.active, z["ff_active_out"])
    assert int(counters[0]) == int(z["ff_errors"])
