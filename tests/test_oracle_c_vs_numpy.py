"""The C restatement (oracle/ref_cpu.c, the timed CPU baseline) must equal the NumPy restatement bit for bit."""
import numpy as np

import helpers as H
from oracle import d3q19_ref as R
from oracle import ref_cpu as RC


def _compare(st, cs):
    for name in ("f", "f_new", "rho", "u", "nu_sgs"):
        assert np.array_equal(getattr(st, name), getattr(cs, name), equal_nan=True), name


def test_c_equals_numpy_v60_air_phase():
    st = H.reference_v60_state(24, seed=1, gravity=2e-5, body=1e-5, phase_mode="none")
    st.phase[:] = np.random.default_rng(9).uniform(0, 0.5, st.phase.shape).astype(np.float32)
    cs = RC.CState(st)
    for _ in range(12):
        R.step(st)
    cs.step(12)
    _compare(st, cs)


def test_c_equals_numpy_water_phase_les_default_gravity():
    st = H.reference_v60_state(20, seed=3, gravity=R.RefConfig().GRAVITY_LU, body=1e-4, phase_mode="split")
    cs = RC.CState(st)
    for _ in range(8):
        R.step(st)
    cs.step(8)
    assert st.nu_sgs.max() > 0
    _compare(st, cs)


def test_c_equals_numpy_open_box_no_geometry():
    cfg = R.RefConfig(NX=12, NY=10, NZ=14, GRAVITY_LU=1e-4)
    st = R.init_fields(cfg)
    u0 = H.smooth_velocity(12, 0.02, 4, nz=14, ny=10)
    for q in range(R.Q):
        st.f[q] = R.equilibrium_ref(st.rho, u0[..., 0], u0[..., 1], u0[..., 2], q, "config"); st.f_new[q] = st.f[q]
    cs = RC.CState(st)
    for _ in range(6):
        R.step(st)
    cs.step(6)
    _compare(st, cs)


def test_c_fused_pointer_swap_variant_equals_copy_swap():
    st = H.reference_v60_state(16, seed=7, gravity=2e-5, body=1e-5, phase_mode="none")
    a, b = RC.CState(st), RC.CState(st)
    a.step(5); b.step(5, fused=True)
    _compare(a, b)


def test_c_geometry_equals_numpy():
    for n in (32, 64):
        cfg = R.RefConfig(NX=n, NY=n, NZ=n)
        assert np.array_equal(R.v60_solid(cfg), RC.v60_solid(cfg))
    assert RC.num_threads() >= 1
