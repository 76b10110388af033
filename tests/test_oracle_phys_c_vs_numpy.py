"""oracle/phys_cpu.c (the C twin of d3q19_ref.step_physical, used to check the CUDA kernels at BASELINE's grid sizes) against the
NumPy function that defines compat = physical: bit for bit for every feature combination, periodic and walled boxes, both memory
layouts, and against the committed golden vector."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R
from oracle import ref_cpu as RC

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _state(n, seed, walls):
    rng = np.random.default_rng(seed)
    u0 = H.smooth_velocity(n, 0.03, seed); rho0 = H.smooth_density(n, 0.01, seed)
    g = R.init_equilibrium_phys(rho0, u0)
    g *= (1.0 + 1e-3 * rng.standard_normal(g.shape)).astype(np.float32)
    solid = zone = None
    if walls:
        st = H.reference_v60_state(n, seed=seed, gravity=1e-5, body=1e-5, phase_mode="split")
        solid, zone = st.solid, st.filter_zone
    bf = (2e-5 * rng.standard_normal((n, n, n, 3))).astype(np.float32)
    phase = rng.uniform(0, 1, (n, n, n)).astype(np.float32)
    les_mask = (rng.uniform(0, 1, (n, n, n)) > 0.2).astype(np.int32)
    return g, solid, zone, bf, phase, les_mask


@pytest.mark.parametrize("walls", [False, True])
@pytest.mark.parametrize("feat", ["bgk", "les", "forced", "forced+les+porous", "phase_only"])
def test_c_twin_of_step_physical_is_bit_exact(walls, feat):
    n = 20
    if "porous" in feat and not walls:
        pytest.skip("the filter zone lives in the flag field")
    g, solid, zone, bf, phase, les_mask = _state(n, 3, walls)
    p = R.PhysParams(nx=n, ny=n, nz=n, tau_water=0.56, tau_air=0.8, gravity_lu=2e-5 if feat != "bgk" else 0.0,
                     periodic=(not walls, not walls, True) if walls else (True, True, True),
                     use_force="forced" in feat, use_phase=feat != "bgk" and feat != "les", les="les" in feat, porous="porous" in feat,
                     porous_darcy=0.2, porous_forch=0.5)
    kw = dict(solid=solid, body_force=bf if p.use_force else None, phase=phase if p.use_phase else None,
              filter_zone=zone if p.porous else None, les_mask=les_mask if p.les else None)
    a, b = g, g
    for _ in range(3):
        a, ra, ua = R.step_physical(a, p, **kw)
        b, rb, ub = RC.phys_step(b, p, **kw)
        fluid = np.ones((n, n, n), bool) if solid is None else solid == 0
        assert np.array_equal(a[:, fluid], b[:, fluid]) and np.array_equal(ra, rb) and np.array_equal(ua, ub)
    # the same step on device-ordered arrays ([q][z][y][x], vectors [3][z][y][x])
    dev = lambda x: None if x is None else H.to_dev_scalar(x)
    c, rc, uc = RC.phys_step(H.to_dev_pop(g), p, solid=dev(solid), body_force=None if not p.use_force else H.to_dev_vec(bf),
                             phase=dev(kw["phase"]), filter_zone=dev(kw["filter_zone"]), les_mask=dev(kw["les_mask"]), layout="device")
    a1, r1, u1 = R.step_physical(g, p, **kw)
    assert np.array_equal(np.transpose(c, (0, 3, 2, 1))[:, fluid], a1[:, fluid]) and np.array_equal(np.transpose(rc, (2, 1, 0)), r1)
    assert np.array_equal(np.transpose(uc, (3, 2, 1, 0)), u1)
    assert RC.count_mismatch(H.to_dev_pop(a1), c, dev(solid)) == 0


def test_c_twin_reproduces_the_golden_vector():
    """tests/golden/step_physical_24.npz (periodic 24^3, LES, made by tests/golden/make_golden.py from the NumPy oracle)."""
    z = np.load(os.path.join(GOLD, "step_physical_24.npz"))
    n = int(z["n"])
    p = R.PhysParams(nx=n, ny=n, nz=n, tau_water=float(z["tau"]), les=True)
    g = R.init_equilibrium_phys(z["rho0"], z["u0"])
    for _ in range(int(z["steps"])):
        g, rho, u = RC.phys_step(g, p)
    assert np.array_equal(g, z["g"]) and np.array_equal(rho, z["rho"]) and np.array_equal(u, z["u"])
