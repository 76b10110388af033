"""The collision operator of compat = physical -- collide_phys<float, ...> in csrc/lbm_phys.cuh, the arithmetic contract of
the headline kernel -- compiled by g++ and applied after a periodic pull (tests/emu/emu_collision.cpp), against
oracle/d3q19_ref.py:step_physical.  Bit for bit, several steps, every feature combination of the operator.
(The packed f32x2 instantiation the GPU runs is inline PTX; lbm_selftest_math() proves it equal to this scalar one on the
device.)  Test infrastructure only."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    return H.build_emu("emu_collision", ['lbm_phys.cuh', 'lbm_common.cuh'])


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@pytest.mark.parametrize("case", ["bgk", "les", "forced", "forced_les_porous", "mrt_bgk", "mrt_forced_les_porous"])
def test_emulated_collision_operator_matches_the_oracle(emu, case):
    nx, ny, nz, steps = 12, 10, 8, 4
    rng = np.random.default_rng(9)
    u0 = H.smooth_velocity(nx, 0.05, 21, nz=nz, ny=ny); rho0 = H.smooth_density(nx, 0.02, 21, nz=nz, ny=ny)
    mrt = case.startswith("mrt_")          # the two-rate MRT collision (lbm_params.mrt_magic = 3/16)
    case = case[4:] if mrt else case
    les = case in ("les", "forced_les_porous"); forced = case.startswith("forced"); porous = case == "forced_les_porous"
    p = R.PhysParams(nx=nx, ny=ny, nz=nz, tau_water=0.53, tau_air=0.8, gravity_lu=1e-4 if forced else 0.0, use_force=forced, use_phase=forced,
                     les=les, porous=porous, porous_darcy=0.37 if porous else 0.0, porous_forch=0.9 if porous else 0.0, mrt_magic=0.1875 if mrt else 0.0)
    sh = (nx, ny, nz)
    bf = (1e-4 * rng.standard_normal(sh + (3,))).astype(np.float32) if forced else None
    phase = rng.uniform(0, 1, sh).astype(np.float32) if forced else None
    zone = (rng.random(sh) < 0.2).astype(np.int32) if porous else None
    les_mask = (rng.random(sh) < 0.8).astype(np.int32) if les else None
    g = R.init_equilibrium_phys(rho0, u0)
    flags = None
    if les or porous:
        fl = np.zeros(sh, np.uint8)
        fl |= (4 * (les_mask if les_mask is not None else np.ones(sh, np.int32))).astype(np.uint8)       # LBM_FLAG_LES
        if porous: fl |= (2 * zone).astype(np.uint8)                                                         # LBM_FLAG_FILTER
        flags = H.to_dev_scalar(fl)
    d_g = H.to_dev_pop(g)
    d_bf = H.to_dev_vec(bf) if forced else None
    d_ph = H.to_dev_scalar(phase) if forced else None
    f32 = lambda v: C.c_float(float(v))
    for _ in range(steps):
        g, rho, u = R.step_physical(g, p, solid=None, body_force=bf, phase=phase, filter_zone=zone, les_mask=les_mask)
        out = np.empty_like(d_g); d_rho = np.empty((nz, ny, nx), np.float32); d_u = np.empty((3, nz, ny, nx), np.float32)
        emu.emu_collide_periodic(C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(d_g), _p(out), _p(d_rho), _p(d_u), _p(d_bf), _p(d_ph), _p(flags),
                                 C.c_int(int(les)), C.c_int(int(porous)), f32(p.tau_water), f32(p.tau_air), f32(p.gravity_lu), f32(p.cs_smag),
                                 f32(p.tau_min), f32(p.tau_max), f32(p.porous_darcy), f32(p.porous_forch), f32(p.mrt_magic))
        d_g = out
        assert np.array_equal(np.transpose(d_g, (0, 3, 2, 1)), g)
        assert np.array_equal(np.transpose(d_rho, (2, 1, 0)), rho) and np.array_equal(np.transpose(d_u, (3, 2, 1, 0)), u)
