"""Taylor-Green decay rate (within 0.5 % of 4 nu k^2) and the ghost-plane slab path on one GPU."""
import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R

pytestmark = pytest.mark.gpu


def _engine(*a, **k):
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine
    return D3Q19Engine(*a, **k)


def _tgv2d(n, u0):
    import torch
    k = 2 * np.pi / n
    x = torch.arange(n, dtype=torch.float64, device="cuda") * k
    X = x[None, None, :]; Y = x[None, :, None]
    one = torch.ones((n, 1, 1), dtype=torch.float64, device="cuda")
    ux = u0 * torch.sin(X) * torch.cos(Y) * one
    uy = -u0 * torch.cos(X) * torch.sin(Y) * one
    rho = (1.0 - (3.0 * u0 * u0 / 4.0) * (torch.cos(2 * X) + torch.cos(2 * Y))) * one
    return rho.float().contiguous(), torch.stack([ux, uy, torch.zeros_like(ux)]).float().contiguous()


@pytest.mark.parametrize("tau,mrt_magic", [(0.53, 0.0), (0.8, 0.0), (0.53, 0.1875), (0.8, 0.25)])
def test_taylor_green_decay_rate_256(tau, mrt_magic):
    """z-invariant Taylor-Green vortex on periodic 256^3 (exact Navier-Stokes solution): kinetic energy decays as
    exp(-4 nu k^2 t), nu = (tau - 1/2)/3.  Fit ln E over steps 200..1000.
    u0 = 0.01 is the reference's own Mach number (config MACH_NUMBER = 0.0173 = u0*sqrt(3)).  The second-order
    equilibrium carries an O(Ma^2) viscosity error: at u0 = 0.04 and tau = 0.53 the measured rate is 3.9 % high in
    f32 AND in an f64 run of the oracle, 0.98 % at 0.02, 0.25 % at 0.01.
    mrt_magic > 0: the two-rate MRT collision (lbm_params.mrt_magic) -- the even moments still relax at 1 / tau, so the viscosity,
    and with it the decay rate, is the same analytic one."""
    import torch
    n, u0 = 256, 0.01
    eng = _engine(n, n, n, compat="physical", tau=tau, mrt_magic=mrt_magic)          # default build = strict
    rho0, uinit = _tgv2d(n, u0)
    eng.init_equilibrium(rho=rho0, u=uinit)
    steps, every = 1000, 50
    ts, es = [], []
    for s in range(0, steps, every):
        eng.step(every, write_macro_every=every)
        e = float((0.5 * eng.rho.double() * (eng.u.double() ** 2).sum(0)).sum())
        # rho,u written by step k are the moments of the state entering step k (t = s + every - 1)
        ts.append(s + every - 1); es.append(e)
    ts, es = np.array(ts, float), np.array(es, float)
    sel = ts >= 200
    slope = np.polyfit(ts[sel], np.log(es[sel]), 1)[0]
    nu = (tau - 0.5) / 3.0
    k = 2 * np.pi / n
    expected = -4.0 * nu * k * k
    assert abs(slope / expected - 1.0) <= 5e-3, (slope, expected)
    # uz stays at rounding level and the flow stays exactly z-invariant
    assert float(eng.u[2].abs().max()) < 1e-4 * u0
    assert float((eng.u[0][0] - eng.u[0][n // 2]).abs().max()) == 0.0


@pytest.mark.parametrize("vec", [1, 4])
def test_ghost_plane_slab_equals_wrapped_single_slab(vec):
    """zghost=1 with the periodic self-exchange (single-GPU 'virtual slab' mode) must reproduce the zghost=0
    in-kernel wrap bit for bit: validates the halo population set {5,11,12,15,16 | 6,13,14,17,18}."""
    import torch
    n, steps = 32, 25
    u0 = H.smooth_velocity(n, 0.04, 21); rho0 = H.smooth_density(n, 0.01, 21)
    ru = torch.from_numpy(H.to_dev_scalar(rho0)).cuda(); uu = torch.from_numpy(H.to_dev_vec(u0)).cuda()
    a = _engine(n, n, n, compat="physical", les=True, strict=True, vec=vec, tau=0.6)
    a.init_equilibrium(rho=ru, u=uu)
    a.step(steps)
    b = _engine(n, n, n, compat="physical", les=True, strict=True, vec=vec, tau=0.6, zghost=1, z0=0, nz_global=n)
    pad = lambda t: torch.nn.functional.pad(t, (0, 0, 0, 0, 1, 1))
    b.init_equilibrium(rho=pad(ru), u=pad(uu))
    b.step(steps)
    assert torch.equal(a.populations, b.populations[:, 1:-1])
    assert torch.equal(a.rho, b.rho[1:-1]) and torch.equal(a.u, b.u[:, 1:-1])


@pytest.mark.parametrize("tma", [False, True])
@pytest.mark.parametrize("periodic", [(True, True, True), (False, False, True)])
def test_ghost_plane_slab_with_walls_equals_wrapped_single_slab(periodic, tma, monkeypatch):
    """Same as above behind walls (compat = physical): obstacles straddle the slab interface, so bounce-back slots
    that live in the ghost planes are overwritten by every exchange and must be rebuilt (refresh after the halo).
    tma: the opt-in TMA-staged kernel (x/y open; ghost planes = tensor planes 0, nz+1)."""
    import torch
    if tma:
        if periodic[0]:
            pytest.skip("TMA-staged kernel not eligible")
        monkeypatch.setenv("LBM_TMA", "1")
    n, steps = 32, 25
    rng = np.random.default_rng(12)
    solid = (rng.random((n, n, n)) < 0.1).astype(np.uint8)
    solid[10:14, 10:14, 0:2] = 1; solid[20:24, 5:9, n - 2:] = 1
    sd = torch.from_numpy(H.to_dev_scalar(solid)).cuda()
    u0 = H.smooth_velocity(n, 0.03, 22); rho0 = H.smooth_density(n, 0.01, 22)
    ru = torch.from_numpy(H.to_dev_scalar(rho0)).cuda(); uu = torch.from_numpy(H.to_dev_vec(u0)).cuda()
    a = _engine(n, n, n, compat="physical", walls=True, les=True, tau=0.6, periodic=periodic)
    a.solid.copy_(sd); a.pack_flags()
    a.init_equilibrium(rho=ru, u=uu)
    a.step(steps)
    b = _engine(n, n, n, compat="physical", walls=True, les=True, tau=0.6, zghost=1, z0=0, nz_global=n, periodic=periodic)
    b.solid[1:-1].copy_(sd); b.solid[0].copy_(sd[-1]); b.solid[-1].copy_(sd[0]); b.pack_flags()
    pad = lambda t: torch.nn.functional.pad(t, (0, 0, 0, 0, 1, 1))
    b.init_equilibrium(rho=pad(ru), u=pad(uu))
    b.step(steps)
    fluid = sd == 0
    assert torch.equal(a.populations[:, fluid], b.populations[:, 1:-1][:, fluid])
    assert torch.equal(a.rho[fluid], b.rho[1:-1][fluid]) and torch.equal(a.u[:, fluid], b.u[:, 1:-1][:, fluid])


def test_strict_and_fast_builds_are_distinct_kernels():
    """compat = reference is built twice (-fmad=false / FMA contraction).  Guards against the two builds being
    merged at link time: contraction must change some low bits.  (compat = physical has one build.)"""
    import torch
    n = 32
    u0 = H.smooth_velocity(n, 0.05, 3); rho0 = H.smooth_density(n, 0.02, 3)
    out = []
    for strict in (True, False):
        e = _engine(n, n, n, compat="reference", strict=strict, tau=0.8)
        e.init_equilibrium(rho=torch.from_numpy(H.to_dev_scalar(rho0)).cuda(), u=torch.from_numpy(H.to_dev_vec(u0)).cuda())
        e.step(20)
        out.append(e.populations.clone())
    assert not torch.equal(out[0], out[1])
    assert float((out[0] - out[1]).abs().max()) < 1e-6


def test_physical_ignores_strict_flag():
    import torch
    n = 32
    u0 = H.smooth_velocity(n, 0.05, 3); rho0 = H.smooth_density(n, 0.02, 3)
    out = []
    for strict in (True, False):
        e = _engine(n, n, n, compat="physical", les=True, strict=strict, tau=0.53)
        e.init_equilibrium(rho=torch.from_numpy(H.to_dev_scalar(rho0)).cuda(), u=torch.from_numpy(H.to_dev_vec(u0)).cuda())
        e.step(20)
        out.append(e.populations.clone())
    assert torch.equal(out[0], out[1])


@pytest.mark.parametrize("compat", ["physical", "reference"])
def test_restart_checkpoint_is_bit_exact(compat, tmp_path):
    """save_checkpoint after 7 steps, load into a fresh engine, 9 more steps: identical to the uninterrupted run
    (V60 mask with every feature; the reference mode carries both u buffers of its lagged LES)."""
    import torch
    from pour_over_coffee_lbm_b200.config import LBMConfig
    n = 32
    cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
    kw = dict(porous_darcy=0.37, porous_forch=0.9) if compat == "physical" else {}

    def make():
        e = _engine(n, n, n, compat=compat, periodic=(False, False, False), walls=True, force=True, phase=True, les=True, porous=True,
                    config=cfg, gravity_lu=1e-5, **kw)
        return e
    a = make(); a.build_v60_geometry()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    a.phase.copy_((torch.rand(a.phase.shape, device="cuda", generator=g) < 0.5).float() * (0.4 if compat == "reference" else 1.0))
    a.body_force.copy_(1e-5 * torch.randn(a.body_force.shape, device="cuda", generator=g))
    a.init_equilibrium(rho=torch.ones(a.rho.shape, device="cuda"), u=1e-2 * torch.randn(a.u.shape, device="cuda", generator=g))
    a.step(7)
    path = str(tmp_path / "slab0.pt")
    a.save_checkpoint(path)
    a.step(9)
    b = make(); b.load_checkpoint(path)
    assert b.steps_done == 7
    b.step(9)
    fluid = a.solid == 0
    assert torch.equal(a.populations[:, fluid], b.populations[:, fluid])
    assert torch.equal(a.rho[fluid], b.rho[fluid]) and torch.equal(a.u[:, fluid], b.u[:, fluid])
    c = _engine(n, n, n, compat=compat, walls=True)
    with pytest.raises(ValueError):
        c.load_checkpoint(path)

