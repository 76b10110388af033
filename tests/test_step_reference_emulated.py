"""The legacy-compatible step kernel's SOURCE -- step_cells<LBM_COMPAT_REFERENCE, MODE_BULK, ..., VEC = 1> of
csrc/lbm_step_kernel.cuh: pull, halfway bounce-back and open-face inflow from the neighbour masks, FD-LES on the lagged u,
moments, clamped Guo term, BGK with the reference's equilibrium table, filter damping, write-back -- compiled by g++ and
run cell by cell on the CPU (tests/emu/emu_step_reference.cpp) against the recorded runs of the reference's own
LBMSolver.step(), the 1000-step recording included.  Bit for bit.

The flag byte, the neighbour masks and the f <-> g conversion are built here in NumPy from their documented meaning
(include/lbm_b200.h; on the device lbm_pack_flags / lbm_import_f / lbm_export_f produce them -- GPU-tested); everything
between the loads and the stores is the product's statement sequence.  Test infrastructure only.
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "reference_run_step_*.npz"))) + sorted(glob.glob(os.path.join(GOLD, "reference_run_long_air_*.npz")))


# three host builds of the same kernel source: the one-cell step as the device runs it (collide_reference), the one-cell step with the
# collision routed through the V-generic template (collide_reference_t<float>), and the two-cells-per-thread step with the packed
# collision (collide_reference_t<P2> on host stand-ins of the f32x2 primitives) -- opt-in on the device (vec = 2)
DEPS = ['lbm_step_kernel.cuh', 'lbm_phys.cuh', 'lbm_common.cuh', '../../tests/emu/emu_step_reference.cpp']


@pytest.fixture(scope="module", params=["emu_step_reference", "emu_step_reference_generic", "emu_step_reference_vec2"],
                ids=["one_cell", "one_cell_generic_collision", "two_cells_packed"])
def emu(request):
    return H.build_emu(request.param, DEPS)


@pytest.fixture(scope="module")
def emu_dense():
    return H.build_emu("emu_step_reference", DEPS)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _shift(a, q, sign):
    """a[x + sign * e_q] with out-of-box entries marked by the returned mask (logical [i,j,k] arrays)."""
    e = (int(R.CX[q]) * sign, int(R.CY[q]) * sign, int(R.CZ[q]) * sign)
    out = np.zeros_like(a); inside = np.zeros(a.shape[-3:], bool)
    n = a.shape[-3:]
    src = [slice(max(0, e[d]), n[d] + min(0, e[d])) for d in range(3)]
    dst = [slice(max(0, -e[d]), n[d] + min(0, -e[d])) for d in range(3)]
    out[(Ellipsis,) + tuple(dst)] = a[(Ellipsis,) + tuple(src)]
    inside[tuple(dst)] = True
    return out, inside


def neighbour_masks(solid):
    """bit q of the low word: the source cell x - e_q of population q is solid (bounce-back); bit q of the high word: it lies
    outside the (non-periodic) box (stale inflow w_q).  NEAR = any bit set."""
    lo = np.zeros(solid.shape, np.uint64); hi = np.zeros(solid.shape, np.uint64)
    for q in range(1, R.Q):
        s, inside = _shift(solid, q, -1)
        lo |= np.where(inside & (s != 0), np.uint64(1 << q), np.uint64(0))
        hi |= np.where(~inside, np.uint64(1 << q), np.uint64(0))
    return lo | (hi << np.uint64(32))


def f_to_g(f, solid):
    """post-collision populations whose pull reproduces the reference's post-stream f (what lbm_import_f does)."""
    g = np.zeros_like(f); g[0] = f[0]
    for q in range(1, R.Q):
        tgt_f, inside = _shift(f[q], q, +1)
        tgt_solid, _ = _shift(solid, q, +1)
        g[q] = np.where(inside & (tgt_solid == 0), tgt_f, np.where(inside, f[int(R.OPP[q])], np.float32(0)))
    return g


def g_to_f(g, solid):
    """the reference's f view of the device state (what lbm_export_f does): pull with bounce-back and the w_q inflow rule."""
    f = np.zeros_like(g); f[0] = g[0]
    for q in range(1, R.Q):
        src_g, inside = _shift(g[q], q, -1)
        src_solid, _ = _shift(solid, q, -1)
        f[q] = np.where(inside & (src_solid == 0), src_g, np.where(inside, g[int(R.OPP[q])], R.W[q]))
    return f


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[14:-4] for p in FILES])
def test_emulated_reference_step_kernel_reproduces_the_reference_run(emu, path):
    from pour_over_coffee_lbm_b200.config import LBMConfig
    z = np.load(path)
    n, steps, gravity = int(z["n"]), int(z["steps"]), float(z["gravity"])
    c = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=gravity)
    cfg = LBMConfig(NX=n, NY=n, NZ=n, TAU_FLUID=c.TAU_WATER, TAU_AIR=c.TAU_AIR, GRAVITY_LU=gravity)
    k_lu, beta_lu = cfg.forchheimer_parameters(); c_darcy, c_forch = cfg.filter_constants()
    solid, zone, les_mask = z["solid"], z["filter_zone"], z["les_mask"]
    nbr = neighbour_masks(solid)
    flags = (solid.astype(np.uint8) * 1) | ((zone != 0).astype(np.uint8) * 2) | ((les_mask != 0).astype(np.uint8) * 4) | ((nbr != 0).astype(np.uint8) * 8)
    d_flags, d_nbr = H.to_dev_scalar(flags), H.to_dev_scalar(nbr)
    bufs = [H.to_dev_pop(f_to_g(z["f"], solid)), None]
    bufs[1] = bufs[0].copy()
    d_force, d_phase = H.to_dev_vec(z["body_force"]), H.to_dev_scalar(z["phase"])
    d_rho = np.ones((n, n, n), np.float32)
    u_bufs = [np.zeros((3, n, n, n), np.float32), np.zeros((3, n, n, n), np.float32)]
    d_blockage = np.zeros((n, n, n), np.float32)
    f32 = lambda v: C.c_float(float(v))
    cur = 0
    for _ in range(steps):
        emu.emu_step_reference(C.c_int(n), C.c_int(n), C.c_int(n), _p(bufs[cur]), _p(bufs[1 - cur]), _p(d_rho), _p(u_bufs[cur]), _p(u_bufs[1 - cur]),
                               _p(d_force), _p(d_phase), _p(d_blockage), _p(d_flags), _p(d_nbr), C.c_int(1), C.c_int(1), f32(cfg.TAU_WATER),
                               f32(cfg.TAU_AIR), f32(gravity), f32(cfg.LES_CS), f32(0.55), f32(1.90), f32(k_lu), f32(beta_lu), f32(c_darcy), f32(c_forch))
        cur = 1 - cur
    fluid = solid == 0
    rho = np.transpose(d_rho, (2, 1, 0)); u = np.transpose(u_bufs[cur], (3, 2, 1, 0))
    f_out = g_to_f(np.transpose(bufs[cur], (0, 3, 2, 1)), solid)
    assert np.array_equal(rho[fluid], z["rho"][fluid]) and np.array_equal(u[fluid], z["u"][fluid])
    assert np.array_equal(f_out[:, fluid], z["f_out"][:, fluid])


def test_emulated_reference_step_kernel_on_the_open_box_recording(emu):
    """No V60 mask, no filter system (the first 30 steps of main.py): every face is open -- the stale-inflow rule of the
    neighbour masks' high word -- and obstacles touch the faces.  u and f are compared everywhere; rho away from the faces
    (the boundary manager's face writes are a separate kernel, lbm_face_bc, GPU-tested)."""
    from pour_over_coffee_lbm_b200.config import LBMConfig
    z = np.load(os.path.join(GOLD, "reference_run_openbox.npz"))
    n, steps, gravity = int(z["n"]), int(z["steps"]), float(z["gravity"])
    c = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=gravity)
    cfg = LBMConfig(NX=n, NY=n, NZ=n, TAU_FLUID=c.TAU_WATER, TAU_AIR=c.TAU_AIR, GRAVITY_LU=gravity)
    solid, les_mask = z["solid"], z["les_mask"]
    nbr = neighbour_masks(solid)
    flags = (solid.astype(np.uint8) * 1) | ((les_mask != 0).astype(np.uint8) * 4) | ((nbr != 0).astype(np.uint8) * 8)
    assert (nbr >> np.uint64(32)).any()                                       # open faces present
    d_flags, d_nbr = H.to_dev_scalar(flags), H.to_dev_scalar(nbr)
    bufs = [H.to_dev_pop(f_to_g(z["f"], solid)), None]; bufs[1] = bufs[0].copy()
    d_force, d_phase = H.to_dev_vec(z["body_force"]), H.to_dev_scalar(z["phase"])
    d_rho = np.ones((n, n, n), np.float32)
    u_bufs = [np.zeros((3, n, n, n), np.float32), np.zeros((3, n, n, n), np.float32)]
    f32 = lambda v: C.c_float(float(v))
    cur = 0
    for _ in range(steps):
        emu.emu_step_reference(C.c_int(n), C.c_int(n), C.c_int(n), _p(bufs[cur]), _p(bufs[1 - cur]), _p(d_rho), _p(u_bufs[cur]), _p(u_bufs[1 - cur]),
                               _p(d_force), _p(d_phase), None, _p(d_flags), _p(d_nbr), C.c_int(1), C.c_int(0), f32(cfg.TAU_WATER), f32(cfg.TAU_AIR),
                               f32(gravity), f32(cfg.LES_CS), f32(0.55), f32(1.90), f32(1.0), f32(1.0), f32(0.0), f32(0.0))
        cur = 1 - cur
    fluid = solid == 0
    u = np.transpose(u_bufs[cur], (3, 2, 1, 0)); rho = np.transpose(d_rho, (2, 1, 0))
    f_out = g_to_f(np.transpose(bufs[cur], (0, 3, 2, 1)), solid)
    assert np.array_equal(u[fluid], z["u"][fluid]) and np.array_equal(f_out[:, fluid], z["f_out"][:, fluid])
    inner = fluid.copy(); inner[[0, -1]] = False; inner[:, [0, -1]] = False; inner[:, :, [0, -1]] = False
    assert np.array_equal(rho[inner], z["rho"][inner])


# ---- compat = physical, the headline configuration: step_cells<PHYSICAL, DENSE, ..., VEC = 1> ---------------------------
@pytest.mark.parametrize("les", [False, True])
def test_emulated_dense_physical_step_kernel_matches_the_oracle_and_its_golden(emu_dense, les):
    """pull with in-kernel periodic wrap + collide_phys<float> + write-back, 10 steps at 24^3 from the committed fixture's
    initial state: bit-exact against oracle.step_physical (and, with LES, against tests/golden/step_physical_24.npz)."""
    emu = emu_dense
    z = np.load(os.path.join(GOLD, "step_physical_24.npz"))
    n, steps = int(z["n"]), int(z["steps"])
    p = R.PhysParams(nx=n, ny=n, nz=n, tau_water=float(z["tau"]), les=les)
    g = R.init_equilibrium_phys(z["rho0"], z["u0"])
    b0 = H.to_dev_pop(g); b1 = np.empty_like(b0)
    rho = np.empty((n, n, n), np.float32); u = np.empty((3, n, n, n), np.float32)
    for _ in range(steps):
        g, rho_o, u_o = R.step_physical(g, p)
    f32 = lambda v: C.c_float(float(v))
    cur = emu.emu_step_physical_dense(C.c_int(n), C.c_int(n), C.c_int(n), C.c_int(steps), _p(b0), _p(b1), _p(rho), _p(u), C.c_int(int(les)),
                                      f32(p.tau_water), f32(p.cs_smag), f32(p.tau_min), f32(p.tau_max), C.c_int(1))
    got = np.transpose((b0, b1)[cur], (0, 3, 2, 1))
    assert np.array_equal(got, g) and np.array_equal(np.transpose(rho, (2, 1, 0)), rho_o) and np.array_equal(np.transpose(u, (3, 2, 1, 0)), u_o)
    if les:
        assert np.array_equal(got, z["g"]) and np.array_equal(np.transpose(rho, (2, 1, 0)), z["rho"])


@pytest.mark.parametrize("tau", [0.53, 0.8])
def test_emulated_taylor_green_decay_rate(emu_dense, tau):
    """BASELINE's third criterion on the CPU, with the product's kernel source: the z-invariant Taylor-Green vortex (exact
    Navier-Stokes solution) on a periodic 128 x 128 x 2 box, u0 = 0.01, 1000 steps; ln E fitted over steps 200..1000 against
    -4 nu k^2, nu = (tau - 1/2)/3, within 0.5 %.  (The GPU test does the same at 256^3.)"""
    emu = emu_dense
    n, nz, u0, steps, every = 128, 2, 0.01, 1000, 50
    k = 2 * np.pi / n
    x = np.arange(n)[:, None, None] * k; y = np.arange(n)[None, :, None] * k
    ux = (u0 * np.sin(x) * np.cos(y) * np.ones((1, 1, nz))).astype(np.float32)
    uy = (-u0 * np.cos(x) * np.sin(y) * np.ones((1, 1, nz))).astype(np.float32)
    rho0 = (1.0 - (3.0 * u0 * u0 / 4.0) * (np.cos(2 * x) + np.cos(2 * y)) * np.ones((1, 1, nz))).astype(np.float32)
    uinit = np.stack([ux, uy, np.zeros_like(ux)], -1)
    g = R.init_equilibrium_phys(rho0, uinit)
    bufs = [H.to_dev_pop(g), None]; bufs[1] = np.empty_like(bufs[0])
    rho = np.empty((nz, n, n), np.float32); u = np.empty((3, nz, n, n), np.float32)
    f32 = lambda v: C.c_float(float(v))
    ts, es, cur = [], [], 0
    for s in range(0, steps, every):
        r = emu.emu_step_physical_dense(C.c_int(n), C.c_int(n), C.c_int(nz), C.c_int(every), _p(bufs[cur]), _p(bufs[1 - cur]), _p(rho), _p(u),
                                        C.c_int(0), f32(tau), f32(0.18), f32(0.55), f32(1.90), C.c_int(1))
        cur = cur if r == 0 else 1 - cur
        ts.append(s + every - 1); es.append(float((0.5 * rho.astype(np.float64) * (u.astype(np.float64) ** 2).sum(0)).sum()))
    ts, es = np.array(ts, float), np.array(es, float)
    sel = ts >= 200
    slope = np.polyfit(ts[sel], np.log(es[sel]), 1)[0]
    expected = -4.0 * ((tau - 0.5) / 3.0) * k * k
    assert abs(slope / expected - 1.0) <= 5e-3, (slope, expected)
    assert float(np.abs(u[2]).max()) < 1e-4 * u0 and float(np.abs(u[0][0] - u[0][1]).max()) == 0.0
