"""The CUDA kernel SOURCE of csrc/lbm_producers.cu, compiled by the host compiler and executed thread by thread on the CPU
(tests/emu/emu_producers.cpp), against the recorded runs of the reference -- the same fixtures the GPU tests use.

Why: the authoring container has no GPU.  The GPU tests (tests/test_producers.py, -m gpu) remain the parity proof for the
compiled sm_100a code; this file checks the kernels' logic -- the 4-cells-per-thread path with its edge lanes and masked
stores, the one-cell path, the launch geometry helpers -- on every CPU run, bit for bit.  Test infrastructure only: the
emulation library is never on the product path.
"""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "reference_run_multiphase.npz")
GOLD_FP = os.path.join(HERE, "golden", "reference_run_filter_particles.npz")


@pytest.fixture(scope="module")
def emu():
    return H.build_emu("emu_producers", ['lbm_producers.cu', 'lbm_common.cuh'])


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f(v):
    return C.c_float(float(v))


@pytest.mark.parametrize("vec", [4, 1])
def test_emulated_multiphase_kernels_reproduce_the_reference_run(emu, vec):
    z = np.load(GOLD)
    n = int(z["n"])
    dims = (C.c_int(vec), C.c_int(n), C.c_int(n), C.c_int(n))
    sigma, mob, dt = float(z["sigma"]), float(z["mobility"]), float(z["dt"])
    rw, ra = C.c_double(float(z["rho_water"])), C.c_double(float(z["rho_air"]))
    flags = H.to_dev_scalar(z["solid"]).astype(np.uint8)
    phi = H.to_dev_scalar(z["phi"]); phi_new = H.to_dev_scalar(z["phi_new_in"]); u = H.to_dev_vec(z["u"])
    rho = H.to_dev_scalar(z["rho"]); bf = H.to_dev_vec(z["body_force"]); phase = np.zeros_like(rho)
    sc = lambda: np.zeros_like(rho)
    vc = lambda: np.zeros_like(bf)
    mu, lap, curv = sc(), sc(), sc()
    gphi, gmu, nrm, sf = vc(), vc(), vc(), vc()
    emu.emu_chemical_potential(*dims, _p(phi), _p(lap), _p(mu), _f(3.0 * sigma * float(z["interface_width"]) / 8.0))
    assert np.array_equal(mu, H.to_dev_scalar(z["mu"])) and np.array_equal(lap, H.to_dev_scalar(z["laplacian_phi"]))

    def chain(body_force):
        emu.emu_surface_tension(*dims, _p(phi), _p(mu), _p(rho), _p(flags), _p(gphi), _p(gmu), _p(nrm), _p(curv), _p(sf), _p(body_force), _f(sigma))
    bf_lean = H.to_dev_vec(z["body_force"])                             # the field-free kernel: body_force only, same bits
    emu.emu_surface_tension_lean(*dims, _p(phi), _p(rho), _p(flags), None, None, _p(bf_lean), _f(sigma))
    assert np.array_equal(bf_lean, H.to_dev_vec(z["st_body_force"]))
    chain(bf)
    for name, got in (("grad_phi", gphi), ("grad_mu", gmu), ("normal", nrm), ("surface_force", sf), ("body_force", bf)):
        assert np.array_equal(got, H.to_dev_vec(z["st_" + name])), name
    assert np.array_equal(curv, H.to_dev_scalar(z["st_curvature"]))
    chain(None)                                                          # MultiphaseFlow3D.step(20, precollision_applied=True)
    emu.emu_phase_field_step(*dims, _p(phi), _p(phi_new), _p(mu), _p(u), _p(rho), _p(phase), _f(mob), _f(dt), rw, ra)
    for name, got in (("phi", phi), ("phi_new", phi_new), ("rho", rho), ("phase", phase)):
        assert np.array_equal(got, H.to_dev_scalar(z["s1_" + name])), name
    assert np.array_equal(bf, H.to_dev_vec(z["s1_body_force"]))
    chain(bf)                                                            # step(21, precollision_applied=False)
    emu.emu_phase_field_step(*dims, _p(phi), _p(phi_new), _p(mu), _p(u), _p(rho), _p(phase), _f(mob), _f(dt), rw, ra)
    for name, got in (("phi", phi), ("rho", rho), ("phase", phase), ("curvature", curv)):
        assert np.array_equal(got, H.to_dev_scalar(z["s2_" + name])), name
    assert np.array_equal(bf, H.to_dev_vec(z["s2_body_force"])) and np.array_equal(sf, H.to_dev_vec(z["s2_surface_force"]))
    # stand-alone entry points
    bf2 = H.to_dev_vec(z["s1_body_force"]); rho1 = H.to_dev_scalar(z["s1_rho"])
    emu.emu_apply_surface_tension(*dims, _p(sf), _p(rho1), _p(flags), _p(bf2))
    assert np.array_equal(bf2, H.to_dev_vec(z["s2_body_force"]))
    rho3, phase3 = sc(), sc()
    emu.emu_density_from_phase(*dims, _p(phi), _p(rho3), _p(phase3), rw, ra)
    assert np.array_equal(rho3, H.to_dev_scalar(z["s2_rho"])) and np.array_equal(phase3, H.to_dev_scalar(z["s2_phase"]))
    # mu == NULL is the all-zero chemical potential
    p_a, p_b = H.to_dev_scalar(z["phi"]), H.to_dev_scalar(z["phi"])
    n_a, n_b = H.to_dev_scalar(z["phi_new_in"]), H.to_dev_scalar(z["phi_new_in"])
    emu.emu_phase_field_step(*dims, _p(p_a), _p(n_a), None, _p(u), _p(sc()), _p(sc()), _f(mob), _f(dt), rw, ra)
    emu.emu_phase_field_step(*dims, _p(p_b), _p(n_b), _p(sc()), _p(u), _p(sc()), _p(sc()), _f(mob), _f(dt), rw, ra)
    assert np.array_equal(p_a, p_b)


def test_emulated_pouring_kernels_reproduce_the_reference_run(emu):
    """The nozzle kernels over their bounding box: glibc's expf here, NumPy's in the recording (both within 1 ulp)."""
    from oracle import producers_ref as P
    z = np.load(GOLD)
    n = int(z["n"])
    flags = H.to_dev_scalar(z["solid"]).astype(np.uint8)
    bf = H.to_dev_vec(z["s2_body_force"]); phi = H.to_dev_scalar(z["s2_phi"])
    p = P.PourState(n, float(z["pour_diameter"]), int(z["pour_height"]), float(z["pour_velocity"]))
    p.start_pouring(pattern="center", flow_rate=0.3)
    p.pour_time = np.float32(p.pour_time + np.float32(0.1))
    x, y = p.position()
    args = (C.c_int(n), C.c_int(n), C.c_int(n), _f(x), _f(y), _f(np.float32(p.POUR_DIAMETER_GRID / 2.0)), C.c_int(p.POUR_HEIGHT),
            _f(np.float32(p.POUR_VELOCITY)), _f(p.flow_rate), _f(np.float32(0.1)))
    emu.emu_pour(C.c_int(0), *args, _p(flags), _p(bf))
    emu.emu_pour(C.c_int(1), *args, _p(flags), _p(phi))
    want_bf, want_phi = H.to_dev_vec(z["p1_body_force"]), H.to_dev_scalar(z["p1_phi"])
    base_bf = H.to_dev_vec(z["s2_body_force"])
    assert np.array_equal(bf[want_bf == base_bf], want_bf[want_bf == base_bf])        # untouched cells bit-identical
    eps = np.finfo(np.float32).eps                                                  # exp's ulp rides on the INCREMENT, not on the sum
    assert np.all(np.abs(bf - want_bf) <= 8 * eps * np.maximum(np.abs(want_bf - base_bf), np.abs(want_bf)))
    assert np.all(np.abs(phi - want_phi) <= 8 * eps * np.maximum(np.abs(want_phi - H.to_dev_scalar(z["s2_phi"])), np.abs(want_phi)))
    assert (want_bf != base_bf).sum() > 50


@pytest.mark.parametrize("cells", [16, 4, 1])
def test_emulated_filter_kernels_reproduce_the_reference_run(emu, cells):
    from oracle import producers_ref as P
    z = np.load(GOLD_FP)
    n = int(z["n"]); npart = z["p_pos"].shape[0]
    flags = (2 * H.to_dev_scalar(z["filter_zone"])).astype(np.uint8)                 # LBM_FLAG_FILTER = 2
    dims = (C.c_int(n), C.c_int(n), C.c_int(n))
    pos = np.ascontiguousarray(z["p_pos"].T); vel = np.ascontiguousarray(z["p_vel"].T); act = z["p_active"].astype(np.int32).copy()
    acc = H.to_dev_scalar(z["accumulated_in"]); blk = H.to_dev_scalar(z["blockage_in"])
    for t in range(2):
        emu.emu_particles_block_at_filter(*dims, C.c_int(npart), _p(pos), _p(vel), _p(act), _p(flags), _p(acc), _f(np.float32(z["scale_length"])),
                                          _f(0.0), C.c_uint(0))
        assert np.array_equal(vel.T, z[f"b{t}_vel"]) and np.array_equal(acc, H.to_dev_scalar(z[f"b{t}_accumulated"]))
    for t in range(2):
        emu.emu_dynamic_resistance(C.c_int(cells), *dims, _p(flags), _p(blk), _p(acc))
        assert np.array_equal(acc, H.to_dev_scalar(z[f"r{t}_accumulated"]))
        assert np.allclose(blk, H.to_dev_scalar(z[f"r{t}_blockage"]), rtol=1e-6, atol=1e-8)
    # the counter-based kick equals the oracle's draw bit for bit
    vel = np.ascontiguousarray(z["p_vel"].T); acc = H.to_dev_scalar(z["accumulated_in"])
    emu.emu_particles_block_at_filter(*dims, C.c_int(npart), _p(pos), _p(vel), _p(act), _p(flags), _p(acc), _f(np.float32(z["scale_length"])),
                                      _f(0.01), C.c_uint(7))
    want = z["p_vel"].copy(); a = z["accumulated_in"].copy()
    P.block_particles_at_filter(z["filter_zone"], z["p_pos"], want, z["p_active"], a, float(z["scale_length"]), 0.01, seed=7)
    assert np.array_equal(vel.T, want) and np.array_equal(acc, H.to_dev_scalar(a))


def test_emulated_vec4_equals_vec1_on_a_non_cubic_box(emu):
    """Index arithmetic of the 4-cells-per-thread path on nx != ny != nz (nx = 24: six vectors per row, both edge lanes live),
    random fields everywhere (outer layers included): identical to the one-cell path, and the NumPy oracle agrees with both."""
    from oracle import producers_ref as P
    nx, ny, nz = 24, 10, 7
    rng = np.random.default_rng(3)
    sh = (nx, ny, nz)
    phi = np.clip(rng.normal(0.0, 0.8, sh), -1, 1).astype(np.float32); phi_new0 = rng.normal(0, 0.1, sh).astype(np.float32)
    mu0 = rng.normal(0, 0.05, sh).astype(np.float32); u = rng.normal(0, 0.05, sh + (3,)).astype(np.float32)
    rho = (1 + 0.1 * rng.standard_normal(sh)).astype(np.float32); rho[3, 4, 2] = 0.0
    bf0 = (1e-3 * rng.standard_normal(sh + (3,))).astype(np.float32); sf0 = (1e-3 * rng.standard_normal(sh + (3,))).astype(np.float32)
    solid = (rng.random(sh) < 0.4).astype(np.uint8)
    nrm0 = rng.normal(0, 0.5, sh + (3,)).astype(np.float32)              # what the never-written outer layer of `normal` holds
    outs = []
    for vec in (4, 1):
        dims = (C.c_int(vec), C.c_int(nx), C.c_int(ny), C.c_int(nz))
        d_phi, d_new, d_mu, d_u = H.to_dev_scalar(phi), H.to_dev_scalar(phi_new0), H.to_dev_scalar(mu0), H.to_dev_vec(u)
        d_rho, d_bf, d_sf, d_flags = H.to_dev_scalar(rho), H.to_dev_vec(bf0), H.to_dev_vec(sf0), H.to_dev_scalar(solid)
        d_phase = np.zeros_like(d_rho); d_curv = np.zeros_like(d_rho)
        d_g, d_gm, d_n = np.zeros_like(d_bf), np.zeros_like(d_bf), H.to_dev_vec(nrm0)
        d_bf_lean = H.to_dev_vec(bf0)
        emu.emu_surface_tension_lean(*dims, _p(d_phi), _p(d_rho), _p(d_flags), _p(H.to_dev_vec(nrm0)), _p(H.to_dev_vec(sf0)), _p(d_bf_lean), _f(0.05))
        emu.emu_surface_tension(*dims, _p(d_phi), _p(d_mu), _p(d_rho), _p(d_flags), _p(d_g), _p(d_gm), _p(d_n), _p(d_curv), _p(d_sf), _p(d_bf), _f(0.05))
        assert np.array_equal(d_bf_lean, d_bf)
        emu.emu_phase_field_step(*dims, _p(d_phi), _p(d_new), _p(d_mu), _p(d_u), _p(d_rho), _p(d_phase), _f(0.001), _f(1.0), C.c_double(1.0),
                                 C.c_double(0.00125))
        outs.append([a.copy() for a in (d_phi, d_new, d_rho, d_phase, d_bf, d_sf, d_curv, d_g, d_gm, d_n)])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    m = P.MultiphaseState(sh); m.phi = phi.copy(); m.phi_new = phi_new0.copy(); m.mu = mu0.copy(); m.surface_force = sf0.copy()
    m.normal = nrm0.copy()
    bf = bf0.copy(); r = rho.copy(); ph = np.zeros_like(r)
    P.accumulate_surface_tension_pre_collision(m, r, solid, bf, 0.05)
    P.update_phase_field_cahn_hilliard(m, u, 0.001, 1.0); P.apply_phase_separation(m, 1.0); m.phi[...] = m.phi_new
    P.update_density_from_phase(m, r, ph, 1.0, 0.00125)
    got = outs[0]
    assert np.array_equal(got[0], H.to_dev_scalar(m.phi)) and np.array_equal(got[2], H.to_dev_scalar(r)) and np.array_equal(got[3], H.to_dev_scalar(ph))
    assert np.array_equal(got[4], H.to_dev_vec(bf)) and np.array_equal(got[5], H.to_dev_vec(m.surface_force))
    assert np.array_equal(got[6], H.to_dev_scalar(m.curvature)) and np.array_equal(got[9], H.to_dev_vec(m.normal))


def test_emulated_one_launch_surface_tension_equals_the_chain_on_a_v60_box(emu):
    """48^3 with the V60 mask and a wavy, noisy interface through the cone: the one-launch kernel's body_force is bit-identical
    to the four-kernel chain's, and only the interface band is touched."""
    from oracle import d3q19_ref as R
    n = 48
    solid = R.v60_solid(R.RefConfig(NX=n, NY=n, NZ=n))
    rng = np.random.default_rng(11)
    x = np.arange(n, dtype=np.float32)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    phi = np.tanh((0.5 * n + 3.0 * np.sin(0.3 * X) + 2.0 * np.cos(0.25 * Y) - Z) / 2.0) + 0.02 * rng.standard_normal((n, n, n))
    phi = np.clip(phi, -1, 1).astype(np.float32)
    rho = (1 + 0.05 * rng.standard_normal((n, n, n))).astype(np.float32)
    bf0 = (1e-4 * rng.standard_normal((n, n, n, 3))).astype(np.float32)
    dims = (C.c_int(4), C.c_int(n), C.c_int(n), C.c_int(n))
    d_phi, d_rho, d_flags = H.to_dev_scalar(phi), H.to_dev_scalar(rho), H.to_dev_scalar(solid).astype(np.uint8)
    chain, lean = H.to_dev_vec(bf0), H.to_dev_vec(bf0)
    z3 = lambda: np.zeros_like(chain)
    emu.emu_surface_tension(*dims, _p(d_phi), None, _p(d_rho), _p(d_flags), _p(z3()), None, _p(z3()), _p(np.zeros_like(d_rho)), _p(z3()), _p(chain),
                            _f(0.05))
    emu.emu_surface_tension_lean(*dims, _p(d_phi), _p(d_rho), _p(d_flags), None, None, _p(lean), _f(0.05))
    assert np.array_equal(chain, lean)
    changed = (chain != H.to_dev_vec(bf0)).any(0).sum()
    assert 500 < changed < 0.2 * n ** 3
