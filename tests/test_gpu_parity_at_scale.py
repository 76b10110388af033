"""Step parity at the grid sizes BASELINE.json names (rho, u, f were compared at 16^3 .. 48^3 only).

  * compat = reference, the reference's own 224^3 V60 box, 20 calls of LBMSolver.step(): strict build against oracle/ref_cpu.c, bit-exact;
  * compat = physical, periodic 256^3 (dense kernel, 5 steps) and V60 512^3 with every feature (chord kernel: tile list, wall links
    and 32-bit cell indices at 134 M cells; 2 steps) against oracle/phys_cpu.c -- the C twin of d3q19_ref.step_physical, run
    directly on a download of the device buffers -- bit-exact on every fluid cell;
  * the pressure-gradient drive fused into the step kernel against producer + step, bit-exact.
The oracle only checks; the populations it starts from are the device's own (downloaded after lbm_init_equilibrium)."""
import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R
from oracle import ref_cpu as RC

pytestmark = pytest.mark.gpu


def _torch(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _enough_host_memory(gb):
    try:
        import psutil
        return psutil.virtual_memory().available > gb * 2 ** 30
    except Exception:
        return True


def test_reference_mode_224_v60_20_steps_bit_exact():
    """BASELINE configs[0]: the reference's default 224^3 box with the V60 mask, filter zone, phase field (air-phase relaxation, gravity
    active), body force; FD-LES + macroscopic + collide/stream + filter damping + face BCs, 20 steps."""
    from pour_over_coffee_lbm_b200.config import LBMConfig
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine
    n, steps = 224, 20
    st = H.reference_v60_state(n, seed=6, gravity=2e-5, body=1e-5, phase_mode="split")
    st.phase *= np.float32(0.3)
    cs = RC.CState(st)
    cs.step(steps)
    eng = D3Q19Engine(n, n, n, compat="reference", periodic=(False, False, False), walls=True, force=True, phase=True, les=True, porous=True,
                      strict=True, config=LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=2e-5))
    eng.solid.copy_(_torch(H.to_dev_scalar(st.solid))); eng.filter_zone.copy_(_torch(H.to_dev_scalar(st.filter_zone)))
    eng.les_mask.copy_(_torch(H.to_dev_scalar(st.les_mask))); eng.pack_flags()
    eng.phase.copy_(_torch(H.to_dev_scalar(st.phase))); eng.body_force.copy_(_torch(H.to_dev_vec(st.body_force)))
    eng.import_f(_torch(H.to_dev_pop(st.f)))
    for _ in range(steps):
        eng.step(1); eng.face_bc()
    fluid = cs.solid == 0
    assert np.array_equal(H.from_dev_scalar(eng.rho)[fluid], cs.rho[fluid])
    assert np.array_equal(H.from_dev_vec(eng.u)[fluid], cs.u[fluid])
    f = H.from_dev_pop(eng.export_f())
    assert np.array_equal(f[:, fluid], cs.f[:, fluid])


def _run_physical_against_c_twin(eng, p, steps, walls):
    import torch
    g = eng.populations.cpu().numpy()
    kw = {}
    if walls:
        kw = dict(solid=eng.solid.cpu().numpy(), body_force=eng.body_force.cpu().numpy(), phase=eng.phase.cpu().numpy(),
                  filter_zone=eng.filter_zone.cpu().numpy(), les_mask=eng.les_mask.cpu().numpy())
    for _ in range(steps):
        g, rho, u = RC.phys_step(g, p, layout="device", **kw)
    eng.step(steps)
    torch.cuda.synchronize()
    solid = kw.get("solid")
    assert RC.count_mismatch(eng.populations.cpu().numpy(), g, solid) == 0
    assert RC.count_mismatch(eng.rho.cpu().numpy()[None], rho[None], solid) == 0
    assert RC.count_mismatch(eng.u.cpu().numpy(), u, solid) == 0


def test_physical_periodic_256_dense_kernel_bit_exact():
    """BASELINE configs[1]: periodic 256^3 Taylor-Green state, BGK + local-stress LES, the headline dense kernel, 5 steps."""
    from bench import tgv_fields
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine
    n = 256
    eng = D3Q19Engine(n, n, n, compat="physical", les=True, tau=0.53)
    rho0, u0 = tgv_fields(n, n, n, 0, n)
    eng.init_equilibrium(rho=rho0.cuda(), u=u0.cuda())
    p = R.PhysParams(nx=n, ny=n, nz=n, tau_water=0.53, les=True, cs_smag=float(np.float32(eng.params.cs_smag)))
    _run_physical_against_c_twin(eng, p, 5, walls=False)


@pytest.mark.parametrize("n", [512])
def test_physical_v60_512_every_feature_bit_exact(n):
    """BASELINE configs[2] at full size: the V60 mask at 512^3 (47.7 M fluid cells, 461 598 chord tiles, 3.98 M wall links), gravity *
    phase + a random body force + LES + porous drag + bounce-back, 2 steps of the chord kernel against the C twin."""
    if not _enough_host_memory(48):
        pytest.skip("needs ~35 GB of host memory for three copies of the 512^3 populations")
    from bench import v60_engine
    import torch
    eng = v60_engine(n)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    eng.body_force.copy_(2e-6 * torch.randn(eng.body_force.shape, device="cuda", generator=g))
    par = eng.params
    p = R.PhysParams(nx=n, ny=n, nz=n, tau_water=par.tau_water, tau_air=par.tau_air, gravity_lu=par.gravity_lu, periodic=(False, False, False),
                     use_force=True, use_phase=True, les=True, cs_smag=par.cs_smag, tau_min=par.tau_min, tau_max=par.tau_max, porous=True,
                     porous_darcy=par.porous_darcy, porous_forch=par.porous_forch)
    _run_physical_against_c_twin(eng, p, 2, walls=True)


@pytest.mark.parametrize("with_body_force,mrt_magic", [(True, 0.0), (False, 0.0), (True, 0.1875)], ids=["body_force", "drive_only", "body_force_mrt"])
def test_fused_pressure_gradient_drive_equals_producer_plus_step(with_body_force, mrt_magic):
    """LBM_FEAT_DRIVE (the drive evaluated inside the step kernel from the previous step's rho) against
    lbm_pressure_gradient_force(_set) + lbm_step on a V60 96^3 box, 10 steps: populations, rho, u bit for bit."""
    import torch
    from bench import v60_engine
    n, steps = 96, 10

    def make(drive, force):
        e = v60_engine(n, drive=drive, force=force)
        g = torch.Generator(device="cuda"); g.manual_seed(9)
        e.init_equilibrium(rho=1.0 + 1e-3 * torch.randn((n, n, n), device="cuda", generator=g),
                           u=1e-3 * torch.randn((3, n, n, n), device="cuda", generator=g))
        if e.body_force is not None:
            e.body_force.copy_(1e-6 * torch.randn((3, n, n, n), device="cuda", generator=g))
        e.set_params(drive_max_force=0.12, drive_scale=0.5, mrt_magic=mrt_magic)      # mrt_magic > 0: the MRT instantiation of the four-cell kernel
        return e
    a = make(False, True)
    base = a.body_force.clone()
    for _ in range(steps):
        if with_body_force:
            a.body_force.copy_(base); a.add_pressure_gradient_force(0.12, 0.5)
        else:
            a.set_pressure_gradient_force(0.12, 0.5)
        a.step(1)
    b = make(True, with_body_force)
    b.step(steps)
    fluid = a.solid == 0
    assert torch.equal(a.populations[:, fluid], b.populations[:, fluid]) and torch.equal(a.rho[fluid], b.rho[fluid])
    assert torch.equal(a.u[:, fluid], b.u[:, fluid])
