"""Ad-hoc parity diagnostics (prints error magnitudes / mismatch locations).  Test infrastructure: lives under tests/ because it
calls the CPU oracle (only tests/, smoke() and bench.py's cpu_baseline leg may)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import d3q19_ref as R, ref_cpu as RC
import test_gpu_step_parity as T
from pour_over_coffee_lbm_b200.engine import D3Q19Engine
from pour_over_coffee_lbm_b200 import _lib as L

def physical_case(vec, strict, steps=30, n=32):
    cfg, solid, zone, les_mask, phase, bf, u0, rho0 = T._physical_v60_case(n, 11)
    p = R.PhysParams(nx=n, ny=n, nz=n, tau_water=0.53, tau_air=0.8, gravity_lu=1e-5, periodic=(False, False, False),
                     use_force=True, use_phase=True, les=True, porous=True, porous_darcy=0.37, porous_forch=0.9)
    g = R.init_equilibrium_phys(rho0, u0)
    for _ in range(steps):
        g, rho, u = R.step_physical(g, p, solid=solid, body_force=bf, phase=phase, filter_zone=zone, les_mask=les_mask)
    eng = D3Q19Engine(n, n, n, compat="physical", periodic=(False, False, False), walls=True, force=True, phase=True, les=True,
                      porous=True, strict=strict, vec=vec, tau=0.53, tau_air=0.8, gravity_lu=1e-5, porous_darcy=0.37, porous_forch=0.9)
    t = T._torch
    eng.solid.copy_(t(H.to_dev_scalar(solid))); eng.filter_zone.copy_(t(H.to_dev_scalar(zone)))
    eng.les_mask.copy_(t(H.to_dev_scalar(les_mask))); eng.pack_flags()
    eng.phase.copy_(t(H.to_dev_scalar(phase))); eng.body_force.copy_(t(H.to_dev_vec(bf)))
    eng.init_equilibrium(rho=t(H.to_dev_scalar(rho0)), u=t(H.to_dev_vec(u0)))
    eng.step(steps)
    fluid = solid == 0
    gg = H.from_dev_pop(eng.populations); rr = H.from_dev_scalar(eng.rho); uu = H.from_dev_vec(eng.u)
    flags = H.from_dev_scalar(eng.flags)
    d = np.abs(gg - g); d[:, ~fluid] = 0
    print(f"physical vec={vec} strict={strict} steps={steps}: max|dg|={d.max():.3e} rho rel={H.rel_err(rr[fluid], rho[fluid]):.3e} "
          f"u rel={H.rel_err(uu[fluid], u[fluid]):.3e} umax={np.abs(u[fluid]).max():.3e} nbad={(d>0).sum()}")
    if strict and d.max() > 0:
        q, i, j, k = np.unravel_index(np.argmax(d), d.shape)
        bad = np.argwhere(d.max(0) > 0)
        print("  worst at q,i,j,k", q, i, j, k, "flag", flags[i, j, k], "n bad cells", len(bad), "of", fluid.sum())
        near = (flags & L.FLAG_NEAR) != 0
        print("  bad cells near-wall fraction", near[tuple(bad.T)].mean(), " x%4 hist", np.bincount(bad[:, 0] % 4, minlength=4))

for steps in (1, 30):
    for vec in (1, 4):
        for strict in (True, False):
            physical_case(vec, strict, steps)
