#!/usr/bin/env python
"""Generates the committed golden fixtures (run in the authoring container, where /root/reference exists).

1. reference_config.json  -- constants read by IMPORTING the reference's own `config` package (it imports without
   Taichi) and by parsing the literal lattice table of src/core/lbm_algorithms.py:158-164 (that module needs Taichi,
   so its table is read from the source text).  These pin the oracle's and the product's constants to the reference.
2. step_reference_24.npz / step_physical_24.npz -- seeded inputs -> outputs of the CPU oracle after 10 steps: regression
   fixtures of the ORACLE (the recordings of the reference's own code live in reference_run_*.npz, written by
   make_reference_goldens.py) that also give the GPU tests a fixture which does not need the oracle at run time.
Nothing under tests/ reads /root/reference at test time.
"""
import io
import json
import os
import re
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def reference_constants():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import config as C
    names = ["NX", "NY", "NZ", "Q_3D", "CS2", "INV_CS2", "TAU_FLUID", "TAU_WATER", "TAU_AIR", "SCALE_LENGTH", "SCALE_TIME",
             "SCALE_VELOCITY", "GRAVITY_LU", "GRAVITY_LU_FULL", "RE_CHAR", "TOP_RADIUS", "BOTTOM_RADIUS", "CUP_HEIGHT",
             "WATER_VISCOSITY_90C", "WATER_DENSITY_90C", "PARTICLE_DIAMETER_MM", "COFFEE_BEAN_DENSITY",
             "COFFEE_PARTICLE_RADIUS", "DT", "RHO_0", "ENABLE_LES", "LES_REYNOLDS_THRESHOLD", "SMAGORINSKY_CONSTANT",
             "PHYSICAL_DOMAIN_SIZE"]
    out = {n: (getattr(C, n) if not isinstance(getattr(C, n), np.generic) else getattr(C, n).item()) for n in names}
    out["CX_3D"] = C.CX_3D.tolist(); out["CY_3D"] = C.CY_3D.tolist(); out["CZ_3D"] = C.CZ_3D.tolist()
    out["WEIGHTS_3D"] = [float(w) for w in C.WEIGHTS_3D]
    src = open(os.path.join(REF, "src/core/lbm_algorithms.py")).read()
    body = src[src.index("def get_d3q19_velocity"):src.index("def get_d3q19_weight")]
    trip = re.findall(r"\[\s*(-?\d)\s*,\s*(-?\d)\s*,\s*(-?\d)\s*\]", body)
    assert len(trip) == 19
    out["EQ_TABLE_lbm_algorithms"] = [[int(a), int(b), int(c)] for a, b, c in trip]
    src = open(os.path.join(REF, "src/physics/les_turbulence.py")).read()
    out["LES_CS_les_turbulence"] = float(re.search(r"self\.cs\s*=\s*([0-9.]+)", src).group(1))
    return out


def oracle_goldens():
    import helpers as H
    from oracle import d3q19_ref as R
    n, steps = 24, 10
    st = H.reference_v60_state(n, seed=123, gravity=2e-5, body=1e-5, phase_mode="none")
    st.phase[:] = np.random.default_rng(5).uniform(0, 0.5, st.phase.shape).astype(np.float32)
    inp = dict(f=st.f.copy(), phase=st.phase.copy(), body_force=st.body_force.copy(), solid=st.solid.copy(),
               filter_zone=st.filter_zone.copy(), les_mask=st.les_mask.copy())
    for _ in range(steps):
        R.step(st)
    np.savez_compressed(os.path.join(HERE, "step_reference_24.npz"), n=n, steps=steps, gravity=2e-5, rho=st.rho, u=st.u,
                        **inp)
    u0 = H.smooth_velocity(n, 0.04, 77); rho0 = H.smooth_density(n, 0.01, 77)
    p = R.PhysParams(nx=n, ny=n, nz=n, tau_water=0.53, les=True)
    g = R.init_equilibrium_phys(rho0, u0)
    for _ in range(steps):
        g, rho, u = R.step_physical(g, p)
    np.savez_compressed(os.path.join(HERE, "step_physical_24.npz"), n=n, steps=steps, tau=0.53, rho0=rho0, u0=u0, g=g, rho=rho, u=u)


if __name__ == "__main__":
    with open(os.path.join(HERE, "reference_config.json"), "w") as fh:
        json.dump(reference_constants(), fh, indent=1, sort_keys=True)
    oracle_goldens()
    print("golden fixtures written to", HERE)
