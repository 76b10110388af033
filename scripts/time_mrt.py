#!/usr/bin/env python
"""Two-rate MRT collision on the V60 512^3 box (every feature on): the MRT instantiation of the four-cell quad-list kernel against the
two-cell kernel that carried MRT before, and against BGK on the same kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

def dense(timer, m=256):
    """periodic m^3 Taylor-Green box (BASELINE configs[1]): dense kernel, four cells per thread (MRT instantiation) / one cell per thread"""
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine
    for vec in (4, 1):
        for magic in (0.0, 0.1875):
            eng = D3Q19Engine(m, m, m, compat="physical", tau=0.53, vec=vec, mrt_magic=magic)
            rho0, u0 = bench.tgv_fields(m, m, m, 0, m)
            eng.init_equilibrium(rho=rho0.cuda(), u=u0.cuda())
            tm = timer.measure(lambda k: eng.step(k, write_macro_every=0), 50, 5)
            frac = m ** 3 * 152 / tm["ms_per_step"] / 1e6 / 6540.8
            print(f"periodic {m}^3, vec={vec}, {'MRT magic 3/16' if magic else 'BGK'}: {tm['ms_per_step']:.4f} ms (min {tm['ms_min']:.4f}) = {frac:.3f} of the measured HBM peak", flush=True)
            del eng; torch.cuda.empty_cache()


if __name__ == "__main__":
    timer = bench.Timer(1, min_seconds=0.3)
    if "--dense" in sys.argv:
        dense(timer); sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    for vec in (4, 2):
        for magic in (0.0, 0.1875):
            eng = bench.v60_engine(n, vec=vec)
            eng.set_params(mrt_magic=magic)
            tm = timer.measure(lambda k: eng.step(k, write_macro_every=0), 20, 5)
            fl = eng.fluid_cells()
            frac = (fl * 165 + (n ** 3 - fl)) / tm["ms_per_step"] / 1e6 / 6540.8
            print(f"V60 {n}^3 step only, vec={vec}, {'MRT magic 3/16' if magic else 'BGK'}: {tm['ms_per_step']:.4f} ms (min {tm['ms_min']:.4f}) = {frac:.3f} of the measured HBM peak, "
                  f"finite: {bool(torch.isfinite(eng.populations[:, eng.solid == 0]).all())}", flush=True)
            del eng; torch.cuda.empty_cache()
