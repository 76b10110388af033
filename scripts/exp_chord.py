#!/usr/bin/env python
"""Round-2 experiment: what does the chord kernel's structure cost by itself?  Fully periodic all-fluid 512^3 box (no wall, no link,
every tile full): dense kernel against the chord kernel with the features switched on one by one."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed  # noqa: E402
from bench import measured_peak  # noqa: E402
from pour_over_coffee_lbm_b200.engine import D3Q19Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
peak, _ = measured_peak()
cases = [("dense BGK", dict()), ("dense LES", dict(les=True)),
         ("chord BGK", dict(walls=True)), ("chord LES", dict(walls=True, les=True)),
         ("chord LES+phase", dict(walls=True, les=True, phase=True)),
         ("chord LES+phase+force", dict(walls=True, les=True, phase=True, force=True)),
         ("chord LES+phase+force+porous", dict(walls=True, les=True, phase=True, force=True, porous=True, porous_darcy=0.37, porous_forch=0.9)),
         ("two-cell walls kernel, all features", dict(walls=True, les=True, phase=True, force=True, porous=True, porous_darcy=0.37, porous_forch=0.9, vec=2))]
for name, kw in cases:
    eng = D3Q19Engine(n, n, n, compat="physical", tau=0.53, gravity_lu=1e-5, **kw)
    if kw.get("porous"):
        eng.filter_zone[n // 2] = 1; eng.pack_flags()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    eng.init_equilibrium(rho=torch.ones((n, n, n), device="cuda"), u=1e-3 * torch.randn((3, n, n, n), device="cuda", generator=g))
    if eng.phase is not None: eng.phase.fill_(1.0)
    b = 152 + (1 if kw.get("walls") else 0) + (4 if kw.get("phase") else 0) + (12 if kw.get("force") else 0)
    ms = timed(lambda: eng.step(1, write_macro_every=0), 20, 5)
    print(json.dumps({"case": name, "ms": round(ms, 4), "bytes_per_cell_moved": b, "GBs": round(n ** 3 * b / ms / 1e6), "frac_of_measured_peak": round(n ** 3 * b / ms / 1e6 / peak, 4)}), flush=True)
    del eng; torch.cuda.empty_cache()
