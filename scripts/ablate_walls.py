#!/usr/bin/env python
"""Feature ablation of the compat = physical walls kernels on an all-fluid 512^3 box with solid faces:
which input / feature costs what.  Prints ms per step and the fraction of the measured HBM peak on the bytes moved."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed  # noqa: E402
from pour_over_coffee_lbm_b200.config import LBMConfig  # noqa: E402
from pour_over_coffee_lbm_b200.engine import D3Q19Engine  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=512); ap.add_argument("--vecs", default="4,2")
args = ap.parse_args()
n = args.n
cases = [("walls", {}), ("walls+les", dict(les=True)), ("walls+force", dict(force=True)), ("walls+force+phase", dict(force=True, phase=True)),
         ("walls+force+phase+les", dict(force=True, phase=True, les=True)),
         ("all", dict(force=True, phase=True, les=True, porous=True))]
for vec in [int(v) for v in args.vecs.split(",")]:
    for name, kw in cases:
        cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
        pk = dict(porous_darcy=0.37, porous_forch=0.9) if kw.get("porous") else {}
        eng = D3Q19Engine(n, n, n, compat="physical", periodic=(False, False, False), walls=True, vec=vec, config=cfg, gravity_lu=1e-5, **kw, **pk)
        eng.solid.zero_(); eng.solid[0] = 1; eng.solid[-1] = 1; eng.solid[:, 0] = 1; eng.solid[:, -1] = 1; eng.solid[:, :, 0] = 1; eng.solid[:, :, -1] = 1
        eng.filter_zone.zero_(); eng.filter_zone[n // 2] = 1; eng.pack_flags()
        if kw.get("phase"):
            z = torch.arange(n, device="cuda")[:, None, None]
            eng.phase.copy_(((z < int(0.6 * n)) & (eng.solid == 0)).float())
        g = torch.Generator(device="cuda"); g.manual_seed(1234)
        eng.init_equilibrium(rho=torch.ones((n, n, n), device="cuda"), u=1e-3 * torch.randn((3, n, n, n), device="cuda", generator=g))
        ms = timed(lambda: eng.step(1, write_macro_every=0), 20, 5)
        fluid = eng.fluid_cells()
        b = 152 + 1 + (12 if kw.get("force") else 0) + (4 if kw.get("phase") else 0)
        print(json.dumps({"vec": vec, "case": name, "ms": round(ms, 4), "bytes_per_cell": b,
                          "frac_of_measured_peak": round(fluid * b / ms / 1e6 / 6540.8, 3)}), flush=True)
        del eng
        torch.cuda.empty_cache()
