#!/usr/bin/env python
"""Tuning sweep of the V60 full-feature step (compat = physical): TMA-staged variants and the register-staged kernels."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed, v60_engine  # noqa: E402
from pour_over_coffee_lbm_b200.config import LBMConfig  # noqa: E402
from pour_over_coffee_lbm_b200.engine import D3Q19Engine  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=512); ap.add_argument("--steps", type=int, default=20); ap.add_argument("--box", action="store_true", help="all-fluid box with solid faces instead of the V60 mask")
args = ap.parse_args()
n = args.n
# (vec, block, LBM_TMA_VARIANT): vec = 0 is the TMA-staged kernel (variants: tile 64 x TY, ring depth, CTAs per SM --
# csrc/lbm_step_tma.cu), vec = 2 / 1 the register-staged kernels
cases = [(2, 64, 0), (2, 65, 0), (4, 256, 0)] + ([(0, 0, v) for v in (1, 3, 6)] if os.environ.get("LBM_TMA") == "1" else [])
for strict in (True,):       # compat = physical has a single build
    for vec, block, variant in cases:
        os.environ["LBM_TMA_VARIANT"] = str(variant)
        cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
        eng = D3Q19Engine(n, n, n, compat="physical", periodic=(False, False, False), walls=True, force=True, phase=True, les=True,
                          porous=True, strict=strict, vec=vec, block=block, config=cfg, gravity_lu=1e-5, porous_darcy=0.37, porous_forch=0.9)
        if args.box:
            eng.solid.zero_(); eng.solid[0] = 1; eng.solid[-1] = 1; eng.solid[:, 0] = 1; eng.solid[:, -1] = 1; eng.solid[:, :, 0] = 1; eng.solid[:, :, -1] = 1
            eng.filter_zone.zero_(); eng.filter_zone[n // 2] = 1; eng.pack_flags()
        else:
            eng.build_v60_geometry()
        z = torch.arange(n, device="cuda")[:, None, None]
        eng.phase.copy_(((z < int(0.6 * n)) & (eng.solid == 0)).float())
        g = torch.Generator(device="cuda"); g.manual_seed(1234)
        eng.init_equilibrium(rho=torch.ones((n, n, n), device="cuda"), u=1e-3 * torch.randn((3, n, n, n), device="cuda", generator=g))
        ms = timed(lambda: eng.step(1, write_macro_every=0), args.steps, 5)
        fluid = eng.fluid_cells()
        print(json.dumps({"vec": vec, "block": block, "tma_variant": variant if vec == 0 else None, "ms": round(ms, 4),
                          "MFLUPS": round(fluid / ms / 1e3), "frac": round((fluid * 165 + (n ** 3 - fluid)) / ms / 1e6 / 6540.8, 3)}), flush=True)
        del eng
        torch.cuda.empty_cache()
