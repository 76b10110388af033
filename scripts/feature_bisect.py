"""Time the step on an all-fluid 256^3 box (solid faces) for feature subsets, VEC=1/4 (diagnostic)."""
import json, os, sys, itertools
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed
from pour_over_coffee_lbm_b200.config import LBMConfig
from pour_over_coffee_lbm_b200.engine import D3Q19Engine
n = 256
for vec in (1, 4):
    for walls, force, phase, les, porous in [(0,0,0,0,0),(0,0,0,1,0),(1,0,0,0,0),(1,0,0,1,0),(1,1,0,0,0),(1,1,1,0,0),(1,1,1,1,0),(1,0,0,0,1),(1,1,1,1,1)]:
        cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
        eng = D3Q19Engine(n, n, n, compat="physical", periodic=(True,)*3 if not walls else (False,)*3, walls=bool(walls), force=bool(force), phase=bool(phase),
                          les=bool(les), porous=bool(porous), strict=True, vec=vec, config=cfg, gravity_lu=1e-5, porous_darcy=0.37, porous_forch=0.9)
        if walls:
            eng.solid.zero_(); eng.solid[0] = 1; eng.solid[-1] = 1; eng.solid[:, 0] = 1; eng.solid[:, -1] = 1; eng.solid[:, :, 0] = 1; eng.solid[:, :, -1] = 1
            eng.filter_zone.zero_(); eng.filter_zone[n // 2] = 1; eng.pack_flags()
        if phase: eng.phase.fill_(1.0)
        g = torch.Generator(device="cuda"); g.manual_seed(1234)
        eng.init_equilibrium(rho=torch.ones((n, n, n), device="cuda"), u=1e-3 * torch.randn((3, n, n, n), device="cuda", generator=g))
        l0 = eng.launch_count()
        ms = timed(lambda: eng.step(1, write_macro_every=0), 40, 5)
        print(json.dumps({"vec": vec, "walls": walls, "force": force, "phase": phase, "les": les, "porous": porous, "ms": round(ms, 4),
                          "GLUPS": round(n**3 / ms / 1e6, 2), "launches_per_step": (eng.launch_count() - l0) / 45}), flush=True)
        del eng; torch.cuda.empty_cache()
