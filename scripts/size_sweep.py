#!/usr/bin/env python
"""Does the fraction of the HBM peak depend on the box size?  Dense periodic kernel and featureless walls kernel at
several sizes (working set 2 x 19 x 4 B x cells)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed  # noqa: E402
from pour_over_coffee_lbm_b200.engine import D3Q19Engine  # noqa: E402

for shape in ((256, 256, 256), (512, 256, 256), (512, 512, 256), (512, 512, 512), (1024, 1024, 128), (256, 256, 2048)):
    nx, ny, nz = shape
    for kind in ("dense", "walls"):
        if kind == "dense":
            eng = D3Q19Engine(nx, ny, nz, compat="physical", tau=0.6)
            b = 152
        else:
            eng = D3Q19Engine(nx, ny, nz, compat="physical", periodic=(False, False, False), walls=True, tau=0.6)
            eng.solid.zero_(); eng.solid[0] = 1; eng.solid[-1] = 1; eng.solid[:, 0] = 1; eng.solid[:, -1] = 1; eng.solid[:, :, 0] = 1; eng.solid[:, :, -1] = 1
            eng.pack_flags()
            b = 153
        eng.init_equilibrium(1.0, (0.01, 0.0, 0.0))
        ms = timed(lambda: eng.step(1, write_macro_every=0), 20, 5)
        fluid = eng.fluid_cells()
        print(json.dumps({"shape": shape, "kind": kind, "ms": round(ms, 4), "GB": round(fluid * b / 1e9, 2),
                          "frac_of_measured_peak": round(fluid * b / ms / 1e6 / 6540.8, 3)}), flush=True)
        del eng
        torch.cuda.empty_cache()
