"""periodic 256^3 dense kernel (bench calibration) timing under the library LBM_B200_LIB points to"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed
from pour_over_coffee_lbm_b200.engine import D3Q19Engine
e = D3Q19Engine(256, 256, 256, compat="physical", tau=0.53)
e.init_equilibrium(rho0=1.0, u0=(0.01, 0.0, 0.0))
ms = timed(lambda: e.step(1, write_macro_every=0), 50, 10)
print(json.dumps({"case": "tgv_256_dense", "ms": round(ms, 4), "frac": round(152 * 256**3 / ms / 1e6 / 6540.8, 4)}))
