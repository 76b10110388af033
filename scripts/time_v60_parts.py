#!/usr/bin/env python
"""Where does the time of the full V60 512^3 step sequence go?  Times each call of the sequence on its own."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed, v60_engine  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
eng = v60_engine(n, True, 0)
eng.step(2, write_macro_every=1)
parts = {
    "clear_body_force": lambda: eng.clear_body_force(),
    "add_pressure_gradient_force": lambda: eng.add_pressure_gradient_force(0.12, 1.0),
    "set_pressure_gradient_force": lambda: eng.set_pressure_gradient_force(0.12, 1.0),
    "step(write_macro=0)": lambda: eng.step(1, write_macro_every=0),
    "step(write_macro=1)": lambda: eng.step(1, write_macro_every=1),
}
for k, fn in parts.items():
    print(json.dumps({"part": k, "ms": round(timed(fn, 20, 5), 4)}), flush=True)
def full():
    eng.set_pressure_gradient_force(0.12, 1.0); eng.step(1, write_macro_every=1)
print(json.dumps({"part": "full sequence", "ms": round(timed(full, 20, 5), 4)}), flush=True)
