#!/usr/bin/env python
"""Round-2 experiments on the chord kernel: what does the dependent tile-entry load cost?
  LBM_REGULAR=1    all-fluid box only: tile coordinates computed from the warp index, the entry is loaded after the population loads
  LBM_L2_WINDOW=1  persisting L2 access-policy window over the tile list
Prints one JSON line per case; results of the REGULAR variant are compared bit for bit with the baseline."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed  # noqa: E402
from scripts.tune_chord import make, report  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
base = None
for box in (True, False):
    for env in ({}, {"LBM_L2_WINDOW": "1"}, {"LBM_REGULAR": "1"}):
        if "LBM_REGULAR" in env and not box:
            continue
        for k in ("LBM_L2_WINDOW", "LBM_REGULAR"):
            os.environ.pop(k, None)
        os.environ.update(env)
        eng = make(n, 4, box=box)
        eng.step(7, write_macro_every=0)
        if box and not env:
            base = eng.populations.clone()
        same = bool(torch.equal(base, eng.populations)) if (box and base is not None) else None
        ms = timed(lambda: eng.step(1, write_macro_every=0), 30, 5)
        report(f"{'box' if box else 'v60'}_{n}_vec4_{'+'.join(env) or 'baseline'}", eng, ms, {"same_as_baseline_after_7_steps": same})
        del eng; torch.cuda.empty_cache()
    base = None
