#!/usr/bin/env python
"""Round-2 experiment on the one-tile-per-warp chord kernel: L2 prefetch distance (LBM_PREFETCH = tiles ahead, 0 = off)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed  # noqa: E402
from scripts.tune_chord import make, report  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
block = int(sys.argv[2]) if len(sys.argv) > 2 else 128
for box in (False, True):
    base = None
    for dist in (0, 256, 1024, 4096, 16384):
        os.environ["LBM_PREFETCH"] = str(dist)
        eng = make(n, 4, block, box=box)
        eng.step(7, write_macro_every=0)
        if base is None:
            base = eng.populations.clone()
        same = bool(torch.equal(base, eng.populations))
        ms = timed(lambda: eng.step(1, write_macro_every=0), 30, 5)
        report(f"{'box' if box else 'v60'}_{n}_vec4_block{block}_prefetch{dist}", eng, ms, {"same_as_prefetch0": same})
        del eng; torch.cuda.empty_cache()
