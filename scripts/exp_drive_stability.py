#!/usr/bin/env python
"""How long does the V60 state stay finite under the fused pressure-gradient drive, as a function of the drive's scale?"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bench import v60_engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for scale, maxf in ((1.0, 0.12), (0.5, 0.12), (0.1, 0.12), (0.05, 0.12), (0.02, 0.12), (1.0, 1e-4), (0.0, 0.12)):
    eng = v60_engine(n, drive=True, force=False)
    eng.set_params(drive_scale=scale, drive_max_force=maxf)
    hist = []
    for block in range(10):
        eng.step(200, write_macro_every=1)
        s = eng.field_statistics().tolist()
        hist.append((200 * (block + 1), round(s[0], 5), round(s[1], 4), round(s[2], 4), int(s[5] + s[6])))
        if s[5] + s[6] > 0:
            break
    print(json.dumps({"n": n, "drive_scale": scale, "max_force": maxf, "history(step,max|u|,min rho,max rho,nonfinite)": hist}), flush=True)
    del eng; torch.cuda.empty_cache()
