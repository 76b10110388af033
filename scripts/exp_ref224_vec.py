#!/usr/bin/env python
"""BASELINE configs[0] (224^3 V60 box, compat = reference, strict build): cells per thread and CTA size of the legacy step
kernel behind walls.  Every variant must leave the SAME populations after the timed steps (bit for bit)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    m = 224
    timer = bench.Timer(1, min_seconds=0.3)
    ref = None
    for vec, block in ((1, 0), (1, 128), (2, 0), (4, 0), (4, 128), (1, 0)):
        try:
            eng = bench.v60_engine(m, compat="reference", vec=vec, block=block)
        except Exception as e:
            print(f"vec={vec} block={block}: {type(e).__name__}: {str(e)[:120]}"); continue
        eng.phase.mul_(0.3)
        eng.step(7, write_macro_every=1); torch.cuda.synchronize()
        g = eng.g[eng.cur].clone() if hasattr(eng, "cur") else None
        fluid = (eng.solid == 0)
        rho = eng.rho.clone(); u = eng.u.clone()
        if ref is None:
            ref = (rho, u)
            same = "reference variant"
        else:
            same = f"rho {'==' if torch.equal(rho[fluid], ref[0][fluid]) else '!='} u {'==' if torch.equal(u[:, fluid], ref[1][:, fluid]) else '!='}"
        tm = timer.measure(lambda k: eng.step(k, write_macro_every=1), 50, 5)
        print(f"ref_224 vec={vec} block={block or 'default'}: {tm['ms_per_step']:.4f} ms/step (min {tm['ms_min']:.4f} max {tm['ms_max']:.4f})  {same}", flush=True)
        del eng; torch.cuda.empty_cache()
