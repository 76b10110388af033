#!/usr/bin/env python
"""One short device run for the two-cells-per-thread legacy-compatible kernel (compat = reference, vec = 2: the legacy arithmetic on
packed f32x2, collide_reference_t<P2>): the recorded runs of the reference's own LBMSolver.step() through tests/test_gpu_vs_reference_run.py
with vec = 2 (three step scenarios, the open box, every 1000-step recording), the newest 1000-step recording on the one- / four-cell kernels as
well, and BASELINE configs[0] (224^3 V60 box) timed on vec = 1 against vec = 2."""
import glob
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import test_gpu_vs_reference_run as T  # noqa: E402

if __name__ == "__main__":
    t0 = time.time()
    ok = True

    def run(name, fn, *a):
        global ok
        try:
            fn(*a); print(f"PASS {name}", flush=True)
        except BaseException:
            ok = False
            print(f"FAIL {name}", flush=True); traceback.print_exc(limit=3)

    for p in T.STEP_FILES:
        run(f"step {os.path.basename(p)[19:-4]} vec=2", T.test_step_kernel_reproduces_the_reference_run, p, 2)
    run("open box vec=2", T.test_open_box_step_and_face_bc_reproduce_the_reference_run, 2)
    for p in T.LONG_FILES:
        for vec in ((2,) if p.endswith("long_air_1000.npz") else (2, 1, 4)):
            run(f"1000 steps {os.path.basename(p)[19:-4]} vec={vec}", T.test_step_kernel_reproduces_1000_steps_of_the_reference_run, p, vec)
    print(f"parity: {'ALL PASS' if ok else 'FAILURES'} ({time.time() - t0:.1f} s)", flush=True)
    timer = bench.Timer(1, min_seconds=0.25)
    ref = None
    for vec in (1, 2):
        eng = bench.v60_engine(224, compat="reference", vec=vec)
        eng.phase.mul_(0.3)
        eng.step(7, write_macro_every=1); torch.cuda.synchronize()
        fluid = eng.solid == 0
        cur = (eng.rho[fluid].clone(), eng.u[:, fluid].clone())
        same = "" if ref is None else f"  rho {'==' if torch.equal(cur[0], ref[0]) else '!='} u {'==' if torch.equal(cur[1], ref[1]) else '!='} (vs vec=1, 7 steps)"
        ref = ref or cur
        tm = timer.measure(lambda k: eng.step(k, write_macro_every=1), 50, 5)
        fl = int(fluid.sum().item())
        frac = (fl * 165 + (224 ** 3 - fl)) / tm["ms_per_step"] / 1e6 / 6540.8
        print(f"ref_224 vec={vec}: {tm['ms_per_step']:.4f} ms/step (min {tm['ms_min']:.4f}) = {frac:.3f} of the measured HBM peak{same}", flush=True)
        del eng; torch.cuda.empty_cache()
    print(f"total {time.time() - t0:.1f} s")
