#!/usr/bin/env python
"""lbm_field_statistics on the V60 512^3 box: time per call and the eight values against torch reductions over the same fields."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    eng = bench.v60_engine(n)
    eng.step(5, write_macro_every=1)
    torch.cuda.synchronize()
    got = eng.field_statistics().clone().cpu().double()
    fluid = eng.solid == 0
    r = eng.rho[fluid].double(); u = eng.u[:, fluid]
    um = torch.sqrt((u[0] * u[0] + u[1] * u[1]) + u[2] * u[2])
    want = torch.tensor([um.max().item(), r.min().item(), r.max().item(), r.sum().item(),
                         (0.5 * r * (u.double() ** 2).sum(0)).sum().item(), 0.0, 0.0, float(fluid.sum().item())], dtype=torch.float64)
    print("device :", got.tolist())
    print("torch  :", want.tolist())
    exact = all(got[i] == want[i] for i in (0, 1, 2, 5, 6, 7))
    close = all(abs(got[i] - want[i]) <= 1e-10 * abs(want[i]) for i in (3, 4))
    print("max / min / counts exact:", exact, "| sums within 1e-10:", close)
    timer = bench.Timer(1, min_seconds=0.2)
    tm = timer.measure(lambda k: [eng.field_statistics() for _ in range(k)], 20, 3)
    fl = int(fluid.sum().item())
    gb = (n ** 3 + 16 * fl) / 1e9
    print(f"lbm_field_statistics V60 {n}^3: {tm['ms_per_step']:.4f} ms per call (min {tm['ms_min']:.4f}), {gb:.3f} GB algorithmic = {gb / tm['ms_per_step'] * 1e3:.0f} GB/s")
    # the stand-alone pressure-gradient producer over the quad list (set mode: 4 B rho + 12 B force per fluid cell + flags)
    tm = timer.measure(lambda k: [eng.set_pressure_gradient_force(0.12, 0.1) for _ in range(k)], 20, 3)
    gb = (n ** 3 + 16 * fl) / 1e9
    print(f"lbm_pressure_gradient_force_set V60 {n}^3: {tm['ms_per_step']:.4f} ms per call (min {tm['ms_min']:.4f}), {gb:.3f} GB algorithmic = {gb / tm['ms_per_step'] * 1e3:.0f} GB/s")
    tm = timer.measure(lambda k: [eng.add_pressure_gradient_force(0.12, 0.1) for _ in range(k)], 20, 3)
    gb = (n ** 3 + 28 * fl) / 1e9
    print(f"lbm_pressure_gradient_force (accumulate) V60 {n}^3: {tm['ms_per_step']:.4f} ms per call (min {tm['ms_min']:.4f}), {gb:.3f} GB algorithmic = {gb / tm['ms_per_step'] * 1e3:.0f} GB/s")
    # the two reference-mode producers that were still cell-at-a-time: Forchheimer force (filter shell only) and the reaction add
    tm = timer.measure(lambda k: [eng.add_forchheimer_force() for _ in range(k)], 20, 3)
    print(f"lbm_forchheimer_force V60 {n}^3: {tm['ms_per_step']:.4f} ms per call (min {tm['ms_min']:.4f})")
    reaction = torch.zeros_like(eng.body_force)
    tm = timer.measure(lambda k: [eng.add_reaction_force(reaction) for _ in range(k)], 20, 3)
    gb = (n ** 3 + 36 * fl) / 1e9
    print(f"lbm_add_reaction_force V60 {n}^3: {tm['ms_per_step']:.4f} ms per call (min {tm['ms_min']:.4f}), {gb:.3f} GB algorithmic = {gb / tm['ms_per_step'] * 1e3:.0f} GB/s")
