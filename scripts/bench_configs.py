#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs on ONE B200 (bench.py carries the headline config):

  v60_512            configs[2]  V60 512^3, compat=physical: bounce-back + Guo force (gravity*phase + pressure-gradient
                                 drive recomputed every step) + local-stress Smagorinsky + porous drag
  v60_512_particles  configs[3]  same + 1 M coffee particles, two-way trilinear coupling every step
  ref_224            configs[0]  the reference's default box, compat=reference (all legacy quirks), LBMSolver.step()
  tgv_256_les / tgv_256_macro    periodic 256^3 with LES / with rho,u written every step

Roofline bookkeeping: bytes = N_fluid*B_alg + N_solid*1 with B_alg = 165 (V60) or 152 (periodic).
Prints one JSON line per config.  Not a driver contract file -- evidence for DESIGN.md / profiles/.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import measured_peak  # noqa: E402
from pour_over_coffee_lbm_b200.config import LBMConfig  # noqa: E402
from pour_over_coffee_lbm_b200.engine import D3Q19Engine, ParticleState, particles_couple  # noqa: E402


def timed(fn, steps, warmup, min_seconds=0.4):
    """ms per call.  Warm up until the GPU has been busy for >= min_seconds (clock ramp from idle takes ~100 ms on
    B200: a 20 ms window reads 40 % slow), then time enough calls to cover >= min_seconds."""
    import time
    t0 = time.perf_counter(); n = 0
    while n < warmup or time.perf_counter() - t0 < min_seconds:
        fn(); n += 1
        if n % 8 == 0: torch.cuda.synchronize()
    torch.cuda.synchronize()
    per = max(1e-6, (time.perf_counter() - t0) / n)
    steps = max(steps, int(min_seconds / per) + 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def report(name, eng, ms, b_alg, extra=None):
    peak, _ = measured_peak()
    cells, fluid = eng.cells(), eng.fluid_cells()
    bytes_ = fluid * b_alg + (cells - fluid) * 1
    line = {"config": name, "grid": [eng.nx, eng.ny, eng.nz], "ms_per_step": ms, "MLUPS": cells / ms / 1e3,
            "MFLUPS": fluid / ms / 1e3, "fluid_fraction": fluid / cells, "bytes_alg_per_fluid_cell": b_alg,
            "achieved_GBs": bytes_ / ms / 1e6, "roofline_frac_of_measured": bytes_ / ms / 1e6 / peak, "peak_GBs": peak}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def v60_engine(n, strict, vec, compat="physical"):
    cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
    kw = dict(porous_darcy=0.37, porous_forch=0.9) if compat == "physical" else {}
    eng = D3Q19Engine(n, n, n, compat=compat, periodic=(False, False, False), walls=True, force=True, phase=True, les=True,
                      porous=True, strict=strict, vec=vec, config=cfg, gravity_lu=1e-5, **kw)
    eng.build_v60_geometry()
    z = torch.arange(n, device="cuda")[:, None, None]
    eng.phase.copy_(((z < int(0.6 * n)) & (eng.solid == 0)).float())
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    u = 1e-3 * torch.randn((3, n, n, n), device="cuda", generator=g)
    eng.init_equilibrium(rho=torch.ones((n, n, n), device="cuda"), u=u)
    return eng


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--fast", action="store_true")
    ap.add_argument("--vec", type=int, default=0)
    ap.add_argument("--only", default="")
    ap.add_argument("--n", type=int, default=512)
    args = ap.parse_args()
    strict = not args.fast
    want = lambda k: (not args.only) or k in args.only.split(",")

    if want("tgv_256_les") or want("tgv_256_macro"):
        for name, les, macro in (("tgv_256_les", True, 0), ("tgv_256_macro", False, 1)):
            if not want(name):
                continue
            eng = D3Q19Engine(256, 256, 256, compat="physical", les=les, strict=strict, vec=args.vec, tau=0.53)
            ms = timed(lambda: eng.step(1, write_macro_every=macro), args.steps, args.warmup)
            report(name, eng, ms, 152, {"note": "rho,u written every step (+16 B/cell not counted)" if macro else "local-stress LES"})
            del eng

    if want("v60_512") or want("v60_512_particles"):
        n = args.n
        eng = v60_engine(n, strict, args.vec)
        if want("v60_512"):
            ms_kernel = timed(lambda: eng.step(1, write_macro_every=0), args.steps, args.warmup)
            report(f"v60_{n}_step_kernel_only", eng, ms_kernel, 165)

            def full():
                eng.set_pressure_gradient_force(0.12, 1.0)      # = clear_body_force + add_pressure_gradient_force on fluid cells
                eng.step(1, write_macro_every=1)
            ms = timed(full, args.steps, args.warmup)
            report(f"v60_{n}", eng, ms, 165, {"note": "pressure-gradient drive (written, not accumulated: no clear pass) + fused step with rho,u write-out"})
        if want("v60_512_particles"):
            P = 1_000_000
            ps = ParticleState(P, eng.device)
            rng = np.random.default_rng(42)
            cfg = eng.cfg
            # uniform in the coffee-bed frustum (bottom 30 % of the cone)
            zb, zt = 5.0, 5.0 + 0.3 * cfg.CUP_HEIGHT / cfg.SCALE_LENGTH
            z = rng.uniform(zb + 1, zt, P)
            rr = (cfg.BOTTOM_RADIUS + (cfg.TOP_RADIUS - cfg.BOTTOM_RADIUS) * (z - zb) * cfg.SCALE_LENGTH / cfg.CUP_HEIGHT) / cfg.SCALE_LENGTH
            r = np.sqrt(rng.uniform(0, 1, P)) * 0.8 * rr
            th = rng.uniform(0, 2 * np.pi, P)
            pos = np.stack([n / 2 + r * np.cos(th), n / 2 + r * np.sin(th), z]).astype(np.float32)
            ps.pos.copy_(torch.from_numpy(pos))
            rad = np.clip(rng.normal(3.25e-4, 0.3 * 3.25e-4, P), 0.5 * 3.25e-4, 1.5 * 3.25e-4).astype(np.float32)
            ps.radius.copy_(torch.from_numpy(rad))
            ps.mass.copy_(torch.from_numpy(((np.float32(4 / 3) * np.float32(3.14159)) * rad ** 3 * np.float32(1200.0)).astype(np.float32)))
            ps.active.fill_(1)
            react = torch.zeros_like(eng.u)
            eng.step(1, write_macro_every=1)
            ms_p = timed(lambda: particles_couple(eng, ps, react, relax=0.8), args.steps, args.warmup)

            def coupled():
                # step_with_two_way_coupling (legacy/lbm_solver.py:1485-1509) + the drive.  body_force = drive + reaction:
                # the drive is WRITTEN first (no clear pass), the reaction added on top -- the same two-term f32 sum as
                # clear -> += reaction -> += drive
                eng.set_pressure_gradient_force(0.12, 1.0)
                particles_couple(eng, ps, react, relax=0.8)
                eng.add_reaction_force(react)
                eng.step(1, write_macro_every=1)
            ms = timed(coupled, args.steps, args.warmup)
            report(f"v60_{n}_particles_1M", eng, ms, 165,
                   {"particle_kernel_ms": ms_p, "particle_fraction_of_step": ms_p / ms, "particles": P,
                    "note": "step_with_two_way_coupling sequence: drive (written), couple (gather+drag+scatter+relax, incl. memset of reaction), add reaction, step"})
        del eng

    if want("ref_224"):
        n = 224
        eng = v60_engine(n, strict, args.vec, compat="reference")
        eng.phase.mul_(0.3)      # tau_air: the stable regime of the legacy solver (quirk Q1)
        ms = timed(lambda: eng.step(1, write_macro_every=1), args.steps, args.warmup)
        report("ref_224_compat_reference", eng, ms, 165, {"note": "LBMSolver.step(): FD-LES on lagged u + macroscopic + collide/stream + filter damping, one kernel"})


if __name__ == "__main__":
    main()
