#!/usr/bin/env python
"""Round-2 tuning / cross-check of the chord-fitted four-cell walls kernel (csrc/lbm_phys_chord.cuh) on one B200.

  check   V60 n^3 (default 128): vec = 4 (chord tiles + wall links) against vec = 2 (round-1 kernel, bit-exact against the oracle in the
          GPU tests) after `--check-steps` steps -- populations on fluid cells, rho, u bit for bit; and the fused pressure-gradient
          drive (LBM_FEAT_DRIVE) against pressure-gradient producer + step, with and without a body-force field.
  time    V60 n^3 (default 512), all features: vec 2 (64-thread CTAs) / vec 4 at 16 warps (64- and 128-thread CTAs) / vec 4 at 12
          warps per SM; the same on an all-fluid box with solid faces; the drive sequence unfused and fused.
Prints one JSON line per measurement.  Evidence for DESIGN.md / profiles/, not a driver contract file.
"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed  # noqa: E402
from bench import measured_peak  # noqa: E402
from pour_over_coffee_lbm_b200.config import LBMConfig  # noqa: E402
from pour_over_coffee_lbm_b200.engine import D3Q19Engine  # noqa: E402


def make(n, vec, block=0, box=False, drive=False, force=True, seed=1234, periodic=(False, False, False)):
    cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
    eng = D3Q19Engine(n, n, n, compat="physical", periodic=periodic, walls=True, force=force, phase=True, les=True, porous=True,
                      vec=vec, block=block, config=cfg, gravity_lu=1e-5, porous_darcy=0.37, porous_forch=0.9, drive=drive)
    if box == "periodic":
        eng.filter_zone[n // 2] = 1; eng.pack_flags()
    elif box:
        eng.solid.zero_(); eng.solid[0] = 1; eng.solid[-1] = 1; eng.solid[:, 0] = 1; eng.solid[:, -1] = 1; eng.solid[:, :, 0] = 1; eng.solid[:, :, -1] = 1
        eng.filter_zone.zero_(); eng.filter_zone[n // 2] = 1; eng.pack_flags()
    else:
        eng.build_v60_geometry()
    z = torch.arange(n, device="cuda")[:, None, None]
    eng.phase.copy_(((z < int(0.6 * n)) & (eng.solid == 0)).float())
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    rho = 1.0 + 1e-3 * torch.randn((n, n, n), device="cuda", generator=g)
    eng.init_equilibrium(rho=rho, u=1e-3 * torch.randn((3, n, n, n), device="cuda", generator=g))
    if force:
        eng.body_force.copy_(1e-6 * torch.randn((3, n, n, n), device="cuda", generator=g))
    return eng


def same(a, b, fluid=None):
    if fluid is not None:
        a, b = a[..., fluid], b[..., fluid]
    return bool(torch.equal(a, b))


def check(n, steps):
    out = {}
    ref = make(n, 2); ref.step(steps)
    fluid = ref.solid == 0
    for block in (0, 128, 256):
        e4 = make(n, 4, block); e4.step(steps)
        out[f"vec4_block{block}_equals_vec2"] = same(e4.populations, ref.populations, fluid) and same(e4.rho, ref.rho, fluid) and same(e4.u, ref.u, fluid)
        del e4
    for box in (True, "periodic"):
        per = (True, True, True) if box == "periodic" else (False, False, False)
        r2 = make(n, 2, box=box, periodic=per); r2.step(steps)
        e4 = make(n, 4, box=box, periodic=per); e4.step(steps)
        fl = r2.solid == 0
        out[f"box_{box}_vec4_equals_vec2"] = same(e4.populations, r2.populations, fl) and same(e4.rho, r2.rho, fl) and same(e4.u, r2.u, fl)
        del r2, e4
    # fused drive against producer + step (both on the four-cell kernel), with a body-force field (accumulate) and without (set)
    for with_bf in (True, False):
        a = make(n, 4)
        base = a.body_force.clone()
        for _ in range(steps):
            if with_bf:
                a.body_force.copy_(base); a.add_pressure_gradient_force(0.12, 1.0)
            else:
                a.set_pressure_gradient_force(0.12, 1.0)
            a.step(1)
        b = make(n, 4, drive=True, force=with_bf)
        if not with_bf:
            pass
        b.step(steps)
        out[f"fused_drive_equals_producer_plus_step_bodyforce_{with_bf}"] = same(a.populations, b.populations, fluid) and same(a.rho, b.rho, fluid) and same(a.u, b.u, fluid)
        del a, b
    del ref
    print(json.dumps({"check": f"V60 {n}^3, {steps} steps", **out}), flush=True)
    return all(out.values())


def report(name, eng, ms, extra=None):
    peak, _ = measured_peak()
    n3 = eng.cells(); fluid = eng.fluid_cells()
    b = fluid * 165 + (n3 - fluid)
    line = {"case": name, "ms": round(ms, 4), "MLUPS": round(n3 / ms / 1e3), "MFLUPS": round(fluid / ms / 1e3), "frac_of_measured_peak": round(b / ms / 1e6 / peak, 4)}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512); ap.add_argument("--check-n", type=int, default=128); ap.add_argument("--check-steps", type=int, default=12)
    ap.add_argument("--steps", type=int, default=30); ap.add_argument("--skip-check", action="store_true"); ap.add_argument("--skip-time", action="store_true")
    ap.add_argument("--only-default", action="store_true")
    args = ap.parse_args()
    ok = True
    if not args.skip_check:
        ok = check(args.check_n, args.check_steps)
    if not args.skip_time:
        n = args.n
        cases = [(2, 0), (4, 0), (4, 128), (4, 256)] if not args.only_default else [(0, 0)]
        for box in (False, True):
            for vec, block in cases:
                eng = make(n, vec, block, box=box)
                ms = timed(lambda: eng.step(1, write_macro_every=0), args.steps, 5)
                report(f"{'box' if box else 'v60'}_{n}_step_only_vec{vec}_block{block}", eng, ms)
                if not box:
                    ms = timed(lambda: eng.step(1, write_macro_every=1), args.steps, 5)
                    report(f"v60_{n}_step_macro_vec{vec}_block{block}", eng, ms)
                del eng; torch.cuda.empty_cache()
        for vec in ((2, 4) if not args.only_default else (0,)):
            eng = make(n, vec)

            def seq():
                eng.set_pressure_gradient_force(0.12, 1.0)
                eng.step(1, write_macro_every=1)
            ms = timed(seq, args.steps, 5)
            report(f"v60_{n}_sequence_unfused_vec{vec}", eng, ms, {"note": "pressure-gradient producer (set) + step with rho,u write-out"})
            del eng; torch.cuda.empty_cache()
        eng = make(n, 4, drive=True, force=False)
        ms = timed(lambda: eng.step(1, write_macro_every=1), args.steps, 5)
        report(f"v60_{n}_sequence_fused_drive", eng, ms, {"note": "drive evaluated inside the step kernel from the previous rho; rho,u written"})
        del eng; torch.cuda.empty_cache()
        eng = make(n, 4, drive=True, force=True)
        ms = timed(lambda: eng.step(1, write_macro_every=1), args.steps, 5)
        report(f"v60_{n}_sequence_fused_drive_plus_bodyforce", eng, ms, {"note": "drive fused + a body_force field (the particles' reaction) read"})
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
