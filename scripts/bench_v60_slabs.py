#!/usr/bin/env python
"""BASELINE configs[4] analogue: a V60 box of NX x NX x NZ (default 1024^3) with every feature, z-slab partitioned over
the ranks of one box (torchrun, one rank per GPU), NCCL halo exchange overlapped with the interior update.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/bench_v60_slabs.py --size 1024 --steps 30

Prints one JSON line on rank 0: whole-job MLUPS / MFLUPS (max over ranks of the device time), per-rank fluid cells and
per-rank step time without the exchange (load imbalance of equal-thickness slabs through a cone), roofline fraction.
"""
import argparse, json, os, sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pour_over_coffee_lbm_b200 import slab  # noqa: E402
from pour_over_coffee_lbm_b200.config import LBMConfig  # noqa: E402
from pour_over_coffee_lbm_b200.engine import D3Q19Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=1024); ap.add_argument("--size-z", dest="nz", type=int, default=0)
    ap.add_argument("--steps", type=int, default=30); ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--equal", action="store_true", help="equal-thickness slabs instead of fluid-balanced ones")
    ap.add_argument("--particles", type=int, default=0, help="BASELINE configs[4]: N coffee particles, two-way coupled every step (replicated "
                    "on every rank, owner computes; engine.particles_couple_slab -- written at the end of round 1, first GPU run is round 2's)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, nzg = args.n, (args.nz or args.n)
    cfg = LBMConfig(NX=n, NY=n, NZ=nzg, GRAVITY_LU=1e-5)
    if args.equal or world == 1:
        part = slab.partition_z(nzg, world)[rank]
    else:      # equal WORK: slabs cut at the prefix sums of the per-plane fluid count (every rank computes the same cuts)
        from pour_over_coffee_lbm_b200.engine import v60_fluid_cells_per_plane
        part = slab.partition_z_balanced(v60_fluid_cells_per_plane(cfg, local), world, min_planes=3)[rank]
        torch.cuda.empty_cache()
    zg = 1 if world > 1 else 0
    eng = D3Q19Engine(n, n, part.nz, compat="physical", periodic=(False, False, False), walls=True, force=True, phase=True, les=True,
                      porous=True, config=cfg, gravity_lu=1e-5, porous_darcy=0.37, porous_forch=0.9, zghost=zg, z0=part.z0,
                      nz_global=nzg, device=local)
    if world > 1:
        eng.attach_process_group()
    eng.build_v60_geometry()
    zglob = torch.arange(part.z0 - zg, part.z0 + part.nz + zg, device="cuda")[:, None, None]
    eng.phase.copy_(((zglob < int(0.6 * nzg)) & (eng.solid == 0)).float())
    g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
    shp = (part.nz + 2 * zg, n, n)
    eng.init_equilibrium(rho=torch.ones(shp, device="cuda"), u=1e-3 * torch.randn((3,) + shp, device="cuda", generator=g))
    if world > 1:
        eng.halo_exchange()
    fluid = eng.fluid_cells()

    ps = react = None
    if args.particles > 0:
        from pour_over_coffee_lbm_b200.engine import ParticleState, particles_couple, particles_couple_slab
        ps = ParticleState(args.particles, eng.device)
        gp = torch.Generator(device="cuda"); gp.manual_seed(42)                      # the same particles on every rank
        ps.pos[0].uniform_(0.35 * n, 0.65 * n, generator=gp); ps.pos[1].uniform_(0.35 * n, 0.65 * n, generator=gp)
        ps.pos[2].uniform_(6.0, 0.45 * nzg, generator=gp)
        ps.radius.fill_(3.25e-4); ps.mass.fill_(float(4.0 / 3.0 * 3.14159 * 3.25e-4 ** 3 * 1200.0)); ps.active.fill_(1)
        react = torch.zeros_like(eng.u)

    def coupled_step():
        eng.clear_body_force()
        if world > 1: particles_couple_slab(eng, ps, react, relax=0.8)
        else: particles_couple(eng, ps, react, relax=0.8)
        eng.add_reaction_force(react)
        eng.step(1, write_macro_every=1)                                             # the coupling reads this step's u

    def timed(steps):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if ps is None:
            eng.step(steps, write_macro_every=0)
        else:
            for _ in range(steps): coupled_step()
        e1.record()
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        return e0.elapsed_time(e1) / steps

    timed(args.warmup)
    ms = timed(args.steps)
    t = torch.tensor([ms, float(fluid)], device="cuda", dtype=torch.float64)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
    else:
        allt = [t]
    if rank == 0:
        per_ms = [float(a[0]) for a in allt]; per_fluid = [int(a[1]) for a in allt]
        ms_max = max(per_ms)
        cells = n * n * nzg; tot_fluid = sum(per_fluid)
        peak = 6540.8
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        bytes_ = tot_fluid * 165 + (cells - tot_fluid)
        print(json.dumps({"config": f"V60 {n}x{n}x{nzg}, all features, z-slabs" + (f", {args.particles} particles two-way coupled" if args.particles else ""),
                          "n_gpus": world, "ms_per_step": ms_max,
                          "MLUPS": cells / ms_max / 1e3, "MFLUPS": tot_fluid / ms_max / 1e3, "fluid_fraction": tot_fluid / cells,
                          "partition": "equal thickness" if (args.equal or world == 1) else "fluid-balanced", "per_rank_ms": [round(x, 4) for x in per_ms], "per_rank_fluid_Mcells": [round(x / 1e6, 2) for x in per_fluid],
                          "slab_imbalance_max_over_mean": max(per_fluid) / (tot_fluid / world),
                          "achieved_GBs_total": bytes_ / ms_max / 1e6, "roofline_frac_of_measured_per_gpu": bytes_ / ms_max / 1e6 / peak / world}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
