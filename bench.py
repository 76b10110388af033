#!/usr/bin/env python
"""bench.py -- MLUPS of the fused D3Q19 step (BASELINE.json metric) on N B200s of one box.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (C restatement of its Taichi kernels)

One "step" = one pass of the hot path over the whole lattice.

N = 1   headline = BASELINE.json configs[2], the largest single-GPU configuration: V60 geometry 512^3, compat = physical,
        halfway bounce-back + Guo force (gravity * phase + pressure-gradient drive) + local-stress Smagorinsky + porous drag,
        rho and u written every step -- ONE kernel launch per step (the drive is fused into the step kernel).  Sub-records
        (`configs`): the same box without the drive / without the rho,u write-out, the unfused producer + step sequence,
        configs[3] (+ 1 M two-way coupled particles), configs[1] (periodic 256^3 Taylor-Green, the roofline calibration) and
        configs[0] (the reference's own 224^3 box in compat = reference).
N > 1   BASELINE.json configs[4]: the V60 box at 1024^3 with 1 M two-way coupled particles, fluid-balanced z-slabs, the 5 + 5
        outgoing populations (+ rho for the fused drive) per interface over NCCL, overlapped with the interior; before the
        timing every rank checks its slab of a 64^3 V60 run against a single-GPU run of the whole box, bit for bit
        (`parity_ok`).  Sub-record: the periodic Taylor-Green ring (256^3 per rank, weak scaling) of round 1.

Every measurement is the median over blocks of exactly `--steps` steps, repeated until at least 0.5 s of GPU time has been
timed (CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import datetime
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0       # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
B_V60 = 165                     # algorithmic bytes per fluid-cell update: 19 x 4 x 2 populations + 1 flag + 12 force
B_PERIODIC = 152                # 19 x 4 B read + 19 x 4 B write
METRIC = "MLUPS (D3Q19 fused step)"
# PressureGradientDrive on compat = physical: F = -cs^2 grad(rho)/rho from the PREVIOUS step's rho doubles the lattice's own pressure
# force explicitly; at the reference's scale 1 the V60 state is non-finite within 200 steps, at 0.5 within 400, at <= 0.1 it is stable
# for thousands (profiles/r02_exp_drive_stability.log).  The benchmark runs the drive at 0.1: same kernel, same arithmetic, finite state
# (a non-finite state sends the packed reciprocal / square root down their scalar fall-backs and times those instead).
CHORD_END_COST = float(os.environ.get('LBM_BENCH_CHORD_END_COST', '16'))      # slab cuts by kernel cost, not by fluid count (engine.v60_fluid_cells_per_plane)
DRIVE_MAX_FORCE, DRIVE_SCALE = 0.12, 0.1


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def box_copy_bandwidth():
    """STREAM-style copy on THIS box, same recipe as MEASURED_PEAKS.json (b.copy_(a) over 1 Gi bf16 elements, read+write
    bytes, best of 10, CUDA events).  Reported next to the official peak because the leases of this pool differ."""
    import torch
    a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
    a.fill_(1.0)
    best = 0.0
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
        best = max(best, 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del a, b
    torch.cuda.empty_cache()
    return best


def profile_traffic(key):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/step_kernel_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as fh:
            return json.load(fh).get(key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi sampled while the GPU is under load (B200_PROFILING.md clocks line).  Started BEFORE the warm-up and the
    barrier -- round 1 forked it between the barrier and the first event and the other ranks' timed region swallowed rank 0's
    fork latency."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# timing
# ---------------------------------------------------------------------------------------------------
class Timer:
    """Blocks of exactly `steps` calls of the hot path, each bracketed by barrier + synchronize and timed with CUDA events on
    the current stream; the blocks repeat until `min_seconds` of device time is covered (the same count on every rank: it is
    agreed from rank 0's first block).  Per block the time is the MAX over ranks; the figure reported is the MEDIAN block."""

    def __init__(self, world: int, min_seconds: float = 0.5, max_blocks: int = 400):
        self.world, self.min_seconds, self.max_blocks = world, min_seconds, max_blocks

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def _block(self, run, steps):
        import torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(self, run, steps: int, warmup: int):
        """run(k) enqueues k steps.  Returns dict(ms_per_step, ms_min, ms_max, blocks, steps_per_block)."""
        import torch
        run(max(warmup, 3))
        # clock ramp: a B200 coming from idle needs ~100 ms under load before it boosts.  The number of calls is agreed between the
        # ranks (rank 0 times one call and broadcasts the count): `run` may contain collectives -- the particle coupling on slabs
        # all-reduces on every n-th call -- and a loop bounded by each rank's own wall clock lets the ranks drift apart in call count
        # until one of them waits in a collective the others never enter.
        self.barrier()
        t0 = time.perf_counter()
        run(max(1, steps)); torch.cuda.synchronize()
        reps = int(min(200, max(1, np.ceil(0.25 / max(time.perf_counter() - t0, 1e-4)))))
        if self.world > 1:
            import torch.distributed as dist
            nr = torch.tensor([reps], dtype=torch.int64, device="cuda")
            dist.broadcast(nr, src=0)
            reps = int(nr.item())
        for _ in range(reps):
            run(max(1, steps)); torch.cuda.synchronize()
        first = self._block(run, steps)
        blocks = int(min(self.max_blocks, max(3, np.ceil(self.min_seconds * 1e3 / max(first, 1e-3)))))
        if self.world > 1:
            import torch.distributed as dist
            nb = torch.tensor([blocks], dtype=torch.int64, device="cuda")
            dist.broadcast(nb, src=0)
            blocks = int(nb.item())
        times = [first] + [self._block(run, steps) for _ in range(blocks - 1)]
        per = sorted(t / steps for t in times)
        return {"ms_per_step": statistics.median(per), "ms_min": per[0], "ms_max": per[-1], "blocks": blocks, "steps_per_block": steps}


def record(name, cells, fluid, b_alg, tm, peak, note=None, launches_per_step=None, extra=None):
    ms = tm["ms_per_step"]
    bytes_ = fluid * b_alg + (cells - fluid) * (1 if b_alg == B_V60 else 0)
    r = {"config": name, "ms_per_step": ms, "ms_min": tm["ms_min"], "ms_max": tm["ms_max"], "blocks": tm["blocks"],
         "MLUPS": cells / ms / 1e3, "MFLUPS": fluid / ms / 1e3, "fluid_fraction": fluid / cells,
         "bytes_alg_per_fluid_cell": b_alg, "achieved_GBs": bytes_ / ms / 1e6, "roofline_frac": bytes_ / ms / 1e6 / peak}
    if launches_per_step is not None:
        r["launches_per_step"] = launches_per_step
    if note:
        r["note"] = note
    r.update(extra or {})
    return r


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
def v60_engine(n, *, nz_global=None, z0=0, nz=None, zghost=0, device=0, drive=False, force=True, compat="physical", seed=1234,
               vec=0, block=0):
    """BASELINE configs[2]: V60 mask of an n x n x nz_global box: phase = 1 inside the cone below 0.6 NZ,
    GRAVITY_LU = 1e-5, every feature on, velocity perturbation N(0, 1e-3) folded into f_eq."""
    import torch
    from pour_over_coffee_lbm_b200.config import LBMConfig
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine
    nzg = nz_global or n
    nz = nz or nzg
    cfg = LBMConfig(NX=n, NY=n, NZ=nzg, GRAVITY_LU=1e-5)
    kw = dict(porous_darcy=0.37, porous_forch=0.9) if compat == "physical" else {}
    eng = D3Q19Engine(n, n, nz, compat=compat, periodic=(False, False, False), walls=True, force=force, phase=True, les=True, porous=True,
                      config=cfg, gravity_lu=1e-5, zghost=zghost, z0=z0, nz_global=nzg, device=device, drive=drive, vec=vec, block=block,
                      drive_max_force=DRIVE_MAX_FORCE, drive_scale=DRIVE_SCALE, **kw)
    eng.build_v60_geometry()
    dev = eng.device
    zglob = torch.arange(z0 - zghost, z0 + nz + zghost, device=dev)[:, None, None]
    eng.phase.copy_(((zglob < int(0.6 * nzg)) & (eng.solid == 0)).float())
    g = torch.Generator(device=dev); g.manual_seed(seed + z0)
    shp = (nz + 2 * zghost, n, n)
    eng.init_equilibrium(rho=torch.ones(shp, device=dev), u=1e-3 * torch.randn((3,) + shp, device=dev, generator=g))
    return eng


def bed_particles(eng, count, seed=42):
    """`count` particles uniform in the coffee-bed frustum (bottom 30 % of the cone, 80 % of the local
    radius), radius N(3.25e-4, 30 %) clipped to [0.5, 1.5] x mean, at rest, stored in cell order (z slowest).  Same particles on every
    rank."""
    import torch
    from pour_over_coffee_lbm_b200.engine import ParticleState
    cfg = eng.cfg
    n = eng.nx
    rng = np.random.default_rng(seed)
    zb = 5.0; zt = zb + 0.3 * cfg.CUP_HEIGHT / cfg.SCALE_LENGTH
    z = rng.uniform(zb + 1, zt, count)
    rr = (cfg.BOTTOM_RADIUS + (cfg.TOP_RADIUS - cfg.BOTTOM_RADIUS) * (z - zb) * cfg.SCALE_LENGTH / cfg.CUP_HEIGHT) / cfg.SCALE_LENGTH
    r = np.sqrt(rng.uniform(0, 1, count)) * 0.8 * rr
    th = rng.uniform(0, 2 * np.pi, count)
    ps = ParticleState(count, eng.device)
    ps.pos.copy_(torch.from_numpy(np.stack([n / 2 + r * np.cos(th), n / 2 + r * np.sin(th), z]).astype(np.float32)))
    rad = np.clip(rng.normal(3.25e-4, 0.3 * 3.25e-4, count), 0.5 * 3.25e-4, 1.5 * 3.25e-4).astype(np.float32)
    ps.radius.copy_(torch.from_numpy(rad))
    ps.mass.copy_(torch.from_numpy(((np.float32(4 / 3) * np.float32(3.14159)) * rad ** 3 * np.float32(1200.0)).astype(np.float32)))
    ps.active.fill_(1)
    ps.sort_by_cell(eng.nx, eng.ny)     # layer by layer, as the reference creates its bed (coffee_particles.py:220-412)
    return ps


def tgv_fields(nx, ny, nz_total, z0, nz, u0=0.04):
    """3-D Taylor-Green initial state of the slab [z0, z0+nz) in device layout ([z,y,x], [c,z,y,x]).  u0 = 0.04 is SURVEY 8d's
    throughput state (the decay-rate gate of the tests runs the z-invariant vortex at u0 = 0.01)."""
    import torch
    k = 2.0 * np.pi / nx
    x = torch.arange(nx, dtype=torch.float64) * k
    y = torch.arange(ny, dtype=torch.float64) * (2.0 * np.pi / ny)
    z = (torch.arange(nz, dtype=torch.float64) + z0) * (2.0 * np.pi / nz_total)
    X = x[None, None, :]; Y = y[None, :, None]; Z = z[:, None, None]
    ux = u0 * torch.sin(X) * torch.cos(Y) * torch.cos(Z)
    uy = -u0 * torch.cos(X) * torch.sin(Y) * torch.cos(Z)
    uz = torch.zeros_like(ux)
    rho = 1.0 + (3.0 * u0 * u0 / 16.0) * (torch.cos(2 * X) + torch.cos(2 * Y)) * (torch.cos(2 * Z) + 2.0)
    return rho.float().contiguous(), torch.stack([ux, uy, uz]).float().contiguous()


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own algorithm (C restatement of its Taichi kernels), all host threads
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, n=256, budget_s=150.0):
    """LBMSolver.step() as the reference runs it on a V60 box -- FD-LES pre-pass, macroscopic, collide + push-stream + bounce-back,
    COPY-swap, filter damping, face BCs (oracle/ref_cpu.c, OpenMP) -- on the headline workload's state at n^3 (bounded sample:
    the GPU arm's box is 512^3; the reference's cost per cell does not depend on the box size).  The phase field is scaled to
    0.3 so that the legacy solver relaxes with tau_air: with tau_water it is linearly unstable and would time NaNs.
    Returns (mlups, ms_per_step, threads, sample description)."""
    from oracle import d3q19_ref as R, ref_cpu as RC
    threads = os.cpu_count() or 1
    try:      # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every host thread it can get, whatever the launcher set
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(threads)
    except Exception:
        pass
    nz = n

    def make(nz_):
        cfg = R.RefConfig(NX=n, NY=n, NZ=nz_, GRAVITY_LU=1e-5)
        st = R.init_fields(cfg)
        st.solid = RC.v60_solid(cfg)
        st.filter_zone = R.filter_zones(cfg)
        st.filter_blockage = np.zeros(st.rho.shape, np.float32)
        st.K_lu, st.beta_lu = R.forchheimer_params(cfg)
        st.les_mask = np.where(st.filter_zone == 1, 0, st.les_mask).astype(np.int32)
        st.apply_filter = True
        st.phase[:, :, : int(0.6 * nz_)] = 0.3
        st.phase[st.solid != 0] = 0.0
        rng = np.random.default_rng(1234)
        u = (1e-3 * rng.standard_normal((3,) + st.rho.shape)).astype(np.float32)
        one = np.ones(st.rho.shape, np.float32)
        for q in range(R.Q):
            st.f[q] = R.equilibrium_ref(one, u[0], u[1], u[2], q, "config")
            st.f_new[q] = st.f[q]
        return RC.CState(st)

    cs = make(nz)
    t0 = time.perf_counter(); cs.step(1); t1 = time.perf_counter() - t0
    if t1 * (steps + warmup) > budget_s and nz > 64:     # bounded sample: the lower part of the same box
        n_fit = max(64, int(nz * budget_s / (t1 * (steps + warmup))) // 8 * 8)
        nz = n_fit
        cs = make(nz)
    cs.step(max(1, warmup))
    t0 = time.perf_counter(); cs.step(steps); dt = time.perf_counter() - t0
    cells = n * n * nz
    fluid = int((cs.solid == 0).sum())
    sample = (f"{steps} calls of the C restatement of LBMSolver.step() (FD-LES + macroscopic + collide/push-stream/bounce-back + copy-swap + "
              f"filter damping + face BCs; Taichi not installable) on the V60 box at {n}x{n}x{nz} ({100 * fluid / cells:.1f} % fluid), f32, "
              f"OpenMP, {RC.num_threads()} threads")
    return cells * steps / dt / 1e6, dt / steps * 1e3, RC.num_threads(), sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mlups, ms, cores, sample = cpu_reference_run(args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "V60 geometry, D3Q19 + FD-LES + Guo-like force + filter damping + halfway bounce-back: LBMSolver.step() of the "
                               "reference (BASELINE configs[2] physics) on a bounded 256^3 sample of the box the GPU arm runs at 512^3 (N = 1) / "
                               "1024^3 (N > 1); MLUPS counts every lattice cell, as the reference does",
                   "threads": cores},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm, N = 1
# ---------------------------------------------------------------------------------------------------
def run_single(args, local):
    import torch
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine, particles_couple
    timer = Timer(1)
    peak, peak_src = measured_peak()
    n = args.size
    sampler = ClockSampler(local)
    subs = []
    want = (lambda k: True) if not args.only else (lambda k: k in args.only.split(","))

    # ---- headline: configs[2] sequence, drive fused into the step kernel, rho,u written every step ---------------
    eng = v60_engine(n, drive=True, force=False, device=local, vec=args.vec)
    cells, fluid = eng.cells(), eng.fluid_cells()
    l0 = eng.launch_count()
    head = timer.measure(lambda k: eng.step(k, write_macro_every=1), args.steps, args.warmup)
    l1 = eng.launch_count()
    launches_per_step = None
    eng.step(args.steps, write_macro_every=1); torch.cuda.synchronize()
    launches_per_step = (eng.launch_count() - l1) / args.steps
    headline = record(f"v60_{n}_sequence", cells, fluid, B_V60, head, peak, launches_per_step=launches_per_step,
                      note="configs[2]: pressure-gradient drive (fused into the step kernel, from the previous step's rho) + gravity*phase + LES + "
                           "porous drag + bounce-back, rho,u written every step; ONE launch per step")
    subs.append(headline)
    t0 = time.perf_counter()
    eng.step(args.steps, write_macro_every=1)
    host_enqueue_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    torch.cuda.synchronize()
    stats = eng.field_statistics().tolist()
    finite = stats[5] == 0 and stats[6] == 0
    del eng; torch.cuda.empty_cache()

    if want("v60_step"):
        eng = v60_engine(n, device=local, vec=args.vec)
        tm = timer.measure(lambda k: eng.step(k, write_macro_every=0), args.steps, args.warmup)
        subs.append(record(f"v60_{n}_step_only", cells, fluid, B_V60, tm, peak, launches_per_step=1,
                           note="the step kernel alone: body_force read from a field (no drive), no rho,u write-out"))
        tm = timer.measure(lambda k: eng.step(k, write_macro_every=1), args.steps, args.warmup)
        subs.append(record(f"v60_{n}_step_macro", cells, fluid, B_V60, tm, peak, launches_per_step=1, note="+ rho,u written every step (16 B per fluid cell, not in the denominator)"))

        def unfused(k):
            for _ in range(k):
                eng.set_pressure_gradient_force(DRIVE_MAX_FORCE, DRIVE_SCALE)
                eng.step(1, write_macro_every=1)
        tm = timer.measure(unfused, args.steps, args.warmup)
        subs.append(record(f"v60_{n}_sequence_unfused", cells, fluid, B_V60, tm, peak, launches_per_step=2,
                           note="round-1 form of the sequence: pressure-gradient producer (written into body_force) + step, two launches"))
        del eng; torch.cuda.empty_cache()

    if want("particles"):
        eng = v60_engine(n, drive=True, force=True, device=local, vec=args.vec)
        ps = bed_particles(eng, args.particles)
        eng.step(1, write_macro_every=1)

        def coupled(k):
            # LBMSolver.step_with_two_way_coupling (legacy/lbm_solver.py:1485-1509) + the drive: clear -> coupling on the current u ->
            # under-relaxation -> reaction into body_force -> step.  The coupling scatters straight into body_force (clear + "+= reaction"
            # in one; the clear is sparse: last step's deposits, cell by cell), the drive is added inside the step kernel.
            for _ in range(k):
                particles_couple(eng, ps, eng.body_force, relax=0.8, sparse_clear=True)
                eng.step(1, write_macro_every=1)
        tm = timer.measure(coupled, args.steps, args.warmup)
        tp = timer.measure(lambda k: [particles_couple(eng, ps, eng.body_force, relax=0.8, sparse_clear=True) for _ in range(k)], args.steps, args.warmup)
        subs.append(record(f"v60_{n}_particles_{args.particles}", cells, fluid, B_V60, tm, peak, launches_per_step=2,
                           note="configs[3]: + two-way coupled particles (trilinear gather, Schiller-Naumann drag, warp-aggregated atomic scatter, "
                                "under-relaxation) every step",
                           extra={"particles": args.particles, "particle_kernel_ms": tp["ms_per_step"]}))
        del eng, ps; torch.cuda.empty_cache()

    if want("tgv"):
        m = 256
        eng = D3Q19Engine(m, m, m, compat="physical", device=local, tau=0.53)
        rho0, u0 = tgv_fields(m, m, m, 0, m)
        eng.init_equilibrium(rho=rho0.cuda(), u=u0.cuda())
        tm = timer.measure(lambda k: eng.step(k, write_macro_every=0), max(args.steps, 50), args.warmup)
        subs.append(record("tgv_256", m ** 3, m ** 3, B_PERIODIC, tm, peak, launches_per_step=1,
                           note="configs[1], roofline calibration: periodic 256^3 Taylor-Green (u0 = 0.04), BGK, dense kernel, rho,u on demand"))
        tm = timer.measure(lambda k: eng.step(k, write_macro_every=1), max(args.steps, 50), args.warmup)
        subs.append(record("tgv_256_macro", m ** 3, m ** 3, B_PERIODIC, tm, peak, launches_per_step=1, note="+ rho,u written every step (16 B per cell, not in the denominator)"))
        del eng; torch.cuda.empty_cache()

    if want("ref224"):
        m = 224
        eng = v60_engine(m, device=local, compat="reference")
        eng.phase.mul_(0.3)      # tau_air: the stable regime of the legacy solver (quirk Q1)
        tm = timer.measure(lambda k: eng.step(k, write_macro_every=1), max(args.steps, 50), args.warmup)
        subs.append(record("ref_224", eng.cells(), eng.fluid_cells(), B_V60, tm, peak, launches_per_step=1,
                           note="configs[0]: the reference's default box, compat = reference (legacy arithmetic, bit-exact build): FD-LES on the lagged u + "
                                "macroscopic + collide/stream + filter damping in one kernel"))
        del eng; torch.cuda.empty_cache()

    clocks = sampler.stop()
    e2e = run_e2e_single(args, local, n) if want("e2e") else None
    box_gbs = box_copy_bandwidth()
    cpu = None
    if not args.no_cpu_baseline:
        mlups_cpu, ms_cpu, cores, sample = cpu_reference_run(args.cpu_steps, 2, budget_s=25.0)
        cpu = {"value": mlups_cpu, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample, "ms_per_step": ms_cpu}
    bytes_launch = fluid * B_V60 + (cells - fluid)
    line = {
        "metric": METRIC, "value": headline["MLUPS"], "unit": "MLUPS", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": headline["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[2]: V60 geometry {n}^3, D3Q19 + Smagorinsky LES + Guo force (gravity*phase + pressure-gradient drive) + "
                               "filter-paper porous drag + halfway bounce-back, rho,u written every step",
                   "grid": [n, n, n], "fluid_cells": fluid, "MFLUPS": headline["MFLUPS"], "compat": "physical",
                   "kernel": "phys_chord_kernel<FORCED,LES,POROUS,DRIVE>: packed list of active quads (32 per warp), cp.async-staged populations, packed f32x2 "
                             "collision (one copy in the instruction cache), halfway bounce-back through a per-link value buffer, whole-quad stores, "
                             "fused drive; explicitly rounded (one build, bit-exact vs the oracle)",
                   "cache": "working set 27 GB >> 126 MB L2 (inputs larger than L2, no flush needed)",
                   "timing": f"median of {head['blocks']} blocks of {args.steps} steps (min {head['ms_min']:.4f} / max {head['ms_max']:.4f} ms per step)",
                   "drive": {"max_force": DRIVE_MAX_FORCE, "scale": DRIVE_SCALE, "why": "scale 1 of the explicit lagged-density drive is unstable in compat = physical"},
                   "state_finite_after_run": bool(finite)},
        "roofline": {"bound": "hbm", "achieved": bytes_launch / headline["ms_per_step"] / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": headline["roofline_frac"], "traffic": profile_traffic(f"v60_{n}_sequence_dram_bytes_per_launch"), "peak_source": peak_src,
                     "copy_gbs_this_box": box_gbs, "algorithmic_bytes_per_launch": bytes_launch,
                     "basis": "165 B per fluid cell (19x4x2 populations + 1 flag + 12 force) + 1 B per solid cell; one launch = one step of the whole box"},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(round(launches_per_step * args.steps)), "host_enqueue_ms_per_step": host_enqueue_ms,
        "clocks": clocks, "configs": subs,
    }
    print(json.dumps(line), flush=True)


def run_e2e_single(args, local, n):
    """The same metric through the reference-facing API (LBMSolver.step() of the facade) with HOST buffers: every step copies the body_force
    field the orchestration rewrites each step (main.py:770-800) from pinned host memory -- double-buffered on a copy stream, so the copy of
    step k+1 overlaps step k -- steps, and reads the step statistics (max |u|, mean rho: main.py:907-912) back to the host."""
    import torch
    from pour_over_coffee_lbm_b200.config import LBMConfig
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    from pour_over_coffee_lbm_b200.physics import FilterPaperSystem
    cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
    solver = LBMSolver(nx=n, ny=n, nz=n, config=cfg, compat="physical", periodic=(False, False, False), geometry=True, device=local,
                       les=True, phase=True, force=True, gravity_lu=1e-5, porous_darcy=0.37, porous_forch=0.9)
    FilterPaperSystem(solver).initialize_filter_geometry()
    e = solver.engine
    z = torch.arange(n, device=e.device)[:, None, None]
    e.phase.copy_(((z < int(0.6 * n)) & (e.solid == 0)).float())
    g = torch.Generator(device=e.device); g.manual_seed(1234)
    e.init_equilibrium(rho=torch.ones((n, n, n), device=e.device), u=1e-3 * torch.randn((3, n, n, n), device=e.device, generator=g))
    host_force = torch.zeros(e.body_force.shape, dtype=torch.float32).pin_memory()
    dev_force = [e.body_force, torch.empty_like(e.body_force)]
    stats_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(e.device)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    steps = min(args.steps, 20)
    main = torch.cuda.current_stream(e.device)

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i & 1])                      # the step that read this buffer last has finished
            dev_force[i & 1].copy_(host_force, non_blocking=True)     # H2D of step i's input
            ready[i & 1].record(copy_stream)

    def one_step(i):
        main.wait_event(ready[i & 1])
        e.body_force = dev_force[i & 1]
        upload(i + 1)                                                 # next step's input travels while this step runs
        solver.step()
        freed[i & 1].record(main)
        stats = solver.step_statistics()                              # device reduction of the step's result
        stats_host.copy_(stats, non_blocking=True)                    # D2H
        main.synchronize()

    for ev in freed:
        ev.record(main)
    upload(0)
    for i in range(3):
        one_step(i)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(3, 3 + steps):
        one_step(i)
    ev1.record()
    torch.cuda.synchronize()
    ms_step = ev0.elapsed_time(ev1) / steps
    # the variant main.py actually runs: the producers live on the device, the host sends nothing and reads the statistics
    e.body_force = dev_force[0]
    solver2_ms = None
    try:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            e.set_pressure_gradient_force(DRIVE_MAX_FORCE, DRIVE_SCALE); solver.step(); stats_host.copy_(solver.step_statistics(), non_blocking=True); main.synchronize()
        ev0.record()
        for _ in range(steps):
            e.set_pressure_gradient_force(DRIVE_MAX_FORCE, DRIVE_SCALE)                  # PressureGradientDrive.apply() on the device
            solver.step()
            stats_host.copy_(solver.step_statistics(), non_blocking=True)
            main.synchronize()
        ev1.record(); torch.cuda.synchronize()
        solver2_ms = ev0.elapsed_time(ev1) / steps
    except Exception:
        solver2_ms = None
    out = {"value": n ** 3 / (ms_step * 1e-3) / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": int(host_force.numel() * 4),
           "d2h_bytes_per_step": int(stats_host.numel() * 4), "ms_per_step": ms_step, "steps": steps,
           "api": "LBMSolver.step() of the facade on the V60 box; body_force fed from pinned host memory every step (double-buffered on a copy "
                  "stream), step statistics read back",
           "device_producers": None if solver2_ms is None else
           {"value": n ** 3 / (solver2_ms * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": solver2_ms, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
            "api": "what main.py does: PressureGradientDrive on the device + LBMSolver.step() + statistics read back every step"}}
    del solver, e, dev_force
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------
# our arm, N > 1
# ---------------------------------------------------------------------------------------------------
def slab_parity_check(world, rank, local, n=64, steps=12):
    """Every rank's slab of a V60 n^3 run (all features, fused drive, fluid-balanced cuts) against a single-GPU run of the whole box
    on the same rank: populations of the owned fluid cells, rho, u -- bit for bit.  A wrong halo cannot print a fast MLUPS."""
    import torch
    import torch.distributed as dist
    from pour_over_coffee_lbm_b200 import slab
    from pour_over_coffee_lbm_b200.engine import v60_fluid_cells_per_plane
    from pour_over_coffee_lbm_b200.config import LBMConfig
    cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
    part = slab.partition_z_balanced(v60_fluid_cells_per_plane(cfg, local, chord_end_cost=CHORD_END_COST), world, min_planes=3)[rank]
    # one global random state, cut per rank: the per-slab generator of v60_engine would differ from the whole-box run
    g = torch.Generator(device="cuda"); g.manual_seed(777)
    u_all = 1e-3 * torch.randn((3, n, n, n), device="cuda", generator=g)
    rho_all = 1.0 + 1e-3 * torch.randn((n, n, n), device="cuda", generator=g)
    whole = v60_engine(n, device=local, drive=True, force=False)
    whole.init_equilibrium(rho=rho_all, u=u_all)
    whole.step(steps, write_macro_every=1)
    mine = v60_engine(n, nz_global=n, z0=part.z0, nz=part.nz, zghost=1, device=local, drive=True, force=False)
    mine.attach_process_group()
    pad = lambda t: torch.nn.functional.pad(t, (0, 0, 0, 0, 1, 1))
    zs = slice(part.z0, part.z0 + part.nz)
    rho_s, u_s = pad(rho_all[zs]), pad(u_all[:, zs])
    if part.z0 > 0: rho_s[0] = rho_all[part.z0 - 1]; u_s[:, 0] = u_all[:, part.z0 - 1]
    if part.z0 + part.nz < n: rho_s[-1] = rho_all[part.z0 + part.nz]; u_s[:, -1] = u_all[:, part.z0 + part.nz]
    mine.init_equilibrium(rho=rho_s, u=u_s)
    mine.step(steps, write_macro_every=1)
    torch.cuda.synchronize()
    fluid = whole.solid[zs] == 0
    ok = bool(torch.equal(mine.populations[:, 1:-1][:, fluid], whole.populations[:, zs][:, fluid]) and
              torch.equal(mine.rho[1:-1][fluid], whole.rho[zs][fluid]) and torch.equal(mine.u[:, 1:-1][:, fluid], whole.u[:, zs][:, fluid]))
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    del whole, mine
    torch.cuda.empty_cache()
    return bool(t.item() == 1)


def run_multi(args, world, rank, local):
    import torch
    import torch.distributed as dist
    from pour_over_coffee_lbm_b200 import slab
    from pour_over_coffee_lbm_b200.config import LBMConfig
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine, particles_couple_slab, v60_fluid_cells_per_plane
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=150))      # a mismatched collective aborts in minutes, not in ten
    timer = Timer(world)
    peak, peak_src = measured_peak()
    sampler = ClockSampler(local) if rank == 0 else None
    parity_ok = slab_parity_check(world, rank, local)

    # ---- headline: configs[4], V60 1024^3 + particles on fluid-balanced z-slabs ----------------------------------
    n = args.size_multi
    cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
    part = slab.partition_z_balanced(v60_fluid_cells_per_plane(cfg, local, chord_end_cost=CHORD_END_COST), world, min_planes=3)[rank]
    torch.cuda.empty_cache()
    eng = v60_engine(n, nz_global=n, z0=part.z0, nz=part.nz, zghost=1, device=local, drive=True, force=True, vec=args.vec)
    eng.attach_process_group()
    eng.halo_exchange()
    ps = bed_particles(eng, args.particles)
    fl = torch.tensor([float(eng.fluid_cells())], device="cuda", dtype=torch.float64)
    allf = [torch.zeros_like(fl) for _ in range(world)]
    dist.all_gather(allf, fl)
    per_fluid = [int(a.item()) for a in allf]
    fluid, cells = sum(per_fluid), n ** 3
    eng.step(1, write_macro_every=1)

    def coupled(k):
        for _ in range(k):
            particles_couple_slab(eng, ps, eng.body_force, relax=0.8, sparse_clear=True, interface_guard=True)
            eng.step(1, write_macro_every=1)
    l0 = eng.launch_count()
    head = timer.measure(coupled, args.steps, args.warmup)
    l1 = eng.launch_count()
    coupled(args.steps); torch.cuda.synchronize()
    launches_per_step = (eng.launch_count() - l1) / args.steps
    t0 = time.perf_counter(); coupled(args.steps); host_enqueue_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    torch.cuda.synchronize()
    fluid_only = timer.measure(lambda k: eng.step(k, write_macro_every=1), args.steps, args.warmup)
    stats = eng.field_statistics().tolist()
    finite = torch.tensor([1 if (stats[5] == 0 and stats[6] == 0) else 0], device="cuda")
    dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    del eng, ps; torch.cuda.empty_cache()

    # ---- sub-record: periodic Taylor-Green ring, 256^3 per rank (weak scaling, round 1's SCALE workload) ----------
    m = 256
    ring = D3Q19Engine(m, m, m, compat="physical", device=local, zghost=1, z0=rank * m, nz_global=m * world, tau=0.53)
    ring.attach_process_group()
    rho0, u0 = tgv_fields(m, m, m * world, rank * m, m)
    pad = lambda t: torch.nn.functional.pad(t, (0, 0, 0, 0, 1, 1))
    ring.init_equilibrium(rho=pad(rho0).cuda(), u=pad(u0).cuda())
    tg = timer.measure(lambda k: ring.step(k, write_macro_every=0), max(args.steps, 50), args.warmup)
    del ring; torch.cuda.empty_cache()
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        bytes_total = fluid * B_V60 + (cells - fluid)
        headline = record(f"v60_{n}_particles_{args.particles}_slabs", cells, fluid, B_V60, head, peak * world, launches_per_step=launches_per_step,
                          note="configs[4]: V60 box on fluid-balanced z-slabs, drive fused, 1 M replicated particles (owner computes), rho,u every step",
                          extra={"per_rank_fluid_Mcells": [round(x / 1e6, 2) for x in per_fluid], "particles": args.particles})
        subs = [headline,
                record(f"v60_{n}_sequence_slabs", cells, fluid, B_V60, fluid_only, peak * world, note="the same slabs without the particle coupling"),
                record("tgv_256_per_rank_ring", m ** 3 * world, m ** 3 * world, B_PERIODIC, tg, peak * world,
                       note="periodic Taylor-Green ring, 256^3 per rank (weak scaling), dense kernel, NCCL halo overlapped with the interior")]
        line = {
            "metric": METRIC, "value": headline["MLUPS"], "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": headline["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: V60 geometry {n}^3 + LES + drive + porous drag + bounce-back + {args.particles} two-way coupled particles, "
                                   f"{world} fluid-balanced z-slabs, NCCL halo of 5+5 populations (+ rho) per interface overlapped with the interior",
                       "grid": [n, n, n], "fluid_cells": fluid, "MFLUPS": headline["MFLUPS"], "compat": "physical",
                       "cache": "working set >> 126 MB L2 per GPU (inputs larger than L2, no flush needed)",
                       "timing": f"median of {head['blocks']} blocks of {args.steps} steps (min {head['ms_min']:.4f} / max {head['ms_max']:.4f} ms per step), max over ranks",
                       "parity_ok": parity_ok, "state_finite_after_run": bool(finite.item() == 1)},
            "parity_ok": parity_ok,
            "roofline": {"bound": "hbm", "achieved": bytes_total / headline["ms_per_step"] / 1e6 / world, "peak": peak, "unit": "GB/s",
                         "frac": bytes_total / headline["ms_per_step"] / 1e6 / world / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_total / world, "basis": "per GPU: 165 B per fluid cell + 1 B per solid cell of the rank's slab (mean over ranks)"},
            "cpu_baseline": None,
            "e2e": {"value": headline["MLUPS"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "ms_per_step": headline["ms_per_step"],
                    "api": "engine.step + particles_couple_slab on device-resident state: at N > 1 no per-step host input exists in this workload (the producers run "
                           "on the device); the host-buffer e2e number is the N = 1 line's"},
            "gpu_launches": int(round(launches_per_step * args.steps)), "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clocks, "configs": subs,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_ours(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        run_multi(args, world, rank, local)
    else:
        run_single(args, local)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vec", type=int, default=0, help="cells per thread (0 = library default)")
    ap.add_argument("--size", type=int, default=512, help="V60 box edge at N = 1")
    ap.add_argument("--size-multi", type=int, default=1024, help="V60 box edge at N > 1")
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only", default="", help="comma list of sub-records to run besides the headline: v60_step,particles,tgv,ref224,e2e")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
