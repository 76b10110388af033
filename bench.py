#!/usr/bin/env python
"""bench.py -- MLUPS of the fused D3Q19 step (BASELINE.json metric) on N B200s of one box.

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (C restatement of its Taichi kernels)

One "step" = one pass of the hot path (LBMSolver.step / execute_collision_streaming) over the
whole lattice = ONE launch of the fused pull collide-stream kernel per GPU.
Workload at N=1: BASELINE.json configs[1], periodic 256^3 Taylor-Green vortex, D3Q19 BGK fp32.
At N>1 (weak scaling) every rank owns a 256x256x256 z-slab of a periodic 256x256x(256N) box and
exchanges the 5+5 outgoing populations per interface over NCCL, overlapped with the interior.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SIDE = 256
BYTES_PER_CELL = 152            # 19 x 4 B read + 19 x 4 B write (periodic BGK)
FALLBACK_HBM_GBS = 6650.0       # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


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


def profile_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as fh:
            return json.load(fh).get("tgv256_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
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


def tgv_fields(nx, ny, nz_total, z0, nz, u0=0.04):
    """3-D Taylor-Green initial state of the slab [z0, z0+nz) in device layout ([z,y,x], [c,z,y,x])."""
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
# CPU baseline: the reference's own algorithm (C restatement of its Taichi kernels), all host threads
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, n=N_SIDE, budget_s=150.0):
    """Times LBMSolver.step() structure (LES pre-pass off for BGK, macroscopic, collide+push, COPY-swap,
    boundary manager) on the TGV state.  Returns (mlups, ms_per_step, cores, sample description)."""
    from oracle import d3q19_ref as R, ref_cpu as RC
    nz = n

    def make(nz_):
        cfg = R.RefConfig(NX=n, NY=n, NZ=nz_, GRAVITY_LU=0.0, USE_LES=False)
        st = R.init_fields(cfg)
        k = 2 * np.pi / n
        i = np.arange(n)[:, None, None] * k; j = np.arange(n)[None, :, None] * k; kk = np.arange(nz_)[None, None, :] * (2 * np.pi / nz_)
        u0 = 0.04
        ux = (u0 * np.sin(i) * np.cos(j) * np.cos(kk)).astype(np.float32)
        uy = (-u0 * np.cos(i) * np.sin(j) * np.cos(kk)).astype(np.float32)
        uz = np.zeros_like(ux); rho = np.ones_like(ux)
        for q in range(R.Q):
            st.f[q] = R.equilibrium_ref(rho, ux, uy, uz, q, "config")
            st.f_new[q] = st.f[q]
        return RC.CState(st)

    cs = make(nz)
    t0 = time.perf_counter(); cs.step(1); t1 = time.perf_counter() - t0
    if t1 * (steps + warmup) > budget_s and nz > 32:     # bounded sample: thinner slab of the same box
        nz = max(32, int(nz * budget_s / (t1 * (steps + warmup))) // 8 * 8)
        cs = make(nz)
    cs.step(max(1, warmup))
    t0 = time.perf_counter(); cs.step(steps); dt = time.perf_counter() - t0
    cells = n * n * nz
    sample = (f"{steps} calls of the C restatement of LBMSolver.step() (macroscopic + collide/push-stream + copy-swap + "
              f"boundary manager; Taichi not installable) on a {n}x{n}x{nz} box, TGV state, f32, OpenMP")
    return cells * steps / dt / 1e6, dt / steps * 1e3, RC.num_threads(), sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mlups, ms, cores, sample = cpu_reference_run(args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "MLUPS (D3Q19 fused step)", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "periodic 256^3 Taylor-Green vortex, D3Q19 BGK fp32 (BASELINE configs[1]); reference has no periodic BC: "
                               "open faces, same lattice size and state"},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = N_SIDE
    nz_total = n * world
    if world > 1:
        eng = D3Q19Engine(n, n, n, compat="physical", device=local, zghost=1, z0=rank * n, nz_global=nz_total, tau=0.53,
                          vec=args.vec, strict=not args.fast)
        eng.attach_process_group()
    else:
        eng = D3Q19Engine(n, n, n, compat="physical", device=local, tau=0.53, vec=args.vec, strict=not args.fast)
    rho0, u0 = tgv_fields(n, n, nz_total, rank * n, n)
    if eng.zghost:
        pad = lambda t: torch.nn.functional.pad(t, (0, 0, 0, 0, 1, 1))
        rho0, u0 = pad(rho0), pad(u0)
    eng.init_equilibrium(rho=rho0.cuda(), u=u0.cuda())
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ------------------------------------------------
    eng.step(max(args.warmup, 3), write_macro_every=0)
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t_host0 = time.perf_counter()
    eng.step(args.steps, write_macro_every=0)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    cells_total = n * n * n * world
    mlups = cells_total / (ms_step * 1e-3) / 1e6

    # ---- end-to-end through the public API with HOST buffers ("e2e") --------------------------
    e2e = run_e2e(args, eng if world == 1 else None, world, rank, local)

    box_gbs = box_copy_bandwidth() if rank == 0 else None
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = BYTES_PER_CELL * n * n * n / (ms_step * 1e-3) / 1e9          # per GPU, per launch
        mlups_cpu, ms_cpu, cores, sample = (None, None, None, None)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            mlups_cpu, ms_cpu, cores, sample = cpu_reference_run(args.cpu_steps, 2, budget_s=25.0)
            cpu = {"value": mlups_cpu, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": "MLUPS (D3Q19 fused step)", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "periodic 256^3 Taylor-Green vortex, D3Q19 BGK fp32 (BASELINE configs[1])" +
                                   (f"; {world} z-slabs of 256^3, NCCL halo of 5+5 populations/interface overlapped with interior" if world > 1 else ""),
                       "grid_per_gpu": [n, n, n], "compat": "physical", "kernel": f"step_kernel<dense> VEC={args.vec or 4}, packed f32x2 collision, explicitly rounded (one build, bit-exact vs oracle)",
                       "cache": "working set 2.55 GB per GPU >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "macro_writeout": "rho,u materialised on demand, not inside the timed steps"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profile_traffic(), "peak_source": peak_src,
                         "copy_gbs_this_box": box_gbs, "frac_of_this_box_copy": achieved / box_gbs if box_gbs else None,
                         "algorithmic_bytes_per_launch": BYTES_PER_CELL * n * n * n},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, eng_unused, world, rank, local):
    """Same metric through the reference-facing API (LBMSolver.step) with host buffers: every step copies the
    body_force field the orchestration rewrites each step (main.py:770-800) from pinned host memory, steps, and
    reads the per-step statistics (max|u|, mean rho: main.py:907-912) back to the host."""
    import torch
    import torch.distributed as dist
    from pour_over_coffee_lbm_b200.solver import LBMSolver
    n = N_SIDE
    if world > 1:
        solver = LBMSolver(nx=n, ny=n, nz=n, compat="physical", periodic=(True, True, True), geometry=False, device=local,
                           les=False, phase=False, zghost=1, z0=rank * n, nz_global=n * world, tau=0.53)
        solver.engine.attach_process_group()
    else:
        solver = LBMSolver(nx=n, ny=n, nz=n, compat="physical", periodic=(True, True, True), geometry=False, device=local,
                           les=False, phase=False, tau=0.53)
    rho0, u0 = tgv_fields(n, n, n * world, rank * n, n)
    if solver.engine.zghost:
        pad = lambda t: torch.nn.functional.pad(t, (0, 0, 0, 0, 1, 1))
        rho0, u0 = pad(rho0), pad(u0)
    solver.engine.init_equilibrium(rho=rho0.cuda(), u=u0.cuda())
    host_force = torch.zeros(solver.engine.body_force.shape, dtype=torch.float32).pin_memory()
    stats_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    steps = min(args.steps, 50)

    def one_step():
        solver.engine.body_force.copy_(host_force, non_blocking=True)      # H2D of this step's input
        solver.step()
        stats = solver.step_statistics()                                    # device reduction of the step's result
        stats_host.copy_(stats, non_blocking=True)                          # D2H
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        one_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        one_step()
    ev1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    return {"value": n * n * n * world / (ms_step * 1e-3) / 1e6, "unit": "MLUPS",
            "h2d_bytes_per_step": int(host_force.numel() * 4), "d2h_bytes_per_step": int(stats_host.numel() * 4),
            "ms_per_step": ms_step, "steps": steps,
            "api": "LBMSolver.step() with body_force fed from pinned host memory each step, statistics read back"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vec", type=int, default=0, help="cells per thread (0 = library default)")
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fast", action="store_true", help="use the FMA-contracted build instead of the default bit-exact (-fmad=false) one")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
