#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/check_slabs.py

Every rank owns a z-slab (+1 ghost plane per side), steps with the NCCL halo exchange overlapped with the interior
(lbm_step with a comm stream), and rank 0 compares the gathered result BIT FOR BIT with a single-GPU run of the whole
box.  Cases: periodic box with LES (physical), V60 box with every feature (physical), legacy solver (reference, FD-LES
needs the u ghost planes too).
"""
import datetime
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from pour_over_coffee_lbm_b200 import slab  # noqa: E402
from pour_over_coffee_lbm_b200.config import LBMConfig  # noqa: E402
from pour_over_coffee_lbm_b200.engine import D3Q19Engine  # noqa: E402


def pad_z(t):
    return torch.nn.functional.pad(t, (0, 0, 0, 0, 1, 1))


def run_case(name, rank, world, local, nx, nzg, steps, make_engine, setup):
    part = slab.partition_z(nzg, world)[rank]
    eng = make_engine(nz=part.nz, zghost=1, z0=part.z0, nz_global=nzg, device=local)
    eng.attach_process_group()
    setup(eng, part.z0, part.nz, True)
    eng.halo_exchange(with_u=True)
    eng.step(steps)
    torch.cuda.synchronize()
    mine = (eng.populations[:, 1:-1].cpu(), eng.rho[1:-1].cpu(), eng.u[:, 1:-1].cpu())
    gathered = [None] * world
    dist.all_gather_object(gathered, (part.z0, mine))
    ok = True
    if rank == 0:
        gathered.sort(key=lambda t: t[0])
        g = torch.cat([m[0] for _, m in gathered], dim=1)
        rho = torch.cat([m[1] for _, m in gathered], dim=0)
        u = torch.cat([m[2] for _, m in gathered], dim=1)
        ref = make_engine(nz=nzg, zghost=0, z0=0, nz_global=nzg, device=local)
        setup(ref, 0, nzg, False)
        ref.step(steps)
        torch.cuda.synchronize()
        fluid = (ref.solid == 0).cpu() if ref.solid is not None else torch.ones_like(rho, dtype=torch.bool)
        e_g = torch.equal(g[:, fluid], ref.populations.cpu()[:, fluid])
        e_r = torch.equal(rho[fluid], ref.rho.cpu()[fluid])
        e_u = torch.equal(u[:, fluid], ref.u.cpu()[:, fluid])
        ok = e_g and e_r and e_u
        print(f"[check_slabs] {name}: world={world} box={nx}x{nx}x{nzg} steps={steps} populations={e_g} rho={e_r} u={e_u}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


def run_particles(rank, world, local, nx, nzg, make_engine, setup, cfg, n_part=20000, steps=6, guard=False, z_hi=None):
    """Two-way coupling on slabs (replicated particles, owner computes, engine.particles_couple_slab over NCCL) against the single-GPU
    coupling: base-cell indices, fluid velocity at the particle and Reynolds number bit for bit; drag within 1e-6 (powf); the reaction
    field and the flow after `steps` coupled steps within 1e-5 of the field scale (a GPU's scatter atomics are unordered)."""
    from pour_over_coffee_lbm_b200.engine import ParticleState, particles_couple, particles_couple_slab

    def particles(dev):
        rng = np.random.default_rng(11)
        ps = ParticleState(n_part, dev)
        pos = np.stack([rng.uniform(0.3 * nx, 0.7 * nx, n_part), rng.uniform(0.3 * nx, 0.7 * nx, n_part),
                        rng.uniform(4.0, (nzg - 4.0) if z_hi is None else z_hi, n_part)])
        ps.pos.copy_(torch.from_numpy(pos.astype(np.float32)))
        ps.vel.copy_(torch.from_numpy((1e-3 * rng.standard_normal((3, n_part))).astype(np.float32)))
        rad = np.clip(rng.normal(3.25e-4, 1e-4, n_part), 1.6e-4, 4.9e-4).astype(np.float32)
        ps.radius.copy_(torch.from_numpy(rad))
        ps.mass.copy_(torch.from_numpy(((np.float32(4 / 3) * np.float32(3.14159)) * rad ** 3 * np.float32(1200.0)).astype(np.float32)))
        ps.active.fill_(1); ps.active[::17] = 0
        return ps

    part = slab.partition_z(nzg, world)[rank]
    eng = make_engine(nz=part.nz, zghost=1, z0=part.z0, nz_global=nzg, device=local)
    eng.attach_process_group()
    setup(eng, part.z0, part.nz, True)
    eng.halo_exchange(with_u=True)
    ps = particles(eng.device)
    eng.body_force.zero_()          # sparse clear: the reaction target starts at zero and only the coupling writes it
    eng.step(3)
    for _ in range(steps):
        particles_couple_slab(eng, ps, eng.body_force, relax=0.8, sparse_clear=True, interface_guard=guard)
        eng.step(1)
    torch.cuda.synchronize()
    mine = (eng.rho[1:-1].cpu(), eng.u[:, 1:-1].cpu(), eng.body_force[:, 1:-1].cpu())
    gathered = [None] * world
    dist.all_gather_object(gathered, (part.z0, mine))
    ok = True
    if rank == 0:
        gathered.sort(key=lambda t: t[0])
        rho = torch.cat([m[0] for _, m in gathered], dim=0); u = torch.cat([m[1] for _, m in gathered], dim=1)
        react = torch.cat([m[2] for _, m in gathered], dim=1)
        ref = make_engine(nz=nzg, zghost=0, z0=0, nz_global=nzg, device=local)
        setup(ref, 0, nzg, False)
        pr = particles(ref.device)
        ref.body_force.zero_()
        ref.step(3)
        for _ in range(steps):
            particles_couple(ref, pr, ref.body_force, relax=0.8)
            ref.step(1)
        torch.cuda.synchronize()
        act = (pr.active != 0).cpu()
        e_cell = torch.equal(ps.cell.cpu()[:, act], pr.cell.cpu()[:, act])
        fluid = (ref.solid == 0).cpu()
        close = lambda a, b, tol: bool(((a - b).abs().max() <= tol * max(float(b.abs().max()), 1e-30)).item())
        e_uf = close(ps.u_fluid.cpu()[:, act], pr.u_fluid.cpu()[:, act], 1e-5)
        e_drag = close(ps.drag.cpu()[:, act], pr.drag.cpu()[:, act], 1e-5)
        e_react = close(react[:, fluid], ref.body_force.cpu()[:, fluid], 1e-5)
        e_rho = close(rho[fluid], ref.rho.cpu()[fluid], 1e-5)
        e_u = close(u[:, fluid], ref.u.cpu()[:, fluid], 1e-5)
        ok = e_cell and e_uf and e_drag and e_react and e_rho and e_u
        print(f"[check_slabs] particles on slabs{' (interface guard, bed below the first cut)' if guard else ''}: world={world} particles={n_part} coupled steps={steps} cell indices (bit-exact)={e_cell} u_fluid={e_uf} "
              f"drag={e_drag} reaction field={e_react} rho={e_rho} u={e_u}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


def run_particle_producers(rank, world, local, nx, nzg, make_engine, setup, cfg, n_part=20000):
    """The two per-particle producers next to the coupling on slabs -- CoffeeParticleSystem.apply_fluid_forces (owner of the base cell
    computes) and FilterPaperSystem.block_particles_at_filter (owner of the first filter plane in gz - 2 .. gz + 2 computes, the range
    may straddle an interface) -- against the single-GPU kernels: force, velocities, the active flags (incl. particles the kernel
    deactivates), the error counter and the accumulated-particles field, all bit for bit."""
    import ctypes as C
    from pour_over_coffee_lbm_b200.engine import ParticleState, particles_fluid_forces_slab, particles_block_at_filter_slab, _ptr
    sl = float(np.float32(cfg.SCALE_LENGTH))

    def particles(dev, scaled):
        rng = np.random.default_rng(13)
        ps = ParticleState(n_part, dev)
        pos = np.stack([rng.uniform(0.2 * nx, 0.8 * nx, n_part), rng.uniform(0.2 * nx, 0.8 * nx, n_part), rng.uniform(1.0, nzg - 2.0, n_part)])
        if scaled:
            pos = pos * sl                      # quirk Q9: the filter lookup divides the position by SCALE_LENGTH
        else:
            pos[0, ::97] = np.nan; pos[2, 5::101] = 1e9     # invalid coordinates: deactivated by the kernel, counted
        ps.pos.copy_(torch.from_numpy(pos.astype(np.float32)))
        vel = (1e-2 * rng.standard_normal((3, n_part))).astype(np.float32)
        ps.vel.copy_(torch.from_numpy(vel))
        rad = np.clip(rng.normal(3.25e-4, 1e-4, n_part), 1.6e-4, 4.9e-4).astype(np.float32)
        ps.radius.copy_(torch.from_numpy(rad))
        ps.mass.copy_(torch.from_numpy(((np.float32(4 / 3) * np.float32(3.14159)) * rad ** 3 * np.float32(1200.0)).astype(np.float32)))
        ps.active.fill_(1); ps.active[::19] = 0
        return ps

    rho_w, mu_w, grav = 997.0, 1.0e-3, 9.81
    part = slab.partition_z(nzg, world)[rank]
    eng = make_engine(nz=part.nz, zghost=1, z0=part.z0, nz_global=nzg, device=local)
    eng.attach_process_group()
    setup(eng, part.z0, part.nz, True)
    eng.halo_exchange(with_u=True)
    eng.step(3)
    pa, pb = particles(eng.device, False), particles(eng.device, True)
    force = torch.zeros_like(pa.pos); counters = torch.zeros(2, dtype=torch.int32, device=eng.device)
    particles_fluid_forces_slab(eng, pa, force, counters, rho_w, mu_w, grav)
    acc = torch.zeros_like(eng.rho)
    particles_block_at_filter_slab(eng, pb, acc, sl, 0.01, 7)
    torch.cuda.synchronize()
    gathered = [None] * world
    dist.all_gather_object(gathered, (part.z0, acc[1:-1].cpu()))
    ok = True
    if rank == 0:
        gathered.sort(key=lambda t: t[0])
        acc_all = torch.cat([m for _, m in gathered], dim=0)
        ref = make_engine(nz=nzg, zghost=0, z0=0, nz_global=nzg, device=local)
        setup(ref, 0, nzg, False)
        ref.step(3)
        ra, rb = particles(ref.device, False), particles(ref.device, True)
        rforce = torch.zeros_like(ra.pos); rcount = torch.zeros(2, dtype=torch.int32, device=ref.device)
        st = ra.struct()
        ref._check(ref.lib.lbm_particles_fluid_forces(ref._ctx, _ptr(ref.u), C.byref(st), _ptr(rforce), rho_w, mu_w, grav, _ptr(rcount), ref.stream), "ff")
        racc = torch.zeros_like(ref.rho)
        st = rb.struct()
        ref._check(ref.lib.lbm_particles_block_at_filter(ref._ctx, C.byref(st), _ptr(ref.flags), _ptr(racc), sl, 0.01, 7, ref.stream), "bf")
        torch.cuda.synchronize()
        same = lambda a, b: torch.equal(a.view(torch.int32), b.view(torch.int32)) if a.dtype == torch.float32 else torch.equal(a, b)
        act = ra.active != 0
        e_force = same(force[:, act], rforce[:, act]); e_vel = same(pa.vel, ra.vel); e_act = same(pa.active, ra.active); e_cnt = same(counters, rcount)
        e_bvel = same(pb.vel, rb.vel); e_acc = same(acc_all, racc.cpu())
        hits = int(round(float(racc.sum()) / 0.01)); deact = int(rcount[0])
        ok = e_force and e_vel and e_act and e_cnt and e_bvel and e_acc and hits > 0 and deact > 0
        print(f"[check_slabs] particle producers on slabs: world={world} particles={n_part} fluid forces: force={e_force} vel={e_vel} active={e_act} "
              f"counters={e_cnt} (deactivated {deact}); filter interception: vel={e_bvel} accumulated={e_acc} ({hits} bounces)", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=150))      # a mismatched collective aborts in minutes, not in ten
    nx = 64
    nzg = 32 * world
    ok = True

    # ---- periodic + LES (physical) -------------------------------------------------------------------
    u0 = H.smooth_velocity(nx, 0.04, 21, nz=nzg); rho0 = H.smooth_density(nx, 0.01, 21, nz=nzg)
    ru = torch.from_numpy(H.to_dev_scalar(rho0)); uu = torch.from_numpy(H.to_dev_vec(u0))

    def mk(nz, zghost, z0, nz_global, device):
        return D3Q19Engine(nx, nx, nz, compat="physical", les=True, tau=0.6, zghost=zghost, z0=z0, nz_global=nz_global, device=device)

    def setup(eng, z0, nz, ghost):
        r = ru[z0:z0 + nz]; v = uu[:, z0:z0 + nz]
        if ghost: r, v = pad_z(r), pad_z(v)
        eng.init_equilibrium(rho=r.cuda(), u=v.cuda())
    ok &= run_case("periodic+LES physical", rank, world, local, nx, nzg, 25, mk, setup)

    # ---- V60, all features (physical) ------------------------------------------------------------------
    cfg = LBMConfig(NX=nx, NY=nx, NZ=nzg, GRAVITY_LU=1e-5)
    bf = torch.from_numpy(H.to_dev_vec((2e-5 * np.random.default_rng(3).standard_normal((nx, nx, nzg, 3))).astype(np.float32)))

    def mk2(nz, zghost, z0, nz_global, device, compat="physical", drive=False):
        kw = dict(porous_darcy=0.37, porous_forch=0.9) if compat == "physical" else {}
        return D3Q19Engine(nx, nx, nz, compat=compat, periodic=(False, False, False), walls=True, force=True, phase=True, les=True,
                           porous=True, config=cfg, gravity_lu=1e-5, zghost=zghost, z0=z0, nz_global=nz_global, device=device, drive=drive,
                           drive_scale=0.5, **kw)

    def setup2(eng, z0, nz, ghost, phase_scale=1.0):
        eng.build_v60_geometry()
        zg = eng.zghost
        zglob = torch.arange(z0 - zg, z0 + nz + zg, device="cuda")[:, None, None]
        eng.phase.copy_(((zglob < int(0.6 * nzg)) & (eng.solid == 0)).float() * phase_scale)
        b = bf[:, z0:z0 + nz]; r = ru[z0:z0 + nz]; v = uu[:, z0:z0 + nz] * 0.5
        if ghost: b, r, v = pad_z(b), pad_z(r), pad_z(v)
        eng.body_force.copy_(b.cuda())
        eng.init_equilibrium(rho=r.cuda(), u=v.cuda())
    ok &= run_case("V60 all features physical", rank, world, local, nx, nzg, 25, mk2, setup2)
    # the pressure-gradient drive fused into the step kernel reads rho across the interface (rho planes travel with the halo)
    ok &= run_case("V60 physical + fused pressure-gradient drive", rank, world, local, nx, nzg, 25, lambda **k: mk2(drive=True, **k), setup2)
    ok &= run_particles(rank, world, local, nx, nzg, mk2, setup2, cfg)
    # a bed that stays planes away from every interface: the guarded calls skip the exchanges and must give the same answer
    ok &= run_particles(rank, world, local, nx, nzg, mk2, setup2, cfg, guard=True, z_hi=0.3 * (nzg / world), steps=8)

    ok &= run_particle_producers(rank, world, local, nx, nzg, mk2, setup2, cfg)

    # ---- legacy solver (reference): FD-LES reads u across the interface ----------------------------------
    ok &= run_case("V60 compat=reference (water phase, FD-LES active)", rank, world, local, nx, nzg, 10,
                   lambda **k: mk2(compat="reference", **k), setup2)
    ok &= run_case("V60 compat=reference (air phase)", rank, world, local, nx, nzg, 40,
                   lambda **k: mk2(compat="reference", **k), lambda e, z0, nz, g: setup2(e, z0, nz, g, 0.3))
    if rank == 0:
        print("[check_slabs] ALL OK" if ok else "[check_slabs] FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
