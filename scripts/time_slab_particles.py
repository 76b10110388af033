#!/usr/bin/env python
"""Per-phase device time of engine.particles_couple_slab on N ranks (torchrun): where do the milliseconds go?"""
import json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bench import v60_engine, bed_particles  # noqa: E402
from pour_over_coffee_lbm_b200 import slab  # noqa: E402
from pour_over_coffee_lbm_b200.config import LBMConfig  # noqa: E402
from pour_over_coffee_lbm_b200.engine import particles_couple, particles_couple_slab, v60_fluid_cells_per_plane  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
part = slab.partition_z_balanced(v60_fluid_cells_per_plane(cfg, local), world, min_planes=3)[rank]
eng = v60_engine(n, nz_global=n, z0=part.z0, nz=part.nz, zghost=1, device=local, drive=True, force=True)
eng.attach_process_group(); eng.halo_exchange()
ps = bed_particles(eng, 1_000_000)
eng.step(2)
react = eng.body_force
per_z = False


def timed(fn, reps=20):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 4)


owned = slab.particle_owner_mask(ps.pos[2], ps.active, eng.z0, eng.nz, eng.nz_global)
outs = [ps.drag_new, ps.u_fluid, ps.reynolds, ps.cd, ps.cell, ps.drag, ps.drag_old]
res = {
    "n": n, "world": world,
    "exchange_planes(u)": timed(lambda: slab.exchange_planes(eng.u, eng.rank, eng.nranks, per_z)),
    "exchange_field(u) [NCCL in the library]": timed(lambda: eng.exchange_field(vec3=eng.u)),
    "owner_mask": timed(lambda: slab.particle_owner_mask(ps.pos[2], ps.active, eng.z0, eng.nz, eng.nz_global)),
    "couple kernel (sparse clear)": timed(lambda: particles_couple(eng, ps, react, relax=0.8, sparse_clear=True)),
    "reduce_ghost_up": timed(lambda: slab.reduce_ghost_up(react, eng.rank, eng.nranks, per_z)),
    "allreduce_owned_packed": timed(lambda: slab.allreduce_owned_packed(outs, owned, ps.active)),
    "allreduce_owned_packed(drag only)": timed(lambda: slab.allreduce_owned_packed([ps.drag], owned, ps.active)),
    "whole particles_couple_slab (sync=all)": timed(lambda: particles_couple_slab(eng, ps, react, relax=0.8, sparse_clear=True)),
    "whole particles_couple_slab (sync=state)": timed(lambda: particles_couple_slab(eng, ps, react, relax=0.8, sparse_clear=True, sync="state")),
    "step(1)": timed(lambda: eng.step(1, write_macro_every=1)),
    "step(20)/20": round(timed(lambda: eng.step(20, write_macro_every=1), reps=3) / 20, 4),
}
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
