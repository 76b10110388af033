"""Timing breakdown of the slab path under torchrun (diagnostic)."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pour_over_coffee_lbm_b200.engine import D3Q19Engine
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("N", 256))
def timeit(fn, steps=300):
    [fn() for _ in range(30)]; dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(steps) if fn.__code__.co_argcount else [fn() for _ in range(steps)]; e1.record()
    dist.barrier(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
eng = D3Q19Engine(n, n, n, compat="physical", device=local, zghost=1, z0=rank * n, nz_global=n * world, tau=0.53)
eng.attach_process_group()
t_full = timeit(lambda k=3: eng.step(k, write_macro_every=0))
t_x = timeit(lambda: eng.halo_exchange())
cs = eng.comm_stream; eng.comm_stream = None
t_noverlap = timeit(lambda k=3: eng.step(k, write_macro_every=0))
eng.comm_stream = cs
single = D3Q19Engine(n, n, n, compat="physical", device=local, tau=0.53)
t_single = timeit(lambda k=3: single.step(k, write_macro_every=0))
ghost1 = D3Q19Engine(n, n, n, compat="physical", device=local, zghost=1, z0=0, nz_global=n, tau=0.53)
t_ghost1 = timeit(lambda k=3: ghost1.step(k, write_macro_every=0))
if rank == 0:
    print(f"n={n} world={world}: overlap step {t_full:.4f} ms | exchange alone {t_x:.4f} ms | serial (no comm stream) {t_noverlap:.4f} ms | "
          f"single-GPU wrap {t_single:.4f} ms | single-GPU ghost planes + local memcpy wrap {t_ghost1:.4f} ms", flush=True)
dist.destroy_process_group()
