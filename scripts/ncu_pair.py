"""Runs a dense and a walls (all-fluid box) no-feature VEC=1 step for ncu comparison."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pour_over_coffee_lbm_b200.engine import D3Q19Engine
n = 256
vec = int(os.environ.get("VEC", "1"))
a = D3Q19Engine(n, n, n, compat="physical", vec=vec, tau=0.6)
b = D3Q19Engine(n, n, n, compat="physical", periodic=(False,) * 3, walls=True, vec=vec, tau=0.6)
b.solid.zero_(); b.solid[0] = 1; b.solid[-1] = 1; b.solid[:, 0] = 1; b.solid[:, -1] = 1; b.solid[:, :, 0] = 1; b.solid[:, :, -1] = 1
b.pack_flags()
for e in (a, b):
    e.init_equilibrium(1.0, (0.01, 0.0, 0.0))
    e.step(10, write_macro_every=0)
torch.cuda.synchronize()
