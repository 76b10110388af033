"""Runs a few full-feature walls steps (V60 mask, or an all-fluid box with --box) for an ncu capture of phys_walls_kernel."""
import argparse, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pour_over_coffee_lbm_b200.config import LBMConfig
from pour_over_coffee_lbm_b200.engine import D3Q19Engine
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=512); ap.add_argument("--box", action="store_true")
ap.add_argument("--vec", type=int, default=0); ap.add_argument("--block", type=int, default=0); ap.add_argument("--steps", type=int, default=8)
args = ap.parse_args()
n = args.n
cfg = LBMConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=1e-5)
eng = D3Q19Engine(n, n, n, compat="physical", periodic=(False, False, False), walls=True, force=True, phase=True, les=True,
                  porous=True, vec=args.vec, block=args.block, config=cfg, gravity_lu=1e-5, porous_darcy=0.37, porous_forch=0.9)
if args.box:
    eng.solid.zero_(); eng.solid[0] = 1; eng.solid[-1] = 1; eng.solid[:, 0] = 1; eng.solid[:, -1] = 1; eng.solid[:, :, 0] = 1; eng.solid[:, :, -1] = 1
    eng.filter_zone.zero_(); eng.filter_zone[n // 2] = 1; eng.pack_flags()
else:
    eng.build_v60_geometry()
z = torch.arange(n, device="cuda")[:, None, None]
eng.phase.copy_(((z < int(0.6 * n)) & (eng.solid == 0)).float())
g = torch.Generator(device="cuda"); g.manual_seed(1234)
eng.init_equilibrium(rho=torch.ones((n, n, n), device="cuda"), u=1e-3 * torch.randn((3, n, n, n), device="cuda", generator=g))
eng.step(args.steps, write_macro_every=0)
torch.cuda.synchronize()
