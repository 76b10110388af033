"""A few steps of the V60 (or box) case on the four-cell quad-list kernel for an ncu capture: --drive = the bench headline
(fused drive, rho/u written), default = the step kernel alone (force from a field, no rho/u write-out)."""
import argparse, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.tune_chord import make
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=512); ap.add_argument("--box", action="store_true")
ap.add_argument("--drive", action="store_true"); ap.add_argument("--steps", type=int, default=6)
args = ap.parse_args()
eng = make(args.n, 4, box=args.box, drive=args.drive, force=not args.drive)
eng.step(args.steps, write_macro_every=1 if args.drive else 0)
torch.cuda.synchronize()
