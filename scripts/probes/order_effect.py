"""Is the V60 step time a property of the kernel or of where the buffers landed / what ran before?  Times the same case several times
in one process, with other allocations in between, and prints the addresses of the population buffers."""
import json, os, sys, time, subprocess, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from scripts.tune_chord import make
from scripts.bench_configs import timed
def clk():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    except Exception as e:
        return str(e)
def one(tag, n=512, hold=None):
    eng = make(n, 4)
    ms = timed(lambda: eng.step(1, write_macro_every=0), 30, 5)
    print(json.dumps({"tag": tag, "ms": round(ms, 4), "g0": hex(eng.g[0].data_ptr()), "g1": hex(eng.g[1].data_ptr()), "bf": hex(eng.body_force.data_ptr()),
                      "flags": hex(eng.flags.data_ptr()) if hasattr(eng, "flags") else None, "smi": clk()}), flush=True)
    del eng; torch.cuda.empty_cache()
one("first")
one("second")
junk = [torch.empty(int(37e6) + 4096 * i, dtype=torch.uint8, device="cuda") for i in range(5)]
one("after_small_allocs_held")
del junk; torch.cuda.empty_cache()
small = [make(96, 4) for _ in range(3)]
for s in small: s.step(3)
del small; torch.cuda.empty_cache()
one("after_small_engines")
time.sleep(5)
one("after_5s_idle")
