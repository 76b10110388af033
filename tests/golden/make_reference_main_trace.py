#!/usr/bin/env python
"""Records what the reference's `main.py` does to the solver and the physics modules (run in the authoring container).

The reference's `CoffeeSimulation` (main.py:445-935) is constructed and stepped UNMODIFIED under the Taichi stand-in
(tests/golden/taichi_shim) on an n^3 grid; every call it makes on LBMSolver, MultiphaseFlow3D, PrecisePouringSystem,
FilterPaperSystem, PressureGradientDrive, CoffeeParticleSystem and the BoundaryConditionManager is logged IN ORDER with its
arguments (scalars verbatim, objects and fields by role), together with
  * the attribute probes (`hasattr` / `getattr(..., None)`) main.py uses to pick its code path, and what they found,
  * which fields the reference's visualisation / diagnostics modules read (`to_numpy`), per module,
  * rho, u, phase, solid after the constructor and after every step (the recorded run itself).
`UnifiedLBMSolver` is replaced by the legacy `LBMSolver` exactly as main.py's own fallback does (main.py:551-561): the
reference's UnifiedLBMSolver + CUDABackend dies in the first pre-stabilisation step (cuda_backend.py:224 indexes a Python list
with a tuple), so the legacy solver IS the path main.py can run.

Output: tests/golden/reference_main_trace.json (+ reference_main_trace_fields.npz).  tests/test_main_trace.py binds every
recorded call against the facade's signatures on the CPU and replays the whole trace on the device.

    python tests/golden/make_reference_main_trace.py [n=16] [steps=12] [pressure_mode=none] [out.json]

Deterministic (NumPy / random seeded before the constructor): a second run reproduced the committed JSON and .npz byte for byte.
About 5 minutes of CPU per recording at 16^3 (the Taichi stand-in executes the reference's kernels as Python loops).
"""
import functools
import inspect
import json
import os
import random
import re
import sys
import time
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_reference_goldens import REF, load_reference, quiet  # noqa: E402

MAIN = os.path.join(REF, "main.py")
TRACE = []                 # ordered calls made from main.py
READERS = {}               # "module" -> sorted set of "field.op"
ROLES = {}                 # id(obj) -> role
FIELD_NAMES = {}           # id(field) -> "role.attr"
SIM = []
KEEP = {}                  # id -> object: keeps wrapped objects alive so that ids stay unique


class StopRecording(BaseException):
    pass


def caller_file(depth=2):
    return sys._getframe(depth).f_code.co_filename


def register(obj, role):
    ROLES[id(obj)] = role; KEEP[id(obj)] = obj
    for k, v in list(vars(obj).items()):
        if hasattr(v, "to_numpy"):
            if id(v) not in FIELD_NAMES:
                FIELD_NAMES[id(v)] = f"{role}.{k}"; KEEP[id(v)] = v
        elif isinstance(v, list) and v and hasattr(v[0], "to_numpy"):
            for i, w in enumerate(v):
                if id(w) not in FIELD_NAMES:
                    FIELD_NAMES[id(w)] = f"{role}.{k}[{i}]"; KEEP[id(w)] = w


def rescan():
    """Fields created after __init__ (lazily allocated ones): look through the registered objects again."""
    for i, role in list(ROLES.items()):
        register(KEEP[i], role)


def enc(v):
    if v is None or isinstance(v, (bool, int, str)):
        return v
    if isinstance(v, float):
        return v
    if isinstance(v, (np.floating, np.integer, np.bool_)):
        return v.item()
    if isinstance(v, dict):
        return {"dict": {str(k): enc(w) for k, w in v.items()}}
    if isinstance(v, (list, tuple)):
        return {"list": [enc(w) for w in v]}
    if id(v) in ROLES:
        return {"obj": ROLES[id(v)]}
    inner = getattr(v, "_solver", None)                    # MinimalAdapter
    if inner is not None and id(inner) in ROLES:
        return {"obj": ROLES[id(inner)]}
    if hasattr(v, "to_numpy"):
        if id(v) not in FIELD_NAMES:
            rescan()
        if id(v) in FIELD_NAMES:
            return {"field": FIELD_NAMES[id(v)]}
    if isinstance(v, np.ndarray):
        return {"ndarray": list(v.shape), "dtype": str(v.dtype)}
    return {"other": type(v).__name__}


def wrap_class(cls, role):
    for name, fn in list(vars(cls).items()):
        if not inspect.isfunction(fn) or (name.startswith("_") and name != "__init__"):
            continue

        def make(name, fn):
            @functools.wraps(fn)
            def wrapper(self, *a, **kw):
                from_main = caller_file() == MAIN
                if name == "__init__":
                    out = fn(self, *a, **kw)
                    register(self, role)
                    if from_main:
                        TRACE.append({"on": role, "call": "__init__", "args": [enc(x) for x in a], "kwargs": {k: enc(x) for k, x in kw.items()}})
                    return out
                if not from_main:
                    return fn(self, *a, **kw)
                rec = {"on": ROLES.get(id(self), role), "call": name, "args": [enc(x) for x in a], "kwargs": {k: enc(x) for k, x in kw.items()}}
                TRACE.append(rec)
                try:
                    out = fn(self, *a, **kw)
                except Exception as e:                   # main.py swallows some of these (try / except around the producers)
                    rec["raised"] = type(e).__name__
                    raise
                rec["returns"] = enc(out)
                return out
            return wrapper
        setattr(cls, name, make(name, fn))


def wrap_field_reads():
    """to_numpy / from_numpy / fill / element access on the solver's and the modules' fields, by calling module."""
    import taichi as ti
    seen = set()
    for cls_name in dir(ti):
        cls = getattr(ti, cls_name)
        if not inspect.isclass(cls):
            continue
        for c in cls.__mro__:
            if c in seen or c is object:
                continue
            seen.add(c)
            for op in ("to_numpy", "from_numpy", "fill", "copy_from"):
                fn = vars(c).get(op)
                if not inspect.isfunction(fn):
                    continue

                def make(op, fn):
                    @functools.wraps(fn)
                    def wrapper(self, *a, **kw):
                        f = caller_file()
                        if f.startswith(REF) and "<taichi_shim" not in f:
                            if id(self) not in FIELD_NAMES:
                                rescan()
                            name = FIELD_NAMES.get(id(self))
                            if name is not None:
                                mod = os.path.relpath(f, REF)
                                READERS.setdefault(mod, set()).add(f"{name}.{op}")
                                if f == MAIN:
                                    TRACE.append({"on": name, "call": op, "args": [enc(x) for x in a], "kwargs": {}})
                        return fn(self, *a, **kw)
                    return wrapper
                setattr(c, op, make(op, fn))


def probes(sim):
    """The attribute probes main.py's CoffeeSimulation and MinimalAdapter make on the solver, and what the reference's
    LBMSolver answers (these choose the code path of step_stable: main.py:803-824)."""
    src = open(MAIN).read()
    names = set(re.findall(r"hasattr\(self\.(?:lbm|_solver), '(\w+)'\)", src)) | set(re.findall(r"getattr\(solver, '(\w+)'", src))
    solver = sim.lbm._solver
    out = {}
    for n in sorted(names):
        present = hasattr(solver, n)
        v = getattr(solver, n, None)
        out[n] = {"present": bool(present), "none": v is None, "kind": "method" if callable(v) and not hasattr(v, "to_numpy") else ("field" if hasattr(v, "to_numpy") else type(v).__name__)}
    return out


def snapshot(sim, tag, store):
    s = sim.lbm._solver
    store[f"{tag}_rho"] = s.rho.to_numpy().astype(np.float32)
    store[f"{tag}_u"] = s.u.to_numpy().astype(np.float32)
    store[f"{tag}_phase"] = s.phase.to_numpy().astype(np.float32)
    store[f"{tag}_solid"] = s.solid.to_numpy().astype(np.uint8)
    store[f"{tag}_body_force"] = s.body_force.to_numpy().astype(np.float32)
    store[f"{tag}_phi"] = sim.multiphase.phi.to_numpy().astype(np.float32)
    ps = sim.particle_system
    store[f"{tag}_particle_count"] = np.int64(ps.particle_count[None])
    store[f"{tag}_particle_active"] = ps.active.to_numpy().astype(np.int32)
    store[f"{tag}_particle_pos"] = ps.position.to_numpy().astype(np.float32)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    pressure_mode = sys.argv[3] if len(sys.argv) > 3 else "none"
    out_json = sys.argv[4] if len(sys.argv) > 4 else os.path.join(HERE, "reference_main_trace.json")
    config = load_reference(n)
    for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.patches", "matplotlib.cm", "matplotlib.gridspec",
              "matplotlib.animation", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1", "seaborn"]:
        sys.modules[m] = mock.MagicMock()
    scratch = "/tmp/reference_main_trace_run"; os.makedirs(scratch, exist_ok=True); os.chdir(scratch)    # main.py writes report/ dirs
    with quiet():
        from src.core.legacy.lbm_solver import LBMSolver
        from src.core.multiphase_3d import MultiphaseFlow3D
        from src.physics.boundary_conditions import BoundaryConditionManager
        from src.physics.coffee_particles import CoffeeParticleSystem
        from src.physics.filter_paper import FilterPaperSystem
        from src.physics.precise_pouring import PrecisePouringSystem
        from src.physics.pressure_gradient_drive import PressureGradientDrive
    # main.py:551-561: `UnifiedLBMSolver(preferred_backend='auto')` failing -> `LBMSolver()`; the stand-in module takes that branch
    fallback = types.ModuleType("src.core.lbm_unified")
    fallback.UnifiedLBMSolver = lambda preferred_backend="auto": LBMSolver()
    sys.modules["src.core.lbm_unified"] = fallback
    for cls, role in ((LBMSolver, "lbm"), (MultiphaseFlow3D, "multiphase"), (PrecisePouringSystem, "pouring"), (FilterPaperSystem, "filter_paper"),
                      (PressureGradientDrive, "pressure_drive"), (CoffeeParticleSystem, "particle_system"), (BoundaryConditionManager, "boundary_manager")):
        wrap_class(cls, role)
    wrap_field_reads()

    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_main", MAIN)
    ref_main = importlib.util.module_from_spec(spec)
    sys.argv = ["main.py"]
    np.random.seed(20240601); random.seed(20240601)
    t0 = time.time()
    fields = {}
    marks = []
    ok_flags = []
    with quiet():
        spec.loader.exec_module(ref_main)
    # `python main.py debug <steps> <pressure_mode>` = run_debug_simulation (main.py:1250-1357, BASELINE configs[0]); the two hooks only
    # take snapshots: after the constructor returns and after every CoffeeSimulation.step()
    ctor, step = ref_main.CoffeeSimulation.__init__, ref_main.CoffeeSimulation.step

    def ctor_hook(self, *a, **kw):
        ctor(self, *a, **kw)
        marks.append({"phase": "constructed", "calls": len(TRACE)})
        snapshot(self, "init", fields)
        SIM.append(self)
        print(f"CoffeeSimulation constructed in {time.time() - t0:.0f} s, {len(TRACE)} calls recorded", file=sys.stderr, flush=True)

    def step_hook(self):
        ok = step(self)
        it = len(ok_flags)
        ok_flags.append(bool(ok))
        marks.append({"phase": f"step_{it}", "calls": len(TRACE), "ok": bool(ok)})
        snapshot(self, f"step{it}", fields)
        print(f"step {it}: ok={ok}, {len(TRACE)} calls, max|u| = {np.abs(fields[f'step{it}_u']).max():.3e}", file=sys.stderr, flush=True)
        if len(ok_flags) == steps:
            raise StopRecording()          # what follows in run() is report generation through matplotlib (mocked here: it never ends)
        return ok

    ref_main.CoffeeSimulation.__init__ = ctor_hook
    ref_main.CoffeeSimulation.step = step_hook
    try:
        with quiet():
            ref_main.run_debug_simulation(max_steps=steps, pressure_mode=pressure_mode)
    except StopRecording:
        pass
    marks.append({"phase": "finished", "calls": len(TRACE)})
    ok_all = all(ok_flags) and len(ok_flags) == steps
    out = {
        "grid": n, "steps": steps, "all_steps_ok": ok_all, "command": f"python main.py debug {steps} {pressure_mode}",
        "solver_class": type(SIM[0].lbm._solver).__name__,
        "constants": {"GRAVITY_LU": float(config.GRAVITY_LU), "DT": float(config.DT), "SCALE_TIME": float(config.SCALE_TIME),
                      "SCALE_LENGTH": float(config.SCALE_LENGTH), "TAU_WATER": float(config.TAU_WATER), "TAU_AIR": float(config.TAU_AIR)},
        "probes": probes(SIM[0]),
        "marks": marks,
        "trace": TRACE,
        "field_readers": {k: sorted(v) for k, v in sorted(READERS.items())},
    }
    with open(out_json, "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    np.savez_compressed(out_json.replace(".json", "_fields.npz"), **fields)
    print("wrote", out_json, len(TRACE), "calls;", {k: len(v) for k, v in out["field_readers"].items()})
