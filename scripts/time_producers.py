#!/usr/bin/env python
"""Times the producers next to the step (csrc/lbm_producers.cu, lbm_particles.cu) on a V60 n^3 box: ms per call and the
algorithmic GB/s (DESIGN.md 3.4b bytes per cell x n^3 cells / time).  One JSON line per entry point."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from scripts.bench_configs import timed, v60_engine  # noqa: E402
from pour_over_coffee_lbm_b200 import _lib as L  # noqa: E402
from pour_over_coffee_lbm_b200.engine import ParticleState, _ptr  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
npart = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
eng = v60_engine(n, True, 0)
eng.step(2, write_macro_every=1)
cells = n ** 3
dev = eng.device
z = torch.arange(n, device=dev, dtype=torch.float32)[:, None, None]
g = torch.Generator(device=dev); g.manual_seed(7)
phi = torch.tanh((0.6 * n - z) / 2.0).expand(n, n, n).contiguous() + 0.01 * torch.randn((n, n, n), device=dev, generator=g)
phi.clamp_(-1.0, 1.0)
sc = lambda: torch.zeros_like(eng.rho)
vc = lambda: torch.zeros_like(eng.body_force)
phi_new, mu, curv, lap = sc(), sc(), sc(), sc()
grad_phi, grad_mu, normal, sf = vc(), vc(), vc(), vc()
blockage, accumulated = sc(), sc()
accumulated.fill_(3.0)
cfg = eng.cfg
sigma = cfg.SURFACE_TENSION_LU
pour = L.LbmPour(pour_x=n / 2, pour_y=n / 2, radius=float(np.float32(0.5 / cfg.GRID_SIZE_CM / 2.0)), pour_z=int(0.7 * n), velocity=0.05,
                 flow_rate=1.0, dt=1.0)

ps = ParticleState(npart, dev)
ps.pos[0].uniform_(0.3 * n, 0.7 * n, generator=g); ps.pos[1].uniform_(0.3 * n, 0.7 * n, generator=g); ps.pos[2].uniform_(6.0, 0.5 * n, generator=g)
ps.vel.normal_(0.0, 0.01, generator=g); ps.vel[2].sub_(0.02)
ps.radius.fill_(3.25e-4); ps.mass.fill_(float(4.0 / 3.0 * 3.14159 * 3.25e-4 ** 3 * 1200.0)); ps.active.fill_(1)
force = torch.zeros_like(ps.pos)
counters = torch.zeros(2, dtype=torch.int32, device=dev)
st = ps.struct()
lib, ctx, s = eng.lib, eng._ctx, eng.stream


def chk(rc):
    assert rc == 0, lib.lbm_last_error(ctx).decode()


parts = {
    # name: (callable, algorithmic bytes per call)
    "lbm_chemical_potential": (lambda: eng.chemical_potential(phi, lap, mu, 3.0 * sigma * 2.0 / 8.0), 12 * cells),
    "lbm_surface_tension (phi, mu -> grad, normal, curvature, F_s, body_force)":
        (lambda: eng.surface_tension(phi, mu, grad_phi, grad_mu, normal, curv, sf, sigma, apply=True), 117 * cells),
    "lbm_surface_tension (no mu / grad_mu)":
        (lambda: eng.surface_tension(phi, None, grad_phi, None, normal, curv, sf, sigma, apply=True), 101 * cells),
    "lbm_surface_tension_body_force (one launch, interface band only, no intermediate fields)":
        (lambda: eng.surface_tension_body_force(phi, sigma, normal, sf), 5 * cells),
    "lbm_apply_surface_tension": (lambda: eng.apply_surface_tension(sf), 41 * cells),
    "lbm_phase_field_step": (lambda: eng.phase_field_step(phi, phi_new, mu, 0.001, 1.0, cfg.RHO_WATER, cfg.RHO_AIR), 40 * cells),
    "lbm_density_from_phase": (lambda: eng.density_from_phase(phi, cfg.RHO_WATER, cfg.RHO_AIR), 12 * cells),
    "lbm_pouring_force (nozzle box only)": (lambda: eng.pouring_force(pour), 0),
    "lbm_pouring_phase_change (nozzle box only)": (lambda: eng.pouring_phase_change(pour, phi), 0),
    "lbm_filter_dynamic_resistance": (lambda: chk(lib.lbm_filter_dynamic_resistance(ctx, _ptr(eng.flags), _ptr(blockage), _ptr(accumulated), s)),
                                      1 * cells),
    f"lbm_particles_fluid_forces ({npart} particles)":
        (lambda: chk(lib.lbm_particles_fluid_forces(ctx, _ptr(eng.u), C.byref(st), _ptr(force), 965.3, 965.3 * 3.15e-7, 9.81, _ptr(counters), s)),
         64 * npart),
    f"lbm_particles_block_at_filter ({npart} particles, lattice-unit cells)":
        (lambda: chk(lib.lbm_particles_block_at_filter(ctx, C.byref(st), _ptr(eng.flags), _ptr(accumulated), 1.0, 0.01, 1, s)), 45 * npart),
    "reference point: lbm_step (V60, all features, no rho/u write-out)": (lambda: eng.step(1, write_macro_every=0), 0),
}
for name, (fn, nbytes) in parts.items():
    ms = timed(fn, 20, 5)
    out = {"part": name, "n": n, "ms": round(ms, 4)}
    if nbytes:
        out["algorithmic_GB/s"] = round(nbytes / (ms * 1e-3) / 1e9, 1)
    print(json.dumps(out), flush=True)
print(json.dumps({"fluid_fraction": float((eng.solid == 0).float().mean()), "active_after": int(ps.active.sum()),
                  "bounced_cells": int((accumulated > 3.0).sum())}), flush=True)
