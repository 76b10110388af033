"""Shared helpers for the parity tests: layout conversion between the oracle's logical index
order ([q,i,j,k], [i,j,k,c]) and the device layout ([q,z,y,x], [c,z,y,x]), and seeded scenarios."""
from __future__ import annotations

import numpy as np

from oracle import d3q19_ref as R


def to_dev_scalar(a):      # [i,j,k] -> [z,y,x]
    return np.ascontiguousarray(np.transpose(a, (2, 1, 0)))


def to_dev_vec(a):         # [i,j,k,c] -> [c,z,y,x]
    return np.ascontiguousarray(np.transpose(a, (3, 2, 1, 0)))


def to_dev_pop(a):         # [q,i,j,k] -> [q,z,y,x]
    return np.ascontiguousarray(np.transpose(a, (0, 3, 2, 1)))


def from_dev_scalar(t):
    return np.ascontiguousarray(np.transpose(t.detach().cpu().numpy(), (2, 1, 0)))


def from_dev_vec(t):
    return np.ascontiguousarray(np.transpose(t.detach().cpu().numpy(), (3, 2, 1, 0)))


def from_dev_pop(t):
    return np.ascontiguousarray(np.transpose(t.detach().cpu().numpy(), (0, 3, 2, 1)))


def smooth_velocity(n, amp, seed, nz=None, ny=None):
    """Smooth seeded velocity field [nx,ny,nz,3] (a few low Fourier modes) -- stable at tau=0.53."""
    ny = ny or n; nz = nz or n
    rng = np.random.default_rng(seed)
    x = np.arange(n)[:, None, None] * (2 * np.pi / n)
    y = np.arange(ny)[None, :, None] * (2 * np.pi / ny)
    z = np.arange(nz)[None, None, :] * (2 * np.pi / nz)
    u = np.zeros((n, ny, nz, 3), np.float64)
    for c in range(3):
        for _ in range(3):
            kx, ky, kz = rng.integers(0, 3, size=3)
            ph = rng.uniform(0, 2 * np.pi, size=3)
            u[..., c] += rng.normal() * np.sin(kx * x + ph[0]) * np.sin(ky * y + ph[1] + 0.5) * np.cos(kz * z + ph[2])
    u *= amp / max(1e-12, np.abs(u).max())
    return u.astype(np.float32)


def smooth_density(n, amp, seed, nz=None, ny=None):
    ny = ny or n; nz = nz or n
    rng = np.random.default_rng(seed + 77)
    x = np.arange(n)[:, None, None] * (2 * np.pi / n)
    y = np.arange(ny)[None, :, None] * (2 * np.pi / ny)
    z = np.arange(nz)[None, None, :] * (2 * np.pi / nz)
    r = 1.0 + amp * np.sin(x + rng.uniform(0, 6)) * np.cos(y + rng.uniform(0, 6)) * np.cos(2 * z + rng.uniform(0, 6))
    return r.astype(np.float32)


def rel_err(a, b):
    """max |a-b| / max |b| (the 1e-5 criterion of BASELINE.md 4)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def reference_v60_state(n, seed=0, gravity=1e-4, body=1e-5, phase_mode="split"):
    """legacy LBMSolver state on an n^3 box with the V60 mask, a phase field, a small seeded
    body force and a smooth velocity perturbation folded into f (config table equilibrium)."""
    cfg = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=gravity)
    st = R.init_fields(cfg)
    R.attach_filter_system(st)
    rng = np.random.default_rng(seed)
    if phase_mode == "split":
        st.phase[:, :, : n // 2] = 1.0
        st.phase[:, :, n // 2:] = rng.uniform(0.0, 1.0, size=st.phase[:, :, n // 2:].shape).astype(np.float32)
    elif phase_mode == "water":
        st.phase[:] = 1.0
    st.body_force[:] = (body * rng.standard_normal(st.body_force.shape)).astype(np.float32)
    u0 = smooth_velocity(n, 0.02, seed)
    rho0 = smooth_density(n, 0.01, seed)
    for q in range(R.Q):
        st.f[q] = R.equilibrium_ref(rho0, u0[..., 0], u0[..., 1], u0[..., 2], q, "config")
        st.f_new[q] = st.f[q]
    return st


def l2_rel_err(a, b):
    """||a-b||_2 / ||b||_2"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.sqrt((b * b).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / (den if den > 0 else 1.0))


def assert_fast_build_close(got, ref, scale, tol=1e-5, what=""):
    """Tolerance bar for the FMA-contracted ("fast") build against the uncontracted oracle.

    BASELINE.json: "rho and u must agree within 1e-5 relative (fp32)".  The strict build meets it with ZERO error.
    For the fast build, relative means relative to the field's characteristic scale `scale` (1 for rho, the initial
    velocity amplitude U0 for u): max|d| <= 1e-5*scale.  A per-cell ratio is meaningless once the legacy solver's
    momentum sink (quirk Q1) has damped |u| to ~1e-6, i.e. to the f32 rounding floor of the populations."""
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    err = float(np.abs(got - ref).max()) / scale
    assert err <= tol, f"{what}: max|d|/scale = {err:.3e} > {tol}"


# ---- CPU emulation of product kernel source (tests/emu) ---------------------------------------------------------------
def build_emu(name, kernel_files):
    """Compile tests/emu/<name>.cpp -- a harness that #includes product CUDA source and runs it thread by thread -- with the
    host compiler and return the ctypes handle.  Rebuilt when the harness or any of `kernel_files` (under csrc/) is newer.
    Skips the calling test when there is no g++ or no CUDA headers (the emulation is extra coverage, not the parity proof)."""
    import ctypes
    import os
    import shutil
    import subprocess
    import pytest
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(os.path.dirname(here), "pour_over_coffee_lbm_b200", "csrc")
    src = os.path.join(here, "emu", name + ".cpp")
    lib = os.path.join(here, "emu", "_build", "lib" + name + ".so")
    inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("kernel-source emulation needs g++ and the CUDA headers")
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    deps = [src] + [os.path.join(csrc, f) for f in kernel_files]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        tmp = lib + f".{os.getpid()}.tmp"
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-w", "-shared", "-fPIC", "-I" + inc, "-include", "algorithm",
                        src, "-o", tmp], check=True)
        os.replace(tmp, lib)
    return ctypes.CDLL(lib)
