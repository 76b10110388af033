"""ctypes wrapper around oracle/ref_cpu.c (TEST INFRASTRUCTURE, not product code).

Builds oracle/_build/libref_cpu.so on demand with the recipe in oracle/Makefile.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import d3q19_ref as R

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libref_cpu.so")


class RefParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("use_les", C.c_int), ("apply_filter", C.c_int), ("apply_faces", C.c_int),
                ("tau_water", C.c_float), ("tau_air", C.c_float), ("gravity_lu", C.c_float),
                ("les_cs", C.c_float), ("K_lu", C.c_float), ("beta_lu", C.c_float),
                ("c_darcy", C.c_float), ("c_forch", C.c_float)]


class RefFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("f", "f_new", "rho", "u", "u_sq", "phase", "body_force", "nu_sgs",
                 "solid", "les_mask", "filter_zone", "filter_blockage")]


class PhysCParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("sx", C.c_longlong), ("sy", C.c_longlong), ("sz", C.c_longlong), ("sq", C.c_longlong),
                ("v_cell", C.c_longlong), ("v_comp", C.c_longlong),
                ("per_x", C.c_int), ("per_y", C.c_int), ("per_z", C.c_int),
                ("use_force", C.c_int), ("use_phase", C.c_int), ("les", C.c_int), ("porous", C.c_int),
                ("tau_water", C.c_float), ("tau_air", C.c_float), ("gravity_lu", C.c_float), ("cs_smag", C.c_float),
                ("tau_min", C.c_float), ("tau_max", C.c_float), ("porous_darcy", C.c_float), ("porous_forch", C.c_float)]


class PhysCFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("g", "g_next", "rho", "u", "solid", "body_force", "phase", "filter_zone", "les_mask")]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "ref_cpu.c"), os.path.join(_HERE, "phys_cpu.c")]
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_SO)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_SO)
        for name in ("ref_les_update", "ref_macroscopic", "ref_collide_stream", "ref_swap_copy",
                     "ref_apply_filter_effects", "ref_face_bcs", "ref_init_fields"):
            getattr(_lib, name).argtypes = [C.POINTER(RefParams), C.POINTER(RefFields)]
            getattr(_lib, name).restype = None
        for name in ("ref_step", "ref_step_fused"):
            getattr(_lib, name).argtypes = [C.POINTER(RefParams), C.POINTER(RefFields), C.c_int]
            getattr(_lib, name).restype = None
        _lib.ref_v60_solid.argtypes = [C.c_int] * 3 + [C.c_float] * 4 + [C.c_void_p]
        _lib.ref_v60_solid.restype = None
        _lib.ref_num_threads.restype = C.c_int
        _lib.ref_fmaf_array.argtypes = [C.c_void_p] * 4 + [C.c_long]
        _lib.ref_fmaf_array.restype = None
        _lib.phys_step.argtypes = [C.POINTER(PhysCParams), C.POINTER(PhysCFields)]
        _lib.phys_step.restype = None
        _lib.phys_count_mismatch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]
        _lib.phys_count_mismatch.restype = C.c_longlong
    return _lib


def fmaf(a, b, c) -> np.ndarray:
    """Exact f32 fused multiply-add, element-wise with NumPy broadcasting (C99 fmaf, single rounding)."""
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float32), np.asarray(b, np.float32), np.asarray(c, np.float32))
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b); c = np.ascontiguousarray(c)
    out = np.empty(a.shape, np.float32)
    lib().ref_fmaf_array(a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data, out.size)
    return out


def num_threads() -> int:
    return int(lib().ref_num_threads())


class CState:
    """Owns contiguous f32 arrays in the reference layout and the C structs pointing at them."""

    def __init__(self, st: R.State):
        cfg = st.cfg
        self.cfg = cfg
        ca = np.ascontiguousarray
        self.f = ca(st.f, np.float32).copy(); self.f_new = ca(st.f_new, np.float32).copy()
        self.rho = ca(st.rho, np.float32).copy(); self.u = ca(st.u, np.float32).copy()
        self.u_sq = ca(st.u_sq, np.float32).copy(); self.phase = ca(st.phase, np.float32).copy()
        self.body_force = ca(st.body_force, np.float32).copy(); self.nu_sgs = ca(st.nu_sgs, np.float32).copy()
        self.solid = ca(st.solid, np.uint8).copy(); self.les_mask = ca(st.les_mask, np.int32).copy()
        self.filter_zone = None if st.filter_zone is None else ca(st.filter_zone, np.int32).copy()
        self.filter_blockage = None if st.filter_blockage is None else ca(st.filter_blockage, np.float32).copy()
        c_darcy, c_forch = R.filter_constants(cfg)
        self.params = RefParams(cfg.NX, cfg.NY, cfg.NZ, int(cfg.USE_LES), int(st.apply_filter),
                                int(st.apply_faces), cfg.TAU_WATER, cfg.TAU_AIR, cfg.GRAVITY_LU,
                                cfg.LES_CS, float(st.K_lu), float(st.beta_lu), float(c_darcy), float(c_forch))
        self._sync_ptrs()

    def _sync_ptrs(self):
        p = lambda a: None if a is None else a.ctypes.data
        self.fields = RefFields(p(self.f), p(self.f_new), p(self.rho), p(self.u), p(self.u_sq),
                                p(self.phase), p(self.body_force), p(self.nu_sgs), p(self.solid),
                                p(self.les_mask), p(self.filter_zone), p(self.filter_blockage))

    def step(self, n: int = 1, fused: bool = False):
        self._sync_ptrs()
        fn = lib().ref_step_fused if fused else lib().ref_step
        fn(C.byref(self.params), C.byref(self.fields), int(n))
        if fused and (n % 2 == 1):   # pointer swap happened inside C: mirror it
            self.f, self.f_new = self.f_new, self.f


def v60_solid(cfg: R.RefConfig) -> np.ndarray:
    out = np.empty((cfg.NX, cfg.NY, cfg.NZ), np.uint8)
    f32 = np.float32
    lib().ref_v60_solid(cfg.NX, cfg.NY, cfg.NZ,
                        float(f32(cfg.TOP_RADIUS / cfg.SCALE_LENGTH)), float(f32(cfg.BOTTOM_RADIUS / cfg.SCALE_LENGTH)),
                        float(f32(cfg.CUP_HEIGHT / cfg.SCALE_LENGTH)), float(f32(0.002 / cfg.SCALE_LENGTH)),
                        out.ctypes.data)
    return out


def phys_step(g, p: "R.PhysParams", solid=None, body_force=None, phase=None, filter_zone=None, les_mask=None, layout: str = "oracle"):
    """oracle/phys_cpu.c: the C twin of d3q19_ref.step_physical -- same arguments, same returns, same bits.
    layout = "oracle": g [19, NX, NY, NZ], body_force / u [NX, NY, NZ, 3] (the reference's index order, z fastest);
    layout = "device": g [19, NZ, NY, NX], body_force / u [3, NZ, NY, NX] (a download of the engine's buffers, x fastest)."""
    nx, ny, nz = p.nx, p.ny, p.nz
    vol = nx * ny * nz
    ca = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
    g = ca(g, np.float32); solid = ca(solid, np.uint8); body_force = ca(body_force, np.float32); phase = ca(phase, np.float32)
    filter_zone = ca(filter_zone, np.int32); les_mask = ca(les_mask, np.int32)
    if layout == "oracle":
        assert g.shape == (19, nx, ny, nz)
        sx, sy, sz, v_cell, v_comp = ny * nz, nz, 1, 3, 1
        rho = np.empty((nx, ny, nz), np.float32); u = np.empty((nx, ny, nz, 3), np.float32)
    else:
        assert g.shape == (19, nz, ny, nx)
        sx, sy, sz, v_cell, v_comp = 1, nx, nx * ny, 1, vol
        rho = np.empty((nz, ny, nx), np.float32); u = np.empty((3, nz, ny, nx), np.float32)
    if p.porous and filter_zone is None:
        filter_zone = np.zeros(rho.shape, np.int32)
    g_next = np.empty_like(g)
    cp = PhysCParams(nx, ny, nz, sx, sy, sz, vol, v_cell, v_comp, int(p.periodic[0]), int(p.periodic[1]), int(p.periodic[2]),
                     int(p.use_force), int(p.use_phase), int(p.les), int(p.porous), p.tau_water, p.tau_air, p.gravity_lu, p.cs_smag,
                     p.tau_min, p.tau_max, p.porous_darcy, p.porous_forch)
    ptr = lambda a: None if a is None else a.ctypes.data
    cf = PhysCFields(ptr(g), ptr(g_next), ptr(rho), ptr(u), ptr(solid), ptr(body_force), ptr(phase), ptr(filter_zone), ptr(les_mask))
    lib().phys_step(C.byref(cp), C.byref(cf))
    return g_next, rho, u


def count_mismatch(a, b, solid=None) -> int:
    """Words that differ (by bit pattern) between two f32 arrays of shape [K, *volume] on the fluid cells of `solid` ([*volume] u8)."""
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape
    vol = int(np.prod(a.shape[1:])) if a.ndim == 4 else int(a.size)
    k = a.shape[0] if a.ndim == 4 else 1
    s = None if solid is None else np.ascontiguousarray(solid, np.uint8)
    return int(lib().phys_count_mismatch(a.ctypes.data, b.ctypes.data, None if s is None else s.ctypes.data, vol, k))
