"""Taichi-field-like shims over torch CUDA tensors.

The reference's callers (main.py MinimalAdapter :378-442, the physics modules, diagnostics) touch
solver fields through the Taichi field surface: `.to_numpy()`, `.from_numpy()`, `.fill()`, `.shape`
and `[i, j, k]` get/set, in the logical index order [i,j,k] (x,y,z) / [i,j,k,c] / [q,i,j,k].
Device memory is x-fastest ([z,y,x]) for coalescing and contiguous z-halo planes, so these
classes expose permuted *views*; ghost planes of a slab are hidden.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch


class _FieldBase:
    owner = None          # the solver whose engine holds the tensor behind this field (set by LBMSolver / MultiphaseFlow3D)

    def __init__(self, on_write: Optional[Callable[[], None]] = None):
        self._on_write = on_write

    def _touch(self):
        if self._on_write is not None:
            self._on_write()

    # logical view (torch, no copy) -- subclasses implement
    def view(self) -> torch.Tensor:
        raise NotImplementedError

    @property
    def shape(self):
        return tuple(self.view().shape)

    @property
    def dtype(self):
        return self.view().dtype

    def to_torch(self) -> torch.Tensor:
        return self.view()

    def to_numpy(self) -> np.ndarray:
        return self.view().contiguous().cpu().numpy()

    def from_numpy(self, arr) -> None:
        v = self.view()
        v.copy_(torch.as_tensor(np.ascontiguousarray(arr)).to(v.device, v.dtype))
        self._touch()

    def from_torch(self, t: torch.Tensor) -> None:
        v = self.view()
        v.copy_(t.to(v.device, v.dtype))
        self._touch()

    def fill(self, value) -> None:
        self.view().fill_(value)
        self._touch()

    def copy_from(self, other: "_FieldBase") -> None:
        self.view().copy_(other.view())
        self._touch()

    def __getitem__(self, idx):
        out = self.view()[idx]
        if out.ndim == 0:
            return out.item()
        return out.cpu().numpy() if out.numel() <= 16 else out

    def __setitem__(self, idx, value):
        v = self.view()
        if isinstance(value, (list, tuple, np.ndarray)):
            value = torch.as_tensor(np.asarray(value)).to(v.device, v.dtype)
        v[idx] = value
        self._touch()


class ScalarField(_FieldBase):
    """[nzp, ny, nx] device tensor seen as [nx, ny, nz]."""

    def __init__(self, getter: Callable[[], torch.Tensor], zghost: int = 0, on_write=None):
        super().__init__(on_write)
        self._get, self._zg = getter, zghost

    def view(self):
        t = self._get()
        if self._zg:
            t = t[self._zg:t.shape[0] - self._zg]
        return t.permute(2, 1, 0)


class VectorField(_FieldBase):
    """[3, nzp, ny, nx] device tensor seen as [nx, ny, nz, 3] (ti.Vector.field surface)."""

    def __init__(self, getter: Callable[[], torch.Tensor], zghost: int = 0, on_write=None):
        super().__init__(on_write)
        self._get, self._zg = getter, zghost

    def view(self):
        t = self._get()
        if self._zg:
            t = t[:, self._zg:t.shape[1] - self._zg]
        return t.permute(3, 2, 1, 0)

    def fill(self, value) -> None:
        v = self.view()
        if isinstance(value, (list, tuple, np.ndarray)):
            for c in range(3):
                v[..., c].fill_(float(value[c]))
        else:
            v.fill_(value)
        self._touch()


class ComponentField(_FieldBase):
    """One component of a vector field (LBMSolver.ux / uy / uz, legacy/lbm_solver.py:270-276)."""

    def __init__(self, getter: Callable[[], torch.Tensor], comp: int, zghost: int = 0, on_write=None):
        super().__init__(on_write)
        self._get, self._c, self._zg = getter, comp, zghost

    def view(self):
        t = self._get()[self._c]
        if self._zg:
            t = t[self._zg:t.shape[0] - self._zg]
        return t.permute(2, 1, 0)


class PopulationField(_FieldBase):
    """The reference's `f[q,i,j,k]` (pre-collision, after streaming).

    The device stores post-collision populations (pull scheme); this field materialises the
    reference view on demand with lbm_export_f and writes through lbm_import_f -- both are exact
    data movement.  Reads are cached until the populations change: the engine bumps `populations_generation` in every call
    that rewrites g (step, init_equilibrium, import_f, load_checkpoint, a geometry change, pack_flags, populations_changed), so a
    reset or a restart can never hand back -- or write through -- the populations of before.
    """

    def __init__(self, engine, on_write=None):
        super().__init__(on_write)
        self._eng = engine
        self._cache = None
        self._cache_gen = -1

    def invalidate(self) -> None:
        self._cache, self._cache_gen = None, -1

    def _materialise(self) -> torch.Tensor:
        if self._cache is None or self._cache_gen != self._eng.populations_generation:
            self._cache = self._eng.export_f()
            self._cache_gen = self._eng.populations_generation
        return self._cache

    def view(self):
        t = self._materialise()
        zg = self._eng.zghost
        if zg:
            t = t[:, zg:t.shape[1] - zg]
        return t.permute(0, 3, 2, 1)

    def _flush(self):
        self._eng.import_f(self._cache)
        self._cache_gen = self._eng.populations_generation        # the cache IS what was just imported

    def from_numpy(self, arr) -> None:
        super().from_numpy(arr); self._flush()

    def from_torch(self, t) -> None:
        super().from_torch(t); self._flush()

    def fill(self, value) -> None:
        super().fill(value); self._flush()

    def __setitem__(self, idx, value):
        super().__setitem__(idx, value); self._flush()


class ConstField:
    """Small read-only lattice tables (cx, cy, cz, w, opposite_dir, e)."""

    def __init__(self, arr: np.ndarray):
        self._a = np.asarray(arr)

    @property
    def shape(self):
        return self._a.shape

    def to_numpy(self):
        return self._a.copy()

    def __getitem__(self, i):
        return self._a[i]


class TensorField(torch.Tensor):
    """A device tensor that also answers the Taichi-field calls the reference's readers make on the particle arrays
    (`particle_system.position.to_numpy()`, `.active.to_numpy()`: lbm_diagnostics.py, visualizer.py; recorded in
    tests/golden/reference_main_trace.json, "field_readers").  It stays a tensor for everything else: `wrap(t)` is a view of `t`."""

    @staticmethod
    def wrap(t: torch.Tensor) -> "TensorField":
        return t.as_subclass(TensorField)

    def to_numpy(self) -> np.ndarray:
        return np.ascontiguousarray(self.detach().as_subclass(torch.Tensor).cpu().numpy())

    def from_numpy(self, arr) -> None:
        self.copy_(torch.as_tensor(np.ascontiguousarray(arr)).to(self.device, self.dtype))

    def fill(self, value) -> None:
        self.fill_(value)

    def to_torch(self) -> torch.Tensor:
        return self.as_subclass(torch.Tensor)


class ScalarCount(int):
    """An int that can be read the way the reference reads its 0-D fields: `particle_count[None]`."""

    def __getitem__(self, _):
        return int(self)
