"""B200Backend: the single compute backend, keeping the `ComputeBackend` interface of
src/core/backends/compute_backends.py:338-539 and the concrete-backend extras the unified solver
calls (`get_backend_info`, `estimate_memory_usage`, `validate_platform`; cuda_backend.py:243-371).

The reference's Apple/CUDA/CPU Taichi backends, the factory singleton and the fallback chain
(compute_backends.py:541-872) are replaced by this one class; there is nothing to fall back to.
"""
from __future__ import annotations

import time
from abc import ABC, abstractmethod
from contextlib import contextmanager
from typing import Any, Dict, Optional

import numpy as np
import torch

from .errors import BackendInitializationError, ComputeExecutionError


class ComputeBackend(ABC):
    def __init__(self, backend_type: str = "base"):
        self.backend_type = backend_type
        self.performance_metrics: Dict[str, Any] = {}
        self.error_history = []
        self.is_initialized = False
        self.creation_time = time.time()
        self.last_error = None
        self.execution_times = []
        self.memory_usage = []
        self.operation_count = 0

    @abstractmethod
    def execute_collision_streaming(self, memory_adapter, params=None, **kwargs) -> None: ...
    @abstractmethod
    def apply_boundary_conditions(self, memory_adapter, params=None, **kwargs) -> None: ...
    @abstractmethod
    def compute_macroscopic_quantities(self, memory_adapter, params=None, **kwargs) -> None: ...
    @abstractmethod
    def get_platform_info(self) -> Dict[str, Any]: ...
    @abstractmethod
    def get_performance_metrics(self) -> Dict[str, Any]: ...

    def initialize_backend(self) -> None:
        try:
            self._perform_initialization()
            self.is_initialized = True
        except Exception as e:
            msg = f"{self.backend_type} backend initialisation failed: {e}"
            self._record_error(msg, "INIT_FAILED")
            raise BackendInitializationError(msg, self.backend_type, "INIT_FAILED")

    def cleanup_backend(self) -> None:
        self._perform_cleanup()
        self.is_initialized = False

    @contextmanager
    def safe_execution(self, operation_name: str):
        start = time.time()
        try:
            yield
        except Exception as e:
            msg = f"{operation_name} failed: {e}"
            self._record_error(msg, "EXECUTION_FAILED")
            raise ComputeExecutionError(msg, self.backend_type, "EXECUTION_FAILED")
        finally:
            self.execution_times.append(time.time() - start)
            self.operation_count += 1
            if len(self.execution_times) > 1000:
                self.execution_times = self.execution_times[-500:]

    def _perform_initialization(self) -> None: ...
    def _perform_cleanup(self) -> None: ...

    def _record_error(self, error_msg: str, error_code: str) -> None:
        rec = {"timestamp": time.time(), "message": error_msg, "code": error_code, "backend_type": self.backend_type}
        self.error_history.append(rec)
        self.last_error = rec
        if len(self.error_history) > 100:
            self.error_history = self.error_history[-50:]

    def get_error_summary(self) -> Dict[str, Any]:
        return {"total_errors": len(self.error_history), "last_error": self.last_error,
                "error_rate": len(self.error_history) / max(1, self.operation_count), "backend_type": self.backend_type}

    def get_basic_metrics(self) -> Dict[str, Any]:
        if not self.execution_times:
            return {"status": "no_data"}
        return {"backend_type": self.backend_type, "total_operations": self.operation_count,
                "avg_execution_time": float(np.mean(self.execution_times)), "min_execution_time": float(np.min(self.execution_times)),
                "max_execution_time": float(np.max(self.execution_times)), "uptime_seconds": time.time() - self.creation_time,
                "is_initialized": self.is_initialized, "error_summary": self.get_error_summary()}


class B200Backend(ComputeBackend):
    def __init__(self):
        super().__init__("b200")
        self._solver = None
        self._ev = None
        self.performance_metrics = {"collision_time": 0.0, "streaming_time": 0.0, "boundary_time": 0.0, "total_time": 0.0}
        self.initialize_backend()

    def _perform_initialization(self) -> None:
        from . import _lib
        _lib.lib()                                        # raises if liblbm_b200.so is not built
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible (B200 / sm_100a required, no CPU fallback)")
        major, _ = torch.cuda.get_device_capability(0)
        if major != 10:
            raise RuntimeError(f"device is sm_{major}x, liblbm_b200 is built for sm_100a only")

    def bind(self, solver) -> None:
        self._solver = solver

    def validate_platform(self) -> bool:
        return self.is_initialized

    def _resolve(self, memory_adapter):
        s = memory_adapter if hasattr(memory_adapter, "engine") else self._solver
        if s is None or not hasattr(s, "engine"):
            raise ValueError("memory_adapter is not bound to a B200 solver (fields must live in the engine's HBM buffers)")
        return s

    def execute_collision_streaming(self, memory_adapter, params: Optional[dict] = None, **kwargs) -> None:
        """cuda_backend.py:197-241: one fused kernel instead of collision + streaming + boundary kernels."""
        with self.safe_execution("execute_collision_streaming"):
            s = self._resolve(memory_adapter)
            p = dict(params or {}, **kwargs)
            tau = p.get("tau")
            # compare in f32: the parameter block holds a c_float, and lbm_set_params re-validates the grid and drops the tensor-map cache
            if tau is not None and np.float32(tau) != np.float32(s.engine.params.tau_water):
                s.engine.set_params(tau_water=float(tau))
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            s.step()
            ev1.record()
            self._ev = (ev0, ev1)

    def apply_boundary_conditions(self, memory_adapter, params=None, **kwargs) -> None:
        with self.safe_execution("apply_boundary_conditions"):
            s = self._resolve(memory_adapter)
            s.boundary_manager.apply_all_boundaries(s)

    def compute_macroscopic_quantities(self, memory_adapter, params=None, **kwargs) -> None:
        with self.safe_execution("compute_macroscopic_quantities"):
            self._resolve(memory_adapter).compute_macroscopic_quantities()

    def get_platform_info(self) -> Dict[str, Any]:
        p = torch.cuda.get_device_properties(0)
        return {"name": p.name, "sm": f"{p.major}{p.minor}", "sm_count": p.multi_processor_count,
                "hbm_gb": p.total_memory / 1e9, "backend": "liblbm_b200 (hand-written sm_100a CUDA)"}

    def get_backend_info(self) -> Dict[str, Any]:
        return {"name": "B200 fused D3Q19 backend", "type": "b200", "platform": self.get_platform_info(),
                "kernels": "1 fused pull collide-stream launch per step"}

    def get_performance_metrics(self) -> Dict[str, Any]:
        """cuda_backend.py:298-329 conventions: throughput_mlups = NX*NY*NZ/1e6/total_time."""
        m = dict(self.performance_metrics)
        if self._ev is not None and self._solver is not None:
            self._ev[1].synchronize()
            t = self._ev[0].elapsed_time(self._ev[1]) * 1e-3
            cells = self._solver.engine.cells()
            m.update(total_time=t, collision_time=t, streaming_time=0.0, boundary_time=0.0,
                     throughput_mlups=cells / 1e6 / t if t > 0 else 0.0,
                     memory_bandwidth=cells * 19 * 4 * 2 / 1e9 / t if t > 0 else 0.0)
        return m

    def estimate_memory_usage(self, nx: int, ny: int, nz: int) -> float:
        n = nx * ny * nz
        return (2 * 19 * 4 + 4 + 2 * 12 + 12 + 4 + 1 + 1 + 4 + 4) * n / 1e9
