"""Error hierarchy of the reference's backend layer (src/core/backends/compute_backends.py:35-65),
kept name-for-name so callers' `except` clauses keep working."""
from __future__ import annotations

from typing import Optional


class BackendError(Exception):
    def __init__(self, message: str, backend_type: Optional[str] = None, error_code: Optional[str] = None):
        self.backend_type = backend_type or "unknown"
        self.error_code = error_code or "UNKNOWN_ERROR"
        super().__init__(message)


class PlatformDetectionError(BackendError):
    pass


class BackendInitializationError(BackendError):
    pass


class ComputeExecutionError(BackendError):
    pass


class MemoryAllocationError(BackendError):
    pass


class PerformanceDegradationError(BackendError):
    pass
