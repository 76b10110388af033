"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/lbm_b200.h declares;
its compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pour_over_coffee_lbm_b200 import _lib
    lib = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lbm_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes prototypes out of sync with the header"
    assert lib.lbm_version() >= 100


def test_struct_layouts_match_header_field_order():
    from pour_over_coffee_lbm_b200 import _lib
    src = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    body = src[src.index("typedef struct {\n    int nx, ny, nz;"):src.index("} lbm_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        parts = decl.split(None, 1)
        fields += [f.strip() for f in parts[1].split(",")]
    assert fields == [f[0] for f in _lib.LbmParams._fields_]
    assert C.sizeof(_lib.LbmParams) == 4 * len(fields)
    fbody = re.sub(r"/\*.*?\*/", "", src[src.index("typedef struct {\n    float *f_src"):src.index("} lbm_fields;")], flags=re.S)
    ffields = [n.strip().lstrip("*") for decl in fbody.replace("typedef struct {", "").split(";") if decl.strip() for n in decl.strip().split(None, 1)[1].split(",")]
    assert ffields == [f[0] for f in _lib.LbmFields._fields_]
    assert C.sizeof(_lib.LbmFields) == 8 * len(ffields) and C.sizeof(_lib.LbmParticles) == 8 * 12 + 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pour_over_coffee_lbm_b200 import _lib
    from pour_over_coffee_lbm_b200.errors import BackendInitializationError
    from pour_over_coffee_lbm_b200.engine import D3Q19Engine
    lib = _lib.lib()
    ctx = C.c_void_p()
    p = _lib.LbmParams(nx=8, ny=8, nz=8, nz_global=8, periodic=7)
    assert lib.lbm_create(C.byref(ctx), 0, C.byref(p)) != 0
    assert b"no CPU fallback" in lib.lbm_last_error(None)
    with pytest.raises(BackendInitializationError):
        D3Q19Engine(8, 8, 8)
    from pour_over_coffee_lbm_b200.solver import UnifiedLBMSolver
    with pytest.raises(Exception):
        UnifiedLBMSolver()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pour_over_coffee_lbm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert not re.search(r"#\s*include\s*[\"<][^\">]*oracle", text), f
                assert "libref_cpu" not in text and "import_module(\"oracle" not in text, f
