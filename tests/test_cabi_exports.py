"""CPU: the C-ABI library loads, exports every symbol include/picgpu.h declares, and refuses to compute without a GPU."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "engineering-degree-in-plasma-simulations_b200"


def _declared():
    src = open(os.path.join(ROOT, "include", "picgpu.h")).read()
    return sorted(set(re.findall(r"PICG_API\s+[\w\s\*]+?\b(picg_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    pg = importlib.import_module(PKG + ".picgpu")
    lib = pg.lib()
    names = _declared()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    pg = importlib.import_module(PKG + ".picgpu")
    if pg.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(pg.PicgError) as e:
        pg.init(0)
    assert e.value.code == -1
    h = ctypes.c_void_p()
    x = (ctypes.c_double * 3)(0, 0, 0)
    y = (ctypes.c_double * 3)(1, 1, 1)
    assert pg.lib().picg_world_create(5, 5, 5, x, y, ctypes.byref(h)) == -1       # PICG_ERR_NO_DEVICE
    assert b"no CPU fallback" in pg.lib().picg_last_error()


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or load anything under oracle/."""
    pkg_dir = os.path.join(ROOT, PKG)
    for dirpath, _, files in os.walk(pkg_dir):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pic_oracle" not in txt and "libref_" not in txt and "oracle/" not in txt, os.path.join(dirpath, f)
