"""The C-ABI library loads and exports exactly what include/sfengine.h declares; without a GPU every entry
point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sfengine.h")
LIB = os.path.join(ROOT, "starflood_b200", "libsfengine.so")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB]).decode()
    exported = set(re.findall(r" T (sfe_[a-z0-9_]+)", out))
    names = declared()
    assert len(names) >= 20
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    extra = sorted(exported - set(names))
    assert not extra, "exported but not declared in include/sfengine.h: %s" % extra


def test_python_binding_covers_header():
    from starflood_b200 import engine
    assert sorted(engine.SIGNATURES) == declared()
    L = engine.lib()
    for name in declared():
        assert hasattr(L, name)


def test_node_layout_is_reference_node_t():
    from starflood_b200 import engine
    assert engine.NODE_DTYPE.itemsize == 80            # tree.h:18-24 in fp32
    assert engine.NODE_DTYPE.fields["bounds"][1] == 40 and engine.NODE_DTYPE.fields["params"][1] == 64


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from starflood_b200 import engine
    assert engine.lib().sfe_init(-1) == engine.STARFLOOD_FAILURE
    assert "no CUDA device" in engine.last_error()
    sim = engine.Sim(64)
    sim.mas[:] = 1.0
    assert engine.solve_gravity_part_part(sim) == engine.STARFLOOD_FAILURE
    assert not sim.acc.any()                              # outputs untouched on failure
    engine.tree_init(sim.tree, 64)
    assert engine.tree_build(sim.tree, 64, sim.pos, sim.mas) == engine.STARFLOOD_FAILURE
    assert engine.solve_gravity_part_tree(sim) == engine.STARFLOOD_FAILURE


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "starflood_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "README.md", (dirpath, f)
