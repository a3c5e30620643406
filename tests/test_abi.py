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


def test_new_entry_points_fail_loudly_without_gpu(tmp_path):
    """Resident leapfrog, snapshots, overlay and the peer exchange have no CPU path either."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from starflood_b200 import engine
    L = engine.lib()
    sim = engine.Sim(32); sim.mas[:] = 1.0
    prm = engine.sim_params()
    assert engine.sim_upload(sim) == engine.STARFLOOD_FAILURE and "no CUDA device" in engine.last_error()
    assert engine.sim_solv_resident(sim, prm) == engine.STARFLOOD_FAILURE
    assert engine.sim_step(sim, prm)[0] == engine.STARFLOOD_FAILURE
    assert engine.sim_download(sim) == engine.STARFLOOD_FAILURE
    assert engine.sim_save_raw_async(str(tmp_path / "x.raw")) == engine.STARFLOOD_FAILURE
    assert not (tmp_path / "x.raw").exists()
    h = C.create_string_buffer(64)
    assert L.sfe_peer_create(0, 1, 1024, h) == engine.STARFLOOD_FAILURE
    assert L.sfe_dev_peer_put(0, 16, None) == engine.STARFLOOD_FAILURE
    assert L.sfe_dev_bh_peer(None, 32, 0, 0, 1.0, 1e-6, 0.5, 0, None) == engine.STARFLOOD_FAILURE
    rb = np.zeros(64, np.float32); m = np.eye(4, dtype=np.float32)
    assert L.sfe_bh_overlay(3, 0.5, m.ctypes.data, rb.ctypes.data, 4, 4) == engine.STARFLOOD_FAILURE and not rb.any()
    vis = engine.Vis(); engine.vis_init(vis, 8, 8)
    assert engine.vis_draw(vis, sim) == engine.STARFLOOD_FAILURE and "no CUDA device" in engine.last_error()
    assert engine.sim_vis_draw(vis, 0) == engine.STARFLOOD_FAILURE
    assert not vis.atomic_buffer.any() and not vis.render_buffer.any()
    assert L.sfe_peer_window() is None
    assert L.sfe_sim_peer_bytes(1000, 13) == 48 * 1024          # pure arithmetic, no device involved


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "starflood_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "README.md", (dirpath, f)
