"""CPU models (numpy float32 = the device's unfused IEEE operations) of two conservative shortcuts of the walk, checked as
properties over random inputs:
  * the leaves' self-subtraction filter (csrc/sfe_tree.cu finalize_kernel -> WalkRec0.rin2, csrc/sfe_walk.cu): a target inside the
    leaf's box (gravity.c:473-476, both bounds inclusive) is never farther from the COM, along any axis, than the stored half edge;
  * the host-closed root gate (csrc/sfe_top.cuh top_nodes_kernel's summary, csrc/sfe_tree.cu tree_root_moments): when
    theta * kr < 0.99 every point inside the root box - except its very centre - OPENS the root (gravity.c:353-377)."""
import numpy as np
import pytest

f32 = np.float32


def rand_boxes(rng, n, cubic):
    centre = (rng.normal(0, 1, (n, 3)) * 10.0 ** rng.uniform(-3, 3, (n, 1))).astype(f32)
    half = (10.0 ** rng.uniform(-4, 2, (n, 1)) * (np.ones((n, 3)) if cubic else rng.uniform(0.05, 1.0, (n, 3)))).astype(f32)
    lo = (centre - half).astype(f32); hi = (centre + half).astype(f32)
    ok = np.all(hi > lo, axis=1)
    return lo[ok], hi[ok]


def points_inside(rng, lo, hi):
    """random points of the closed box, with faces and corners over-represented"""
    u = rng.uniform(0, 1, lo.shape)
    snap = rng.uniform(0, 1, lo.shape)
    u = np.where(snap < 0.15, 0.0, np.where(snap > 0.85, 1.0, u))
    p = (lo.astype(np.float64) + u * (hi.astype(np.float64) - lo.astype(np.float64))).astype(f32)
    return np.minimum(np.maximum(p, lo), hi)


@pytest.mark.parametrize("cubic", [True, False])
def test_leaf_filter_cube_holds_every_point_of_the_box(cubic):
    rng = np.random.default_rng(11)
    lo, hi = rand_boxes(rng, 200000, cubic)
    cm = points_inside(rng, lo, hi)                          # the COM of bodies inside the box (a rounded quotient: any point will do)
    # finalize_kernel: slack, farther face per axis, half edge
    slack = (f32(4e-6) * (np.abs(lo).sum(1, dtype=f32) + np.abs(hi).sum(1, dtype=f32))).astype(f32)
    far = np.maximum(np.abs((cm - lo).astype(f32)), np.abs((hi - cm).astype(f32))).max(1).astype(f32)
    rin = (((far + slack).astype(f32) * f32(1.0001)).astype(f32) + f32(1e-30)).astype(f32)
    for _ in range(4):
        t = points_inside(rng, lo, hi)
        reach = np.abs((t - cm).astype(f32)).max(1)          # the walk: max(|x - rx|, |y - ry|, |z - rz|) <= rin
        assert np.all(reach <= rin)


@pytest.mark.parametrize("cubic", [True, False])
def test_points_inside_the_root_box_open_the_root_when_theta_kr_is_below_one(cubic):
    rng = np.random.default_rng(12)
    lo, hi = rand_boxes(rng, 200000, cubic)
    e = (hi - lo).astype(f32)
    slack = (f32(4e-6) * (np.abs(lo).sum(1, dtype=f32) + np.abs(hi).sum(1, dtype=f32))).astype(f32)
    diag = np.sqrt((e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1] + e[:, 2] * e[:, 2]).astype(f32)).astype(f32)
    w = (f32(1.0 / 3.0) * ((e[:, 0] + e[:, 1]).astype(f32) + e[:, 2]).astype(f32)).astype(f32)       # gravity.c:368-372
    w2 = (w * w).astype(f32)
    hd = (f32(0.5) * diag + slack).astype(f32)
    rin2 = ((hd * hd).astype(f32) * f32(1.0001) + f32(1e-30)).astype(f32)
    kr = (rin2 / w2).astype(f32)
    theta = (f32(0.9899) / kr).astype(f32)                   # the largest theta the host shortcut lets through: theta * kr < 0.99
    assert np.all((theta * kr).astype(f32) < f32(0.99))
    c = (f32(0.5) * (lo + hi).astype(f32)).astype(f32)       # gravity.c:353-357
    for _ in range(4):
        t = points_inside(rng, lo, hi)
        o = (c - t).astype(f32)
        d2 = (((o[:, 0] * o[:, 0]).astype(f32) + (o[:, 1] * o[:, 1]).astype(f32)).astype(f32) + (o[:, 2] * o[:, 2]).astype(f32)).astype(f32)
        pos = d2 > 0
        ratio = (w2[pos] / d2[pos]).astype(f32)              # gravity.c:377
        assert np.all(theta[pos] <= ratio)                   # opened: nobody takes the root as a monopole
