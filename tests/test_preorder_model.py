"""CPU model of the depth-first numbering behind walk_p_kernel's records (csrc/sfe_tree.cu recp_hist_kernel / recp_build_kernel):
pre-order id = (non-empty nodes whose slot range starts earlier) + (depth - depth of the shallowest node starting at the same
slot), successor = first pre-order id at the node's END slot - no traversal, one histogram and one scan over the slots.  Checked
here against an explicit depth-first traversal of oracle-built trees (the device kernels are checked bit for bit through the
walk itself, tests/test_gpu_walk.py); also the depth read off the cell extents, and width^2 = root width^2 / 4^depth."""
import numpy as np
import pytest

from oracle import sfo


def slot_ranges(nodes, n):
    """[start, end) of every node in the depth-first particle order, and its depth (an internal node's children partition
    its range in octant order; a leaf's bodies are its range)."""
    start = np.zeros(n, np.int64); end = np.zeros(n, np.int64); depth = np.zeros(n, np.int64)
    order = []                                         # explicit pre-order of the non-empty nodes
    def visit(i, s, d):
        start[i] = s; depth[i] = d
        b = int(nodes["bodies"][i])
        if b == 0:
            end[i] = s
            return s
        order.append(i)
        if b > 8:
            for c in nodes["branch"][i]:
                s = visit(int(c), s, d + 1)
        else:
            s += b
        end[i] = s
        return s
    import sys
    sys.setrecursionlimit(10000)
    visit(0, 0, 0)
    return start, end, depth, order


@pytest.mark.parametrize("N,box", [(3000, None), (20000, None), (5000, (-0.5, -0.4, -0.6, 0.7, 0.8, 0.6))])
def test_numbering_formula_equals_a_depth_first_traversal(N, box):
    pos, vel, mas = sfo.initcond(N)
    b = sfo.root_box(pos, 0) if box is None else np.array(box, np.float32)
    rc, nodes, n, keys, perm, nin = sfo.tree_build_sorted(b, pos, mas, 4 * N)
    assert rc == 0
    nodes = nodes[:n]
    start, end, depth, order = slot_ranges(nodes, n)
    M = int(end[0])
    assert M == nin
    live = np.array(order)
    # the kernels' arithmetic: one (count << 16 | depth) per node into its first slot, exclusive scan of the counts
    hist = np.zeros(M + 1, np.int64)
    np.add.at(hist, start[live], (1 << 16) | depth[live])
    cnt = hist >> 16; dsum = hist & 0xFFFF
    pref = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    k = cnt[start[live]]
    d0 = (dsum[start[live]] - k * (k - 1) // 2) // k
    pre = pref[start[live]] + (depth[live] - d0)
    assert np.array_equal(pre, np.arange(len(live)))                       # = position in the explicit traversal
    # successor: the first node that starts where the range ends; the first child of a cell is the next id
    nxt = np.where(end[live] < M, pref[np.minimum(end[live], M)], -1)
    pos_of = {int(i): p for p, i in enumerate(order)}
    for p, i in enumerate(order):
        # explicit successor: next node in pre-order outside i's subtree
        q = p + 1
        while q < len(order) and start[order[q]] < end[i]:
            q += 1
        assert nxt[p] == (q if q < len(order) else -1)
        if nodes["bodies"][i] > 8:
            first = next(int(c) for c in nodes["branch"][i] if nodes["bodies"][int(c)] > 0)
            assert pos_of[first] == p + 1
    # depth from the extents along the root's longest axis, width^2 from the depth
    ext = nodes["bounds"][:, 3:6] - nodes["bounds"][:, 0:3]
    ax = int(np.argmax(ext[0]))
    dd = np.rint(np.log2(ext[0, ax] / ext[live, ax])).astype(np.int64)
    assert np.array_equal(dd, depth[live])
    w = (ext[:, 0] + ext[:, 1] + ext[:, 2]).astype(np.float32) * np.float32(1.0 / 3.0)
    w2 = (w * w).astype(np.float32)
    w2a = (w2[0] * np.float32(0.25) ** depth[live].astype(np.float32)).astype(np.float32)
    assert np.max(np.abs(w2a - w2[live]) / w2[live]) <= 1.0e-4             # RECP_TOL: every cell of these trees is RECP_SURE
