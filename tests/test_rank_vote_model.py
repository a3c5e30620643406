"""Host model of rank_kernel's warp-vote scheme (starflood_b200/csrc/sfe_tree.cu): the same mask arithmetic on Python
integers - four votes per group of 32 slots, rank = popc(own octant's mask & lanes before me & lanes from the last range
head on), lane o carrying the running count of octant o, per-warp / per-tile carries cut at range heads, node totals at
the last member - checked against a direct stable count for random range layouts (ranges spanning groups, warps and tiles,
heads and last members on lane 0 / 31, inactive gaps between ranges, tiles without any active slot).  The GPU parity tests
check the kernel itself bit for bit against the reference's tree; this pins the ALGORITHM for layouts real trees rarely have."""
import numpy as np
import pytest

GROUPS, WARPS = 8, 8                 # PT_GROUPS, PT_WARPS: a warp owns 8 groups of 32 slots, a tile 8 warps
TILE = 32 * GROUPS * WARPS
FULL = 0xFFFFFFFF


def popc(x):
    return bin(x & FULL).count("1")


def clz(x):
    return 32 - (x & FULL).bit_length()


def model(active, octant, head, last, node):
    """Returns (rank per slot, {node: totals[8]}) computed the way the kernel does."""
    M = len(active)
    ntiles = (M + TILE - 1) // TILE
    rank = np.zeros(M, np.int64)
    totals = {}
    incl = []                                                  # per tile: inclusive carry per octant (the look-back's result)
    for tile in range(ntiles):
        def slot(w, g, lane):
            return tile * TILE + w * 32 * GROUPS + 32 * g + lane
        def flag(arr, w, g, lane):
            s = slot(w, g, lane)
            return bool(arr[s]) if s < M else False
        # votes per (warp, group)
        hm = [[0] * GROUPS for _ in range(WARPS)]
        bo = [[[0] * 8 for _ in range(GROUPS)] for _ in range(WARPS)]   # bo[w][g][o]: what lane o holds
        for w in range(WARPS):
            for g in range(GROUPS):
                for lane in range(32):
                    if flag(head, w, g, lane):
                        hm[w][g] |= 1 << lane
                    if flag(active, w, g, lane):
                        bo[w][g][int(octant[slot(w, g, lane)])] |= 1 << lane
        if not any(bo[w][g][o] for w in range(WARPS) for g in range(GROUPS) for o in range(8)):
            incl.append([0] * 8)                               # tile without an active slot: zero prefix, nothing to rank
            continue
        # pass A: trailing counts per warp (from the last head of the chunk on)
        wtr = [[0] * 8 for _ in range(WARPS)]; whh = [False] * WARPS
        for w in range(WARPS):
            gl = -1
            for g in range(GROUPS):
                if hm[w][g]:
                    gl = g
            for g in range(GROUPS):
                if g >= gl:
                    for o in range(8):
                        m = bo[w][g][o]
                        if g == gl:
                            m &= (FULL << (31 - clz(hm[w][g]))) & FULL
                        wtr[w][o] += popc(m)
            whh[w] = gl >= 0
        # carries: previous warps of the tile (cw, cut at a head), whole tile (tt, tf), previous tiles (cm)
        tt = [0] * 8; tf = False
        cws = []
        for w in range(WARPS):
            cws.append((list(tt), tf))
            for o in range(8):
                tt[o] = wtr[w][o] if whh[w] else tt[o] + wtr[w][o]
            tf = tf or whh[w]
        cm = incl[tile - 1] if tile > 0 else [0] * 8
        incl.append([tt[o] + (0 if tf else cm[o]) for o in range(8)])
        # pass B
        for w in range(WARPS):
            cw, cf = cws[w]
            run = [cw[o] if cf else cw[o] + cm[o] for o in range(8)]
            for g in range(GROUPS):
                headmask = hm[w][g]
                for lane in range(32):
                    s = slot(w, g, lane)
                    if s >= M or not active[s]:
                        continue
                    oc = int(octant[s])
                    lt = (1 << lane) - 1
                    hb = headmask & (lt | (1 << lane))
                    in_open = hb == 0
                    seg_lo = 0 if in_open else 31 - clz(hb)
                    r = popc(bo[w][g][oc] & lt & ((FULL << seg_lo) & FULL))
                    if in_open:
                        r += run[oc]
                    rank[s] = r
                    if last[s]:
                        upto = FULL >> (31 - lane)
                        hbl = headmask & upto
                        rng = upto & ((FULL << (31 - clz(hbl))) & FULL) if hbl else upto
                        totals[int(node[s])] = [popc(bo[w][g][o] & rng) + (0 if hbl else run[o]) for o in range(8)]
                for o in range(8):
                    if headmask:
                        run[o] = popc(bo[w][g][o] & ((FULL << (31 - clz(headmask))) & FULL))
                    else:
                        run[o] += popc(bo[w][g][o])
    return rank, totals


def layout(rng, M, mean_range, gap_prob):
    """Random ranges of active slots (internal nodes: more than 8 members) with inactive gaps (leaf slots) between them."""
    active = np.zeros(M, bool); head = np.zeros(M, bool); last = np.zeros(M, bool)
    node = np.full(M, -1, np.int64); octant = rng.integers(0, 8, M)
    s, n = 0, 0
    while s < M:
        if rng.random() < gap_prob:
            s += int(rng.integers(1, 40))
            continue
        ln = 9 + int(rng.exponential(mean_range))
        e = min(M, s + ln)
        if e - s < 9:
            break
        active[s:e] = True; head[s] = True; last[e - 1] = True; node[s:e] = n
        n += 1; s = e
    return active, octant, head, last, node


@pytest.mark.parametrize("M,mean_range,gap_prob,seed", [
    (3 * TILE + 77, 20, 0.3, 1),        # many short ranges, several per group
    (2 * TILE, 400, 0.1, 2),            # ranges across groups and warps
    (4 * TILE + 5, 3000, 0.05, 3),      # ranges across tiles
    (TILE - 1, 5, 0.8, 4),              # mostly gaps
    (5 * TILE, 50000, 0.0, 5),          # one range over everything (the root's level)
    (3 * TILE, 30, 0.97, 6),            # whole tiles without an active slot
])
def test_vote_scheme_equals_a_direct_stable_count(M, mean_range, gap_prob, seed):
    rng = np.random.default_rng(seed)
    active, octant, head, last, node = layout(rng, M, mean_range, gap_prob)
    if seed == 2:                        # heads and last members on lanes 0 and 31
        active[:] = False; head[:] = False; last[:] = False; node[:] = -1
        for k, (a, b) in enumerate([(0, 32), (32, 63), (64, 96), (127, 160), (160, 1000), (1024, 1033), (1055, 1088)]):
            active[a:b] = True; head[a] = True; last[b - 1] = True; node[a:b] = k
    rank, totals = model(active, octant, head, last, node)
    want_rank = np.zeros(M, np.int64); want_tot = {}
    cnt = None
    for s in range(M):
        if not active[s]:
            continue
        if head[s]:
            cnt = [0] * 8
        want_rank[s] = cnt[octant[s]]; cnt[octant[s]] += 1
        if last[s]:
            want_tot[int(node[s])] = list(cnt)
    assert np.array_equal(rank[active], want_rank[active])
    assert totals == want_tot and len(want_tot) > 0
