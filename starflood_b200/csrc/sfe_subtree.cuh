// sfe_subtree.cuh - the lower part of the octree build: every node of 9 .. ST_MAX bodies is finished by ONE CTA.
//
// The level-by-level construction of sfe_tree.cu costs three dependent grid-wide kernels per tree level, and below the
// first few levels those kernels are pure latency: a level of the C2 tree moves 20 MB and takes 60-70 us.  Once a node is
// small enough for its members to fit into shared memory nothing it needs lives outside its own slot range any more -
// its subtree is a pure function of its members (in ascending index order) and its bounds.  So the global passes stop at
// such a node ("subtree root", tree.c:74-318 applied to the members of one cell) and one CTA builds everything below it
// in shared memory with __syncthreads() between the levels instead of kernel boundaries:
//
//   per local level   every slot of an internal node picks its octant (tie rule of tree.c:244-247 / :282-285) and gets its
//                     stable rank inside (node, octant) from warp votes (same scheme as rank_kernel, carries between the
//                     warps through shared memory - the whole subtree is one tile, so there is no look-back);
//                     the eight children of every internal node are created (bounds tree.c:219-226, slot ranges,
//                     depth-first successors), the order array is partitioned (ping-pong in shared memory) and the key
//                     digit appended; the moments of the new children are folded in slot (= ascending index) order,
//                     products rounded one by one exactly like tree.c:166-169;
//   at the end        the final (key, index) order of the range is written back over the global slot arrays in place.
//
// Node ids of a subtree are drawn from the global counter with one atomicAdd per CTA and level, so the numbering below
// the subtree roots depends on CTA scheduling; nothing in the engine depends on node numbers (the walk follows links),
// and sfe_tree_download() renumbers the table breadth-first, which is the canonical numbering the tests compare.
#pragma once

namespace sfe {

constexpr int ST_THREADS = 512;
constexpr int ST_MAX = 2048;                         // most bodies a subtree root may hold
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_GROUPS = ST_MAX / ST_THREADS;       // groups of 32 slots per warp
constexpr int ST_NODES = ST_MAX / 9 + 1;             // internal nodes of one level (each holds at least 9 bodies)
constexpr uint32_t ST_LEAF = 0xFFu;
static_assert(ST_NODES < 255, "node indices of a level are stored in bytes");

struct StNode {                 // an internal node of the level that is being split
    uint32_t gid, cbase, next;  // global id, first child id, depth-first successor when not opened
    uint16_t start, end;        // local slot range
    float lo[3], hi[3];
};

struct StShared {
    float4 rec[ST_MAX];              // particle records by local id (= position in the root's range = ascending index)
    unsigned long long key[ST_MAX];  // path key by local id
    uint32_t pidx[ST_MAX];           // particle index by local id
    uint16_t ord[2][ST_MAX];         // slot -> local id, ping-pong between levels
    uint8_t snode[2][ST_MAX];        // slot -> index of the owning internal node of the current level (ST_LEAF: leaf slot)
    StNode nodes[2][ST_NODES];
    float4 nmid[2][ST_NODES];        // split planes of the current / next level (tree.c:159-163)
    uint16_t ncount[ST_NODES][8];    // per-octant member counts
    uint16_t cstart[ST_NODES][8];    // first slot of each child
    uint8_t cidx[ST_NODES][8];       // index of an internal child in the next level's node list
    uint32_t wtr[ST_WARPS][8];       // per warp: counts of the range that is open at the end of its chunk ...
    uint32_t whh[ST_WARPS];          // ... counted from the last head inside the chunk, if there is one
    uint32_t ticket, nnext, base, abort;
};

__global__ void __launch_bounds__(ST_THREADS, 2) subtree_kernel(TreeDev t, int depth_max) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    StShared& S = *reinterpret_cast<StShared*>(st_smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, myo = lane & 7u;
    const uint32_t lt = (1u << lane) - 1u, le = lt | (1u << lane);
    if (t.meta[M_STATUS] != 0u) return;
    const uint32_t nsub = *t.sub_count;
    const uint32_t fin = t.meta[M_FINAL];
    SlotRec* __restrict__ g_rec = t.srec[fin];
    for (;;) {
        __syncthreads();                                   // the previous subtree is completely written back
        if (tid == 0) { S.ticket = atomicAdd(t.sub_ticket, 1u); S.abort = 0u; }
        __syncthreads();
        const uint32_t k = S.ticket;
        if (k >= nsub) return;
        const uint2 ent = t.sub_list[k];                   // {node id, level}
        const uint32_t root = ent.x;
        const uint2 rsg = t.seg[root];
        const uint32_t slot0 = rsg.x, cnt = rsg.y - rsg.x;
        // ids for the children of `count` nodes of level lvl (tree.c:180-189: node budget first, then depth).  One global atomic
        // per CTA and level, issued by thread 0 a phase ahead of its use so that its latency hides behind other work.
        auto alloc_ids = [&](uint32_t count, int lvl) {
            if (depth_max <= lvl) { atomicCAS(&t.meta[M_STATUS], 0u, 2u); S.abort = 1u; }
            else {
                const uint32_t b = atomicAdd(&t.meta[M_COUNT], 8u * count);
                if (t.nodes_max <= b + 8u * count) { atomicCAS(&t.meta[M_STATUS], 0u, 1u); S.abort = 1u; }
                S.base = b;
                atomicMax(&t.meta[M_DEPTH], (uint32_t)lvl + 1u);
            }
        };
        // first child of the nodes [0, n) of list `buf`, whose children start at S.base (read two barriers later)
        auto setup_nodes = [&](uint32_t buf, uint32_t n) {
            for (uint32_t j = tid; j < n; j += ST_THREADS) {
                StNode& nd = S.nodes[buf][j];
                nd.cbase = S.base + 8u * j;
                t.child_base[nd.gid] = nd.cbase;
            }
        };
        int level = (int)ent.y;
        if (tid == 0) { S.nnext = 0u; alloc_ids(1u, level); }     // (in flight while the members are loaded)
        // ---------------------------------------------------------------- load the members (ascending index order)
        for (uint32_t i = tid; i < (uint32_t)ST_MAX; i += ST_THREADS) {
            if (i < cnt) {
                const SlotRec r = slot_load(&g_rec[slot0 + i]);
                S.rec[i] = r.pm;
                S.pidx[i] = r.idx;
                S.key[i] = r.key;
                S.ord[0][i] = (uint16_t)i;
                S.snode[0][i] = 0u;
            } else {
                S.snode[0][i] = (uint8_t)ST_LEAF;
                S.snode[1][i] = (uint8_t)ST_LEAF;
            }
        }
        if (tid == 0) {
            StNode nd;
            nd.gid = root; nd.cbase = 0u; nd.next = t.next[root];
            nd.start = 0; nd.end = (uint16_t)cnt;
            const WalkRec2 cell = t.rec2[root];
            for (int q = 0; q < 3; q++) { nd.lo[q] = cell.lo[q]; nd.hi[q] = cell.hi[q]; }
            S.nodes[0][0] = nd;
            S.nmid[0][0] = make_float4(__fmul_rn(0.5f, __fadd_rn(nd.lo[0], nd.hi[0])), __fmul_rn(0.5f, __fadd_rn(nd.lo[1], nd.hi[1])),
                                       __fmul_rn(0.5f, __fadd_rn(nd.lo[2], nd.hi[2])), 0.f);   // tree.c:159-163
        }
        uint32_t ncur = 1u, cur = 0u;
        bool first = true;
        __syncthreads();
        while (ncur > 0u) {
            const uint32_t nxt = cur ^ 1u;
            if (S.abort) return;                           // the build has failed (uniform: read behind the barrier); tree_status() reports it
            setup_nodes(cur, ncur);                        // (the ids arrived during the previous phase)
            // ------------------------------------------------------------ octants + stable ranks (warp votes)
            uint32_t b[ST_GROUPS], r[ST_GROUPS], lid[ST_GROUPS], nn[ST_GROUPS], hm[ST_GROUPS], bo[ST_GROUPS];
            const uint32_t w0 = warp * (ST_GROUPS * 32) + lane;
#pragma unroll
            for (int g = 0; g < ST_GROUPS; g++) {
                const uint32_t s = w0 + 32u * g;
                const uint32_t n = S.snode[cur][s];
                uint32_t byte = 0u, l = 0u;
                if (n != ST_LEAF) {
                    const uint32_t ns = S.nodes[cur][n].start, ne = S.nodes[cur][n].end;
                    l = S.ord[cur][s];
                    byte = pick_octant(S.nmid[cur][n], S.rec[l], s - ns < 8u) | SB_ACTIVE | (s == ns ? SB_HEAD : 0u) | (s + 1u == ne ? SB_LAST : 0u);
                }
                b[g] = byte; nn[g] = n; lid[g] = l; r[g] = 0u;
            }
            int gl = -1;
            uint32_t live = 0u;                            // groups with an active slot (most slots of the deeper levels sit in leaves)
#pragma unroll
            for (int g = 0; g < ST_GROUPS; g++) {
                hm[g] = 0u; bo[g] = 0u;
                if (__any_sync(0xFFFFFFFFu, (b[g] & SB_ACTIVE) != 0u)) {
                    live |= 1u << g;
                    hm[g] = __ballot_sync(0xFFFFFFFFu, (b[g] & SB_HEAD) != 0u);
                    if (hm[g]) gl = g;
                    bo[g] = octant_mask(true, b[g] & 7u, lane) & __ballot_sync(0xFFFFFFFFu, (b[g] & SB_ACTIVE) != 0u);
                }
            }
            uint32_t tr = 0u;
#pragma unroll
            for (int g = 0; g < ST_GROUPS; g++)
                if (g >= gl) tr += __popc(g == gl ? (bo[g] & (0xFFFFFFFFu << (31 - __clz(hm[g])))) : bo[g]);
            if (lane < 8u) S.wtr[warp][lane] = tr;
            if (lane == 0u) S.whh[warp] = gl >= 0 ? 1u : 0u;
            __syncthreads();
            uint32_t run = 0u;                             // octant myo of the range open at the start of the warp's chunk
            for (uint32_t w = 0; w < warp; w++) {
                const uint32_t v = S.wtr[w][myo];
                if (S.whh[w] != 0u) run = v; else run += v;
            }
#pragma unroll
            for (int g = 0; g < ST_GROUPS; g++) {
                if (!((live >> g) & 1u)) continue;         // nothing to rank, nothing open across the group (ranges are contiguous)
                const uint32_t byte = b[g], oc = byte & 7u;
                const bool act = (byte & SB_ACTIVE) != 0u;
                const uint32_t headmask = hm[g];
                const uint32_t hb = headmask & le;
                const bool in_open = hb == 0u;
                const uint32_t seg_lo = in_open ? 0u : 31u - (uint32_t)__clz(hb);
                uint32_t rr = __popc(__shfl_sync(0xFFFFFFFFu, bo[g], (int)oc) & lt & (0xFFFFFFFFu << seg_lo));
                const uint32_t basev = __shfl_sync(0xFFFFFFFFu, run, (int)oc);
                if (in_open) rr += basev;
                r[g] = rr;
                for (uint32_t m = __ballot_sync(0xFFFFFFFFu, act && (byte & SB_LAST) != 0u); m; m &= m - 1u) {
                    const uint32_t L = (uint32_t)__ffs(m) - 1u, upto = 0xFFFFFFFFu >> (31u - L);
                    const uint32_t hbl = headmask & upto;
                    const uint32_t range = hbl ? (upto & (0xFFFFFFFFu << (31 - __clz(hbl)))) : upto;
                    const uint32_t val = __popc(bo[g] & range) + (hbl ? 0u : run);
                    const uint32_t n = __shfl_sync(0xFFFFFFFFu, nn[g], (int)L);
                    if (lane < 8u) S.ncount[n][lane] = (uint16_t)val;
                }
                if (headmask) run = __popc(bo[g] & (0xFFFFFFFFu << (31 - __clz(headmask)))); else run += __popc(bo[g]);
            }
            __syncthreads();
            // ------------------------------------------------------------ the eight children of every node (tree.c:191-274)
            for (uint32_t q = tid; q < 8u * ncur; q += ST_THREADS) {
                const uint32_t j = q >> 3, o = q & 7u;
                const StNode& nd = S.nodes[cur][j];
                uint32_t start = nd.start, cn = 0u, succ = nd.next;
#pragma unroll
                for (uint32_t o2 = 0; o2 < 8u; o2++) {
                    const uint32_t c2 = S.ncount[j][o2];
                    if (o2 < o) start += c2;
                    if (o2 == o) cn = c2;
                }
#pragma unroll
                for (uint32_t o2 = 7u; o2 > 0u; o2--) if (o2 > o && S.ncount[j][o2] > 0u) succ = nd.cbase + o2;
                const uint32_t gid = nd.cbase + o;
                const float4 mid = S.nmid[cur][j];
                const float md[3] = {mid.x, mid.y, mid.z};
                float clo[3], chi[3];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    clo[a] = (o >> a) & 1u ? md[a] : nd.lo[a];           // tree.c:219-226
                    chi[a] = (o >> a) & 1u ? nd.hi[a] : md[a];
                }
                float4* cell = reinterpret_cast<float4*>(&t.rec2[gid]);
                cell[0] = make_float4(clo[0], clo[1], clo[2], chi[0]);
                cell[1] = make_float4(chi[1], chi[2], __uint_as_float(nd.gid), 0.f);
                t.seg[gid] = make_uint2(slot0 + start, slot0 + start + cn);
                t.next[gid] = succ;
                t.child_base[gid] = NODE_END;                            // set when the child is split itself
                t.counts[(size_t)nd.gid * 8u + o] = cn;
                if (cn == 0u) t.rec1[gid] = make_float4(0.f, 0.f, 0.f, 0.f);   // moments of empty nodes stay zero (tree.c:302-310)
                S.cstart[j][o] = (uint16_t)start;
                uint32_t ci = ST_LEAF;
                if (cn > 8u) {
                    ci = atomicAdd(&S.nnext, 1u);
                    StNode c;
                    c.gid = gid; c.cbase = 0u; c.next = succ; c.start = (uint16_t)start; c.end = (uint16_t)(start + cn);
                    for (int a = 0; a < 3; a++) { c.lo[a] = clo[a]; c.hi[a] = chi[a]; }
                    S.nodes[nxt][ci] = c;
                    S.nmid[nxt][ci] = make_float4(__fmul_rn(0.5f, __fadd_rn(clo[0], chi[0])), __fmul_rn(0.5f, __fadd_rn(clo[1], chi[1])),
                                                  __fmul_rn(0.5f, __fadd_rn(clo[2], chi[2])), 0.f);
                }
                S.cidx[j][o] = (uint8_t)ci;
            }
            __syncthreads();
            // ------------------------------------------------------------ stable 8-way partition of the order array
#pragma unroll
            for (int g = 0; g < ST_GROUPS; g++) {
                const uint32_t s = w0 + 32u * g;
                if (b[g] & SB_ACTIVE) {
                    const uint32_t oc = b[g] & 7u, n = nn[g];
                    const uint32_t dest = (uint32_t)S.cstart[n][oc] + r[g];
                    S.ord[nxt][dest] = (uint16_t)lid[g];
                    S.snode[nxt][dest] = S.cidx[n][oc];
                    if (level < 16) S.key[lid[g]] |= (unsigned long long)oc << (45 - 3 * level);   // digit `level` of the 48-bit path key
                } else if (s < cnt) {
                    S.ord[nxt][s] = S.ord[cur][s];
                    S.snode[nxt][s] = (uint8_t)ST_LEAF;
                }
            }
            __syncthreads();
            // ids for the next level's children: one thread of the last warp (idle or nearly so in the fold phase) fetches them
            // while the folds run
            if (tid == ST_THREADS - 1 && S.nnext > 0u) alloc_ids(S.nnext, level + 1);
            // ------------------------------------------------------------ moments of the new children (tree.c:166-169, :302-310)
            // one lane per (node, component).  A subtree root's own fold is done by the fold kernels of its level like any
            // other node's - except the tree root's (the whole tree is one subtree): item 8*ncur, its range is the identity
            {
                const uint32_t items = 8u * ncur + ((first && root == 0u) ? 1u : 0u);
                const uint32_t comp = tid & 3u;
                for (uint32_t wq = tid >> 2; wq < ((items + 7u) & ~7u); wq += ST_THREADS / 4) {
                    uint32_t s0 = 0u, cn = 0u, gid = 0u;
                    const uint16_t* __restrict__ order = S.ord[nxt];
                    if (wq < 8u * ncur) {
                        const uint32_t j = wq >> 3, o = wq & 7u;
                        cn = S.ncount[j][o]; s0 = S.cstart[j][o]; gid = S.nodes[cur][j].cbase + o;
                    } else if (wq < items) {
                        cn = cnt; s0 = 0u; gid = root; order = nullptr;
                    }
                    // scalar component loads (no select on a loaded float4: ptxas turned that into a branch per element);
                    // the loads do not depend on the running sum, so the unrolled body issues them ahead of the FADD chain
                    const float* __restrict__ recf = reinterpret_cast<const float*>(S.rec);
                    float sum = 0.f;
                    if (order) {
                        const uint16_t* __restrict__ op = order + s0;
#pragma unroll 8
                        for (uint32_t i = 0; i < cn; i++) {
                            const uint32_t l = op[i];
                            const float m = recf[4u * l + 3u], v = recf[4u * l + comp];
                            sum = __fadd_rn(sum, comp == 3u ? m : __fmul_rn(m, v));   // product rounded, then added
                        }
                    } else {
#pragma unroll 8
                        for (uint32_t i = 0; i < cn; i++) {
                            const float m = recf[4u * i + 3u], v = recf[4u * i + comp];
                            sum = __fadd_rn(sum, comp == 3u ? m : __fmul_rn(m, v));
                        }
                    }
                    const float mass = __shfl_sync(0xFFFFFFFFu, sum, (int)(lane | 3u));
                    if (cn) {
                        const float inv = 0.f < mass ? __fdiv_rn(1.0f, mass) : 1.0f;   // tree.c:303-304
                        reinterpret_cast<float*>(&t.rec1[gid])[comp] = comp < 3u ? __fmul_rn(sum, inv) : sum;
                    }
                }
            }
            first = false;
            ncur = S.nnext;
            cur = nxt;
            level++;
            __syncthreads();
            if (tid == 0) S.nnext = 0u;                    // (everybody has read it; the next atomicAdd on it is two barriers away)
        }
        // ---------------------------------------------------------------- final (key, index) order of the range, in place
        for (uint32_t s = tid; s < cnt; s += ST_THREADS) {
            const uint32_t l = S.ord[cur][s];
            slot_store(&g_rec[slot0 + s], S.rec[l], S.key[l], S.pidx[l], NODE_END);
        }
    }
}

}  // namespace sfe
