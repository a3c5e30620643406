// sfe_tree.cuh - device-side view of the octree shared by the build (sfe_tree.cu) and the walk (sfe_walk.cu).
#pragma once
#include "sfe_internal.cuh"

namespace sfe {

constexpr uint32_t NODE_END = 0xFFFFFFFFu;

// meta[] words (device-side counters of one build)
enum : int {
    M_COUNT = 0,       // nodes allocated so far
    M_STATUS = 1,      // 0 ok, 1 node overflow (tree.c:180-183), 2 depth overflow (tree.c:186-189)
    M_NUM_INBOX = 2,   // particles inside the root box
    M_DEPTH = 3,       // deepest level that holds nodes
    M_FINAL = 4,       // level whose slot arrays (perm[], spm4[]) hold the final (key, index)-sorted order
    M_LV_BEGIN = 5,                        // [SFE_MAX_LEVELS + 3] first node id of each level
    M_LV_END = M_LV_BEGIN + SFE_MAX_LEVELS + 3,   // [SFE_MAX_LEVELS + 3] one past the last node id of each level
    M_WORDS = M_LV_END + SFE_MAX_LEVELS + 3
};

// walk records --------------------------------------------------------------------------------------
constexpr uint32_t KIND_EMPTY = 0u, KIND_SINGLE = 1u, KIND_LEAF = 2u, KIND_INTERNAL = 3u;

struct __align__(32) WalkRec0 {   // 32 B = one L1 sector, fetched with ONE 256-bit load: everything a visit decides on
    float a[4];                   // internal: geometric centre xyz (gravity.c:353-357), fl(width^2) (gravity.c:368-372,377)
                                  // leaf    : centre of mass xyz, mass - a leaf is always a monopole (gravity.c:331), so
                                  //           its visit costs exactly this one sector
    uint32_t next;                // next NON-EMPTY node in depth-first order when this node is not opened
    uint32_t child;               // first NON-EMPTY child (internal nodes)
    float rin2;                   // filter in front of the exact bounds test of the self-subtraction (gravity.c:473-476): a point
                                  // is surely outside the box when its squared distance from the centre exceeds this (internal),
                                  // when it is farther than this from the COM along any axis (leaf: the half edge of a cube);
                                  // -1 for a single body (nothing to subtract: gravity.c:463-470 only looks at bodies > 1)
    uint32_t kind;                // KIND_* (a word of its own: no shift / mask on the visit's critical path)
};
struct __align__(16) WalkRec2 {   // 32 B: the cell itself - bounds (gravity.c:473-476 self-subtraction test, tree.c:219-226) + parent
    float lo[3], hi[3];
    uint32_t parent;
    uint32_t pad;
};
static_assert(sizeof(WalkRec0) == 32 && sizeof(WalkRec2) == 32, "walk records are 32 bytes");

// The throughput walk's record: EVERYTHING a visit needs in one 32-byte sector, in DEPTH-FIRST (pre-order) numbering of the
// non-empty nodes - the first non-empty child of node p is p + 1, so only the successor needs a link, and an accepted cell
// costs one scattered request instead of two (decision record + COM record: profiles/r02_walk_exact_ncu.txt).
//   a = {COM x, y, z, mass | cell: centre x, y, z; leaf: half edge of the filter cube (-1: single body), -, -}
//   w = kind << 30 | depth << 24 | successor (RECP_END: none)
// A cell's width^2 is not stored: every cell of depth d is the root halved d times, so width^2 = root width^2 / 4^d up to the
// rounding of the split planes.  finalize checks that per node (RECP_TOL); where it holds (RECP_SURE) the walk settles every
// opening decision farther than RECP_BAND from theta with that value, and only the rest - and the cells where it does not hold
// (RECP_EXACT) - fetch the exact width^2 from the breadth-first record rec0[bfs_of[p]].
struct __align__(32) WalkRecP { float a[7]; uint32_t w; };
static_assert(sizeof(WalkRecP) == 32, "walk records are 32 bytes");
constexpr uint32_t RECP_END = 0x00FFFFFFu;                 // (also the largest node count the numbering can hold)
constexpr uint32_t RECP_SINGLE = 0u, RECP_LEAF = 1u, RECP_EXACT = 2u, RECP_SURE = 3u;   // cells have the sign bit set
constexpr float RECP_TOL = 1.0e-4f, RECP_BAND = 1.0e-3f;
enum { RP_W0 = 0, RP_KR = 1, RP_BAD = 2, RP_COUNT = 3, RP_WORDS = 4 };   // recp_par[]: root width^2 (bits), rin2 / width^2, fallback flag, records

// One slot of a level's particle order: everything that moves with a particle from level to level, as ONE 32-byte record
// (= one sector, written with one 256-bit store).  Separate 16 / 8 / 4-byte arrays made every scattered store a partial
// sector write, which the memory system punishes (measured: 87 us for the top-level scatter against 48 us with the final
// slot as one record, profiles/r02_scatter_store_masks.txt).
struct __align__(32) SlotRec {
    float4 pm;                    // {x, y, z, m}
    unsigned long long key;       // path key so far (3 bits per level, digit d at bit 45 - 3d)
    uint32_t idx;                 // particle index
    uint32_t owner;               // node that still has to be split by the grid-wide passes, NODE_END otherwise
};
static_assert(sizeof(SlotRec) == 32, "slot records are one sector");
#ifdef __CUDACC__
__device__ __forceinline__ SlotRec slot_load(const SlotRec* p) {
    uint32_t r[8];
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p) : "memory");
    SlotRec s;
    s.pm = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
    s.key = ((unsigned long long)r[5] << 32) | r[4]; s.idx = r[6]; s.owner = r[7];
    return s;
}
__device__ __forceinline__ void slot_store(SlotRec* p, float4 pm, unsigned long long key, uint32_t idx, uint32_t owner) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(__float_as_uint(pm.x)), "r"(__float_as_uint(pm.y)),
                 "r"(__float_as_uint(pm.z)), "r"(__float_as_uint(pm.w)), "r"((uint32_t)key), "r"((uint32_t)(key >> 32)), "r"(idx), "r"(owner) : "memory");
}
#endif

struct TopRank;
struct TreeDev {
    const float* bounds;          // the root box {lo xyz, hi xyz} in device memory (the build never needs it on the host)
    uint32_t nodes_max;
    sfe_node_t* nodes;            // reference-layout table (tree.h:18-24), materialised on demand by sfe_tree_download()
    uint2* seg;
    uint32_t* next;
    uint32_t* child_base;
    uint32_t* counts;
    uint64_t* keys;               // 48-bit path keys by particle index (filled on demand by sfe_tree_keys_download)
    SlotRec* srec[SFE_MAX_LEVELS + 3];   // the particle order of every level (nodes own contiguous slot ranges)
    uint32_t* perm_out;
    uint8_t* oct;
    uint32_t* rank;
    uint32_t* meta;
    float4* nmid;                 // internal nodes: {mid x, y, z (tree.c:159-163), first slot as bits}
    uint32_t* ticket;             // [2 * levels] tile tickets of rank_kernel / children_kernel
    uint32_t* tdesc;              // look-back descriptors of rank_kernel, tdesc_stride words per level
    unsigned long long* cdesc;    // look-back descriptors of children_kernel, cdesc_stride entries per level
    uint32_t tdesc_stride, cdesc_stride;
    // exact chunked fold of very long nodes (moments_xl_*): per-level work lists and one scratch area per side stream
    uint32_t* xl_count;           // [levels] number of nodes with more than XL_MIN members
    uint32_t* ml_count;           // [levels] non-zero when the level has a node for moments_long_kernel
    uint32_t* xl_nodes;           // [levels][xl_max]
    uint32_t xl_max, xl_slots;    // list capacity per level, chunk slots per scratch area
    uint32_t xl_min;              // nodes with more members than this use the chunked exact fold
    double* xl_sum;               // [SFE_SIDE_STREAMS][xl_slots][8] per chunk: sums of the rounded products, their inclusive prefix
    int* xl_map;                  // [SFE_SIDE_STREAMS][xl_slots][4][8] {a0, a1, min0, max0, min1, max1, exponent, usable}
    uint32_t* xl_flag;            // [SFE_SIDE_STREAMS][xl_slots] look-back state of every chunk (epoch << 2 | state)
    void* xl_list;                // [SFE_SIDE_STREAMS][xl_max + 1] the level's long nodes as a flat list of chunks (XlNode)
    uint32_t* xl_ticket;          // [levels] chunk tickets of xl_prep_kernel
    // lower part of the build (sfe_subtree.cuh): nodes of 9 .. sub_max bodies are finished by one CTA each
    uint32_t sub_max;             // 8 switches the subtree path off (every level goes through the grid-wide kernels)
    uint2* sub_list;              // {node id, level} of every subtree root, in no particular order
    uint32_t* sub_count;
    uint32_t* sub_ticket;
    // upper levels in one pass (sfe_top.cuh)
    uint32_t* top_table;          // [ranges][TOP_COLS] per-range histograms, then exclusive prefixes over the ranges
    uint32_t* top_tot;            // [TOP_COLS] bodies of every cell of the top levels
    uint32_t* top_start;          // [TOP_COLS] first slot of every cell
    uint32_t* top_info;           // [TOP_COLS] node id | split << 31, NODE_END when the cell has no node
    struct TopRank* top_rank;     // [N] path code + stable rank inside every cell on the path
    uint32_t* top_flag;           // [0] non-zero: the one-pass construction does not apply to this build; [1] candidates noted
    uint2* top_cand;              // particles on a split plane: {particle, level}
    volatile uint32_t* summary;   // mapped host memory: {build id, internal nodes, status, node count} per level
    uint32_t build_id;
    WalkRec0* rec0;
    float4* rec1;      // {com x, y, z, mass}
    WalkRec2* rec2;
    WalkRecP* recp;    // pre-order records of the non-empty nodes (walk_p_kernel)
    uint32_t* bfs_of;  // pre-order id -> node id
    uint32_t* recp_par;   // RP_*
    uint32_t* recp_hist;  // [N + 1] per first slot: (nodes starting there) << 16 | sum of their depths; scratch of the numbering
    uint32_t* recp_pref;  // [N + 1] exclusive prefix of the counts = pre-order id of the shallowest node starting at a slot
};

TreeDev tree_dev(Ctx* c);
int walk_publish_recp(Ctx* c, cudaStream_t s, const char* fn);   // sfe_walk.cu: recp_par -> constant memory
int tree_ensure_recp(Ctx* c, cudaStream_t s, const char* fn);  // builds the pre-order records of the current tree if they do not exist yet
int tree_root_moments(Ctx* c, const float4* pm4, uint32_t i0, uint32_t ni, float theta, bool force, cudaStream_t s);

}  // namespace sfe
