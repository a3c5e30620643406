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
    uint32_t child_kind;          // kind << 30 | first NON-EMPTY child (internal nodes)
    float rin2;                   // beyond this squared distance (internal: from the centre; leaf: from the COM) a point
                                  // is surely outside the box: filter in front of the exact bounds test (gravity.c:473-476)
    float pad;
};
struct __align__(16) WalkRec2 {   // 32 B: the cell itself - bounds (gravity.c:473-476 self-subtraction test, tree.c:219-226) + parent
    float lo[3], hi[3];
    uint32_t parent;
    uint32_t pad;
};
static_assert(sizeof(WalkRec0) == 32 && sizeof(WalkRec2) == 32, "walk records are 32 bytes");

struct TreeDev {
    float bounds[6];
    uint32_t nodes_max;
    sfe_node_t* nodes;            // reference-layout table (tree.h:18-24), materialised on demand by sfe_tree_download()
    uint2* seg;
    uint32_t* next;
    uint32_t* child_base;
    uint32_t* counts;
    uint64_t* keys;               // 48-bit path keys by particle index (filled on demand by sfe_tree_keys_download)
    uint64_t* skey[2];            // path key of every slot, ping-pong between levels (level d lives in skey[d & 1])
    uint32_t* perm[SFE_MAX_LEVELS + 3];
    float4* spm4[SFE_MAX_LEVELS + 3];   // slot-ordered copy of pm4 per level
    uint32_t* perm0;
    uint32_t* perm_out;
    uint32_t* slot_node[2];
    uint32_t* slot_node0;
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
    double* xl_sum;               // [SFE_SIDE_STREAMS][xl_slots][4] chunk sums, then exclusive prefixes, of the rounded products
    long long* xl_map;            // [SFE_SIDE_STREAMS][xl_slots][4][8] {a0, a1, min0, max0, min1, max1, exponent, usable}
    // lower part of the build (sfe_subtree.cuh): nodes of 9 .. sub_max bodies are finished by one CTA each
    uint32_t sub_max;             // 8 switches the subtree path off (every level goes through the grid-wide kernels)
    uint2* sub_list;              // {node id, level} of every subtree root, in no particular order
    uint32_t* sub_count;
    uint32_t* sub_ticket;
    volatile uint32_t* summary;   // mapped host memory: {build id, internal nodes, status, node count} per level
    uint32_t build_id;
    WalkRec0* rec0;
    float4* rec1;      // {com x, y, z, mass}
    WalkRec2* rec2;
};

TreeDev tree_dev(Ctx* c);
int tree_root_moments(Ctx* c, const float4* pm4, uint32_t i0, uint32_t ni, float theta, bool force, cudaStream_t s);

}  // namespace sfe
