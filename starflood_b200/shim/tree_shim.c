/*
 * tree_shim.c - drop-in replacement for the reference's src/tree.c (tree.h:38-44).
 *
 *   tree_init(tree_t*, max_nodes)   tree.c:14-55   keeps max_nodes; the node table itself lives on the device, so no
 *                                                  80 B * max_nodes host block is allocated (its only consumer,
 *                                                  solve_gravity_part_tree, is replaced too; SURVEY 8b "Ownership")
 *   tree_free(tree_t*)              tree.c:57-72
 *   tree_build(tree_t*, N, pos, mas) tree.c:74-318 -> sfe_tree_build(); like the reference it does not write
 *                                                  num_nodes back (tree.c:77 works on a local copy)
 */
#include "tree.h"
#include <stddef.h>
#include "sfe_shim_common.h"

#ifdef ENABLE_TREE
int tree_init(tree_t* restrict tree_ptr, unsigned int max_nodes) {
    tree_ptr->max_nodes = (uint)max_nodes;
    tree_ptr->num_nodes = (uint)0u;
    for (int i = 0; i < 6; i++) tree_ptr->bounds[i] = (real)0;
    tree_ptr->node = (node_t*)NULL;
    return sfe_init(-1);
}

int tree_free(tree_t* restrict tree_ptr) {
    tree_ptr->max_nodes = (uint)0u;
    tree_ptr->num_nodes = (uint)0u;
    for (int i = 0; i < 6; i++) tree_ptr->bounds[i] = (real)0;
    tree_ptr->node = (node_t*)NULL;
    return STARFLOOD_SUCCESS;
}

int tree_build(tree_t* restrict tree_ptr, unsigned int num_parts, const real* restrict pos, const real* restrict mas) {
    /* tree_build is the first engine call of a run (main.c:398 -> sim_solv) and does not see sim_t: pin what it does see -
     * the span of sim.mem from pos to mas - so that the very first upload already runs at PCIe rate; the gravity shim's
     * registration of the whole block replaces this one later */
    {
        const char* a = (const char*)pos < (const char*)mas ? (const char*)pos : (const char*)mas;
        const char* pe = (const char*)(pos + 3u * (size_t)num_parts); const char* me = (const char*)(mas + (size_t)num_parts);
        const char* b = pe > me ? pe : me;
        if (num_parts) sfe_host_register((void*)a, (uint64_t)(b - a));
    }
    return sfe_tree_build(tree_ptr->bounds, pos, mas, num_parts, TREE_DEPTH_MAX, tree_ptr->max_nodes);
}
#endif
