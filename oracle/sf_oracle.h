/*
 * sf_oracle.h - CPU restatement of Starflood's force-evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA engine in
 * starflood_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path never does.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit (integer and
 * fp32 outputs alike) against the unmodified reference compiled from
 * /root/reference/src into oracle/_ref/libsfref_*.so (tests/test_oracle_vs_ref.py)
 * and against the golden vectors generated from those libraries
 * (tests/golden/, tests/golden/make_golden.py).  The Morton keys / permutation
 * have no counterpart in the reference (it inserts particles one by one); they are
 * pinned indirectly: the node table derived from them is bit-identical to the
 * reference's tree_build() output after renumbering (sfo_tree_canonicalise()).
 *
 * File:line citations are into /root/reference/src.
 */
#ifndef SF_ORACLE_H
#define SF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same memory layout as the reference's node_t (tree.h:18-24), 80 bytes. */
typedef struct {
    uint32_t parent;
    uint32_t bodies;
    uint32_t branch[8];
    float    bounds[6]; /* min xyz, max xyz */
    float    params[4]; /* COM xyz, mass */
} sfo_node_t;

typedef struct {
    double iterations;   /* loop iterations of the reference walk, summed over particles */
    double opened;       /* internal nodes opened */
    double evals;        /* monopole evaluations (leaves + accepted internal nodes) */
    double leaf_evals;   /* of which leaves */
    double empty;        /* visits to empty children */
    double self_sub;     /* self-subtractions */
} sfo_walk_stats_t;

/* rng.h:13-40 */
void sfo_pcg4d(uint32_t s[4]);

/* initcond.c:8-263 (uniform unit ball, Maxwellian + Hubble-like outflow).
 * outflow = 4.0e-3 reproduces the reference; 0 gives SURVEY C3's "thermal only" variant.
 * rad may be NULL. */
void sfo_initcond(uint32_t N, float* pos, float* vel, float* mas, float* rad, double outflow);

/* synthetic galaxy pair following the author's disabled sketch (initcond.c:265-329). */
void sfo_initcond_galaxy_pair(uint32_t N, float* pos, float* vel, float* mas);

/* simulation.c:245-301 : jittered root box from mean |pos| */
void sfo_root_box(uint32_t N, const float* pos, uint32_t step_number, float bounds[6]);

/* gravity.c:667-819 : all-pairs softened gravity for targets [i0,i1) over all N sources */
void sfo_pp(uint32_t N, const float* pos, const float* mas, float G, float eps2,
            uint32_t i0, uint32_t i1, float* acc, float* pot);
/* float64 evaluation of the same formula (second opinion) */
void sfo_pp_f64(uint32_t N, const float* pos, const float* mas, double G, double eps2,
                uint32_t i0, uint32_t i1, double* acc, double* pot);

/* tree.c:74-318 : sequential insertion build.  Returns 0 on success, -1 on node / depth overflow
 * (tree.c:180-189).  nodes must hold max_nodes records; *num_nodes receives the node count. */
int sfo_tree_build_insert(const float bounds[6], uint32_t N, const float* pos, const float* mas,
                          int depth_max, uint32_t max_nodes, sfo_node_t* nodes, uint32_t* num_nodes);

/* Sort-based construction that the CUDA engine follows:
 * keys[i]  48-bit octal path key of particle i down to its leaf, zero padded (digit of level l at bits
 *          [45-3l,47-3l]); 0xFFFFFFFFFFFFFFFF for particles outside the root box;
 * perm[k]  particle index at slot k of the (key, index)-sorted order (out-of-box ones last, by index);
 * nodes    canonical table: node 0 = root, children of internal nodes allocated 8 at a time in
 *          breadth-first order (level by level, parents in key order).
 * Returns 0, or -1 on node / depth overflow. */
int sfo_tree_build_sorted(const float bounds[6], uint32_t N, const float* pos, const float* mas,
                          int depth_max, uint32_t max_nodes, sfo_node_t* nodes, uint32_t* num_nodes,
                          uint64_t* keys, uint32_t* perm, uint32_t* num_inbox);

/* Renumber any reference-layout table (e.g. the reference's own) into the canonical numbering.
 * Returns the node count (1 + 8 * #internal). */
uint32_t sfo_tree_canonicalise(const sfo_node_t* in, sfo_node_t* out, uint32_t max_out);

/* gravity.c:245-663 : Barnes-Hut walk for targets [i0,i1) */
void sfo_bh_walk(const sfo_node_t* nodes, uint32_t N, const float* pos, const float* mas,
                 float G, float eps2, float theta, int depth_max,
                 uint32_t i0, uint32_t i1, float* acc, float* pot, sfo_walk_stats_t* stats);

/* gravity.c:16-111,380-445,550-574 : the SOLVER_VIS debug overlay of particle `idx` - cell outlines and interaction
 * segments rasterised with `+=` into an RGBA float buffer through the row-major 4x4 matrix mvp (matMVP) */
void sfo_bh_overlay(const sfo_node_t* nodes, uint32_t N, const float* pos, const float* mas, float theta, int depth_max,
                    uint32_t idx, const float* mvp, float* rb, uint32_t sizex, uint32_t sizey);

/* sph.c:32-145 density, :147-159 EOS, :169-277 pressure force (acc +=) for targets [i0,i1) */
void sfo_sph_density(uint32_t N, const float* pos, const float* mas, const float* rad,
                     uint32_t i0, uint32_t i1, float* rho);
void sfo_sph_eos(uint32_t N, const float* rho, float eos_k, float eos_n, float* prs);
void sfo_sph_force(uint32_t N, const float* pos, const float* mas, const float* rad,
                   const float* rho, const float* prs, float cutoff2,
                   uint32_t i0, uint32_t i1, float* acc);

/* simulation.c:381-451 : kick / drift / per-particle energies, and the double-Kahan totals :454-497 */
void sfo_kick(uint32_t N, float* vel, const float* acc, float dt);
void sfo_drift(uint32_t N, float* pos, const float* vel, float dt);
void sfo_energies(uint32_t N, const float* vel, const float* mas, const float* pot,
                  float* ken, float* pen, double* T, double* V);

/* visualization.c:253-739 (vis_draw), SURVEY 8(f) row N4: mean radius, MVP matrices, particle splat into the integer
 * accumulation buffer, finishing pass into the float render buffer */
double sfo_vis_mean_radius(uint32_t N, const float* pos);
void sfo_vis_matrices(double average_radius, uint32_t step_number, double dt, double ortho_exp2, int autoscale,
                      float mat_vm[16], float mvp[16]);
void sfo_vis_splat(uint32_t N, const float* pos, const float* pot, uint32_t step_number, const float mvp[16],
                   uint32_t sizex, uint32_t sizey, uint32_t samples, int32_t* atomic_buffer);
void sfo_vis_finish(const int32_t* atomic_buffer, uint32_t sizex, uint32_t sizey, double pixel_value_scale, int tonemap,
                    float* render_buffer);

#ifdef __cplusplus
}
#endif
#endif
