/*
 * gravity_shim.c - drop-in replacement for the reference's src/gravity.c.
 *
 * Provides, with the reference's exact signatures (gravity.h:12-22):
 *   solve_gravity_part_part(const sim_t*)                         -> sfe_pp()      (gravity.c:667-819)
 *   solve_gravity_part_tree(const sim_t*[, render_buffer, w, h])  -> sfe_bh()      (gravity.c:245-663)
 * With SOLVER_VIS defined (config.h:32, the default) the walk also draws the debug overlay of particle SOLVER_VIS into
 * the render buffer (gravity.c:16-111, :380-445, :550-574)                         -> sfe_bh_overlay()
 * through visualization.c's matMVP, exactly as gravity.c:17 binds it.
 */
#include "gravity.h"
#include "sfe_shim_common.h"

int solve_gravity_part_part(const sim_t* restrict sim_ptr) {
    const sim_t sim = *sim_ptr;
    sfe_shim_pin(&sim);
    return sfe_pp(sim_find(&sim, SIM_POS), sim_find(&sim, SIM_MAS), sim_find(&sim, SIM_ACC), sim_find(&sim, SIM_POT),
                  sim.N, (real)G, (real)(EPSILON * EPSILON));               /* gravity.c:754 */
}

#ifdef ENABLE_TREE
#ifdef SOLVER_VIS
extern real matMVP[16];   /* visualization.c:251 */
int solve_gravity_part_tree(const sim_t* restrict sim_ptr, f32* render_buffer, unsigned int sizex, unsigned int sizey)
#else
int solve_gravity_part_tree(const sim_t* restrict sim_ptr)
#endif
{
    const sim_t sim = *sim_ptr;
    sfe_shim_pin(&sim);
    /* EXACT: IEEE div/sqrt and the reference's Kahan order -> bit-identical acc/pot (ENABLE_KAHAN_SUMMATION_* on,
     * the shipped default, config.h:194-198).  Define SFE_SHIM_WALK_FAST to trade that for ~15 % walk time. */
#ifdef SFE_SHIM_WALK_FAST
    const uint32_t flavour = SFE_WALK_FAST;
#else
    const uint32_t flavour = SFE_WALK_EXACT;
#endif
    const int rc = sfe_bh(sim_find(&sim, SIM_POS), sim_find(&sim, SIM_MAS), sim_find(&sim, SIM_ACC), sim_find(&sim, SIM_POT),
                          sim.N, (real)G, (real)(EPSILON * EPSILON), (real)TREE_THETA, flavour);   /* gravity.c:448 */
#ifdef SOLVER_VIS
    if (rc == STARFLOOD_SUCCESS && render_buffer != NULL)
        return sfe_bh_overlay((unsigned int)SOLVER_VIS, (real)TREE_THETA, matMVP, render_buffer, sizex, sizey);
#endif
    return rc;
}
#endif
