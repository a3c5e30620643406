/*
 * sph_shim.c - drop-in replacement for the reference's src/sph.c (sph.h:5-7).
 *   solve_sph(const sim_t*) -> sfe_sph()   density sph.c:32-145, EOS :147-159 (K = 0.1, n = 1), force :169-277
 */
#include "sph.h"
#include "sfe_shim_common.h"

#ifdef ENABLE_SPH
int solve_sph(const sim_t* restrict sim_ptr) {
    const sim_t sim = *sim_ptr;
    sfe_shim_pin(&sim);
    return sfe_sph(sim_find(&sim, SIM_POS), sim_find(&sim, SIM_MAS), sim_find(&sim, SIM_RAD), sim_find(&sim, SIM_RHO),
                   sim_find(&sim, SIM_PRS), sim_find(&sim, SIM_ACC), sim.N,
                   (real)0.100, (real)1.000, (real)(1.000 * 1.000));     /* sph.c:148-149, :231 */
}
#endif
