/*
 * sfe_shim_common.h - glue shared by the three reference-side shims.
 *
 * The shims are compiled WITH the reference's own headers (config.h, simulation.h, tree.h ...) in place of
 * gravity.c / tree.c / sph.c (Makefile:125 SRCS), so sim_t / SIM_DOF / real follow the same config.h as the
 * host objects (SURVEY 8b "compile-time coupling").  They unpack sim_t and forward plain pointers to the
 * config-independent C ABI in include/sfengine.h.
 */
#ifndef SFE_SHIM_COMMON_H
#define SFE_SHIM_COMMON_H
#include <stdio.h>
#include "common.h"
#include "config.h"
#include "simulation.h"
#include "types.h"
#include "sfengine.h"

#ifdef STARFLOOD_DOUBLE_PRECISION
#error "the B200 engine is fp32 only (STARFLOOD_DOUBLE_PRECISION, config.h:221, is out of scope)"
#endif
#ifndef EPSILON
#error "the B200 engine implements the softened kernel only: EPSILON (config.h:175) must be defined"
#endif

/* Pin the whole sim.mem block once (simulation.c:89 allocates it with posix_memalign, pageable) so that the
 * per-step copies run at PCIe rate.  Re-registers if the block moves (it never does in the reference). */
static inline void sfe_shim_pin(const sim_t* sim) {
    static void* pinned = NULL;
    if (pinned != sim->mem) {
        if (pinned) sfe_host_unregister(pinned);
        if (SFE_SUCCESS == sfe_host_register(sim->mem, (uint64_t)sizeof(real) * (uint64_t)sim->N * (uint64_t)SIM_DOF))
            pinned = sim->mem;
    }
}
#endif
