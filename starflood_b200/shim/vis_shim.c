/*
 * vis_shim.c - drop-in replacement for vis_draw() (visualization.h:40, visualization.c:253-739).
 *
 * visualization.c stays in the build for vis_init() / vis_free() and the matrix globals; only its vis_draw() is renamed
 * out of the way at compile time (-Dvis_draw=vis_draw_reference on that one file, see build_dropin.sh / INTEGRATION.md).
 * This file keeps the host side of vis_draw() - the model / view matrices, which are a few libm calls on constants
 * (visualization.c:327-446) - and hands the O(N * SPATIAL_SAMPLES) part to the engine: mean radius, projection matrix,
 * clear, particle splat and finishing pass                                                   -> sfe_vis_draw()
 */
#include "visualization.h"
#include "sfe_shim_common.h"
#include <math.h>

#ifdef SHUTTER_SPEED
#error "the B200 engine renders without motion blur: SHUTTER_SPEED (config.h:215, off by default) is not supported"
#endif
#ifndef SPATIAL_SAMPLES
#error "SPATIAL_SAMPLES (config.h:203) must be defined"
#endif

extern real matM[16], matV[16], matP[16], matMVP[16];   /* visualization.c:239-251 */

int vis_draw(const vis_t* vis_ptr, const sim_t* restrict sim_ptr) {
    const sim_t sim = *sim_ptr;
    const vis_t vis = *vis_ptr;
    sfe_shim_pin(&sim);
    {   /* the render buffer crosses PCIe every frame: pin the block vis_init() allocated (visualization.c:94-98) */
        static void* pinned = NULL;
        if (pinned != vis.mem) {
            if (pinned) sfe_host_unregister(pinned);
            const uint64_t px = (uint64_t)vis.sizex * (uint64_t)vis.sizey;
            if (SFE_SUCCESS == sfe_host_register(vis.mem, px * 4u * (sizeof(i32) + sizeof(f32)))) pinned = vis.mem;
        }
    }
    const real time = (real)((double)TIMESTEP_SIZE * (double)sim.step_number);       /* visualization.c:327-331 */

    /* model matrix = rotation (visualization.c:368-415), view = identity, matVM = matV * matM (:432-446) */
    real alp = (real)0.0, bet = (real)TAU * (real)0.1250, gam = (real)TAU * (real)0.0625;
    alp += (real)TAU * (real)1.000e-4 * time;
    bet += (real)TAU * (real)1.000e-4 * time;
    gam += (real)TAU * (real)1.000e-4 * time;
    const real ca = real_cos(alp), sa = real_sin(alp), cb = real_cos(bet), sb = real_sin(bet), cg = real_cos(gam), sg = real_sin(gam);
    const real rot[16] = {
        (real)(ca*cb), (real)(ca*sb*sg-sa*cg), (real)(ca*sb*cg+sa*sg), (real)(0.0),
        (real)(sa*cb), (real)(sa*sb*sg+ca*cg), (real)(sa*sb*cg-ca*sg), (real)(0.0),
        (real)(  -sb), (real)(         cb*sg), (real)(         cb*cg), (real)(0.0),
        (real)(  0.0), (real)(           0.0), (real)(           0.0), (real)(1.0)};
    real matVM[16];
    for (int i = 0; i < 16; i++) { matM[i] = rot[i]; matV[i] = (i % 5 == 0) ? (real)1 : (real)0; }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            matVM[4*i+j] = (matV[4*i+3] * matM[4*3+j]) + (matV[4*i+2] * matM[4*2+j]) + (matV[4*i+1] * matM[4*1+j]) + (matV[4*i+0] * matM[4*0+j]);

#ifdef ORTHO_SCALE
    const real ortho_scale = (real)exp2(ORTHO_SCALE);                               /* visualization.c:418-422 */
#else
    const real ortho_scale = (real)1.0;
#endif
#ifdef AUTOSCALING
    const int autoscale = 1;                                                        /* visualization.c:424-426 */
#else
    const int autoscale = 0;
#endif
    double pixel_value_scale = 0.01 / (double)SPATIAL_SAMPLES;                      /* visualization.c:691-695 */
#ifdef EXPOSURE
    pixel_value_scale *= exp2(EXPOSURE);
#endif
#if (0 < VIS_FILE_FORMAT)
    const int tonemap = 1;                                                          /* visualization.c:719-724 */
#else
    const int tonemap = 0;
#endif
#ifdef SFE_SHIM_VIS_KEEP_ATOMIC
    i32* atomic_out = vis.atomic_buffer;   /* nothing outside vis_draw() reads the accumulation buffer; skip its copy by default */
#else
    i32* atomic_out = NULL;
#endif
    double average_radius = 0.0;
    const int rc = sfe_vis_draw(sim_find(&sim, SIM_POS), sim_find(&sim, SIM_POT), sim.N, sim.step_number, matVM, ortho_scale,
                                autoscale, vis.sizex, vis.sizey, (unsigned int)SPATIAL_SAMPLES, pixel_value_scale, tonemap,
                                atomic_out, vis.render_buffer, &average_radius, matMVP);
    if (rc != SFE_SUCCESS) return STARFLOOD_FAILURE;
    printf("average_radius: % .15f\n", average_radius);                             /* visualization.c:306 */
    /* matP as the reference leaves it (visualization.c:428-432): scaled 3x3 block */
    real scale = ortho_scale;
    if (autoscale) scale *= (real)(1.0 / average_radius);
    for (int i = 0; i < 16; i++) matP[i] = (i % 5 == 0) ? (real)1 : (real)0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) matP[4*i+j] *= scale;
    return STARFLOOD_SUCCESS;
}
