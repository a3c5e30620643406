/*
 * sf_oracle.c - CPU restatement of Starflood's force-evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see sf_oracle.h).  Parity status: PINNED against the
 * unmodified reference (oracle/_ref/libsfref_*.so) and tests/golden/.
 *
 * Must be compiled WITHOUT floating-point contraction (-ffp-contract=off): the
 * reference is built with -std=c99, which disables FMA contraction, and octant
 * decisions / moment folds / Kahan sums depend on every individual rounding.
 *
 * File:line citations are into /root/reference/src.
 */
#include "sf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SFO_PI_4_3 4.1887902047863909846168578443726705122628925325001410946332594564
#define SFO_PI_8_3 8.3775804095727819692337156887453410245257850650002821892665189128
#define SFO_TAU    6.2831853071795864769252867665590057683943387987502116419498891846
#define SFO_INV_SQRT_PI 0.5641895835477562869480794515607725858440506293289988568440857217
/* common.h:28-29 names this INV_POW_PI_3_2 but the value is pi^(3/2) */
#define SFO_POW_PI_3_2 5.5683279968317078452848179821188357020136243902832439107536758188
/* rng.h:9 */
#define SFO_INV_U32 2.3283062977608182109179324470460414886474609375e-10

/* Compensated accumulator reproducing the reference's volatile Kahan blocks
 * (gravity.c:771-799, sph.c:125-139, :259-271). */
typedef struct { float s, c; } kahan_t;
static inline void kahan_add(kahan_t* k, float x) {
    const float y = x - k->c;
    volatile float t = k->s + y;
    volatile float z = t - k->s;
    k->c = z - y;
    k->s = t;
}

/* ------------------------------------------------------------------ rng.h:13-40 */
void sfo_pcg4d(uint32_t s[4]) {
    uint32_t a = s[0], b = s[1], c = s[2], d = s[3];
    a = a * 1664525u + 1013904223u;
    b = b * 1664525u + 1013904223u;
    c = c * 1664525u + 1013904223u;
    d = d * 1664525u + 1013904223u;
    for (int round = 0; round < 2; round++) {
        a += b * d; b += c * a; c += a * b; d += b * c;
        if (round == 0) { a ^= a >> 16; b ^= b >> 16; c ^= c >> 16; d ^= d >> 16; }
    }
    s[0] = a; s[1] = b; s[2] = c; s[3] = d;
}

static void unit_vector(const double u0, const double u1, double out[3]) {
    /* initcond.c:66-72 : phi = tau*u0, theta = acos(2 u1 - 1) */
    const double phi = SFO_TAU * u0, the = acos(2.0 * u1 - 1.0);
    const double cp = cos(phi), sp = sin(phi), ct = cos(the), st = sin(the);
    out[0] = cp * st; out[1] = sp * st; out[2] = ct;
}

/* ------------------------------------------------------------------ initcond.c:8-263 */
void sfo_initcond(uint32_t N, float* pos, float* vel, float* mas, float* rad, double outflow) {
    const double radius = 1.0;
    const double body_mass = (SFO_PI_4_3 * (radius * radius)) / (double)N; /* initcond.c:26-30 */
    for (uint32_t i = 0; i < N; i++) mas[i] = (float)body_mass;
    if (rad) for (uint32_t i = 0; i < N; i++) rad[i] = (float)0.100;          /* initcond.c:43-45 */

    for (uint32_t i = 0; i < N; i++) {                                        /* initcond.c:48-154 */
        uint32_t s[4] = {0xD88D5EFDu + i, 0x203F83D3u, 0x710BE493u, 0xA294E7F1u};
        sfo_pcg4d(s);
        double u[4], p[3];
        for (int j = 0; j < 4; j++) u[j] = SFO_INV_U32 * (double)s[j];
        unit_vector(u[0], u[1], p);
        for (int k = 0; k < 3; k++) { p[k] *= radius * cbrt(u[2]); pos[3u*i+k] = (float)p[k]; }
    }
    for (uint32_t i = 0; i < N; i++) {                                        /* initcond.c:156-263 */
        const double p[3] = {pos[3u*i+0], pos[3u*i+1], pos[3u*i+2]};
        uint32_t s[4] = {0x4AFDC7CFu + i, 0xAE8D7CF2u, 0xE4827F00u, 0xE44CA389u};
        sfo_pcg4d(s);
        double u[4];
        for (int j = 0; j < 4; j++) u[j] = SFO_INV_U32 * (double)s[j];
        double n[3] = {
            sqrt(-2.0 * log(u[0])) * cos(SFO_TAU * u[1]),
            sqrt(-2.0 * log(u[0])) * sin(SFO_TAU * u[1]),
            sqrt(-2.0 * log(u[2])) * cos(SFO_TAU * u[3])};
        const double sigma = 1.000e-4;
        for (int k = 0; k < 3; k++) n[k] *= sigma;
        const double speed = sqrt((n[0]*n[0]) + (n[1]*n[1]) + (n[2]*n[2]));
        sfo_pcg4d(s);
        for (int j = 0; j < 4; j++) u[j] = SFO_INV_U32 * (double)s[j];
        double v[3];
        unit_vector(u[0], u[1], v);
        for (int k = 0; k < 3; k++) v[k] *= speed;
        if (outflow != 0.0) {
            const double G = 1.000;
            for (int k = 0; k < 3; k++) v[k] += outflow * sqrt(SFO_PI_8_3 * G) * p[k]; /* initcond.c:240-242 */
        }
        for (int k = 0; k < 3; k++) vel[3u*i+k] = (float)v[k];
    }
}

/* SURVEY 8(d) C5 (synthetic; follows the disabled sketch initcond.c:265-329 with the survey's choices). */
void sfo_initcond_galaxy_pair(uint32_t N, float* pos, float* vel, float* mas) {
    sfo_initcond(N, pos, vel, mas, NULL, 4.0e-3);
    const double G = 1.000;
    for (uint32_t i = 0; i < N; i++) {
        double p[3] = {pos[3u*i+0], pos[3u*i+1], pos[3u*i+2]};
        double v[3] = {vel[3u*i+0], vel[3u*i+1], vel[3u*i+2]};
        uint32_t s[4] = {0x5D20CDF7u + i, 0xE5C4E174u, 0xCB77B51Eu, 0xA82C87A3u};
        sfo_pcg4d(s);
        double u[4];
        for (int j = 0; j < 4; j++) u[j] = SFO_INV_U32 * (double)s[j];
        if (u[0] < 0.850) {
            p[1] *= 0.025;
            const double r2 = (p[0]*p[0]) + (p[1]*p[1]) + (p[2]*p[2]);
            const double inv_r1 = 0.0 < r2 ? 1.0 / sqrt(r2) : 0.0;
            v[0] = 1.000e-2 * sqrt(SFO_PI_8_3 * G) * inv_r1 *  p[2];
            v[1] = 0.0;
            v[2] = 1.000e-2 * sqrt(SFO_PI_8_3 * G) * inv_r1 * -p[0];
        }
        const int first = u[1] < 0.500;
        p[0] += first ? 2.000 : -2.000;
        v[0] += first ? -0.005 : 0.005;
        for (int k = 0; k < 3; k++) { pos[3u*i+k] = (float)p[k]; vel[3u*i+k] = (float)v[k]; }
    }
}

/* ------------------------------------------------------------------ simulation.c:245-301 */
void sfo_root_box(uint32_t N, const float* pos, uint32_t step_number, float bounds[6]) {
    double sum = 0.0, com = 0.0;
    for (uint32_t i = 0; i < N; i++) {
        const double x = pos[3u*i+0], y = pos[3u*i+1], z = pos[3u*i+2];
        const double r1 = sqrt((x*x) + (y*y) + (z*z));
        const double yy = r1 - com;
        volatile double t = sum + yy;
        volatile double zz = t - sum;
        com = zz - yy;
        sum = t;
    }
    const double rbar = sum * (1.0 / (double)N);
    uint32_t s[4] = {0x638FE17Bu + step_number, 0xB10BDCE7u, 0x7B8ADB9Bu, 0xD8C5850Eu};
    sfo_pcg4d(s);
    for (int k = 0; k < 3; k++) {
        const double jitter = 2.0 * (SFO_INV_U32 * (double)s[k]) - 1.0;
        bounds[k]     = (float)(-2.000 * rbar + 0.5 * jitter * rbar);
        bounds[3 + k] = (float)( 2.000 * rbar + 0.5 * jitter * rbar);
    }
}

/* One softened monopole interaction, accumulated the reference's way
 * (gravity.c:493-548 and :745-799 are the same arithmetic). */
static inline void grav_term(const float ri[3], const float rj[3], float mj, float eps2,
                             kahan_t a[3], kahan_t* p) {
    const float d[3] = {ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]};
    const float r2 = (d[0]*d[0]) + (d[1]*d[1]) + (d[2]*d[2]);
    const float inv_r2 = 1.0f / (r2 + eps2);
    const float inv_r1 = 1.0f / sqrtf(r2 + eps2);
    const float pj = mj * inv_r1;
    for (int k = 0; k < 3; k++) kahan_add(&a[k], pj * inv_r2 * d[k]);
    kahan_add(p, pj);
}

/* ------------------------------------------------------------------ gravity.c:667-819 */
void sfo_pp(uint32_t N, const float* pos, const float* mas, float G, float eps2,
            uint32_t i0, uint32_t i1, float* acc, float* pot) {
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = i0; i < i1; i++) {
        kahan_t a[3] = {{0, 0}, {0, 0}, {0, 0}}, p = {0, 0};
        const float ri[3] = {pos[3u*i+0], pos[3u*i+1], pos[3u*i+2]};
        for (uint32_t j = 0; j < N; j++) grav_term(ri, &pos[3u*j], mas[j], eps2, a, &p); /* j == i included */
        const float mi = mas[i];
        for (int k = 0; k < 3; k++) acc[3u*i+k] = (-G) * mi * a[k].s;                    /* gravity.c:808-810 */
        pot[i] = (-G) * p.s;
    }
}

void sfo_pp_f64(uint32_t N, const float* pos, const float* mas, double G, double eps2,
                uint32_t i0, uint32_t i1, double* acc, double* pot) {
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = i0; i < i1; i++) {
        double a[3] = {0, 0, 0}, p = 0;
        for (uint32_t j = 0; j < N; j++) {
            const double d[3] = {(double)pos[3u*i+0] - pos[3u*j+0], (double)pos[3u*i+1] - pos[3u*j+1],
                                 (double)pos[3u*i+2] - pos[3u*j+2]};
            const double q = d[0]*d[0] + d[1]*d[1] + d[2]*d[2] + eps2;
            const double ir = 1.0 / sqrt(q), pj = mas[j] * ir;
            for (int k = 0; k < 3; k++) a[k] += pj / q * d[k];
            p += pj;
        }
        for (int k = 0; k < 3; k++) acc[3u*i+k] = -G * mas[i] * a[k];
        pot[i] = -G * p;
    }
}

/* ------------------------------------------------------------------ tree.c:74-318 */
static inline int in_box(const float b[6], const float* p) {
    /* tree.c:142-146, inclusive on both faces */
    return !(b[0] > p[0] || p[0] > b[3] || b[1] > p[1] || p[1] > b[4] || b[2] > p[2] || p[2] > b[5]);
}
static inline void node_clear(sfo_node_t* n, uint32_t parent) {
    memset(n, 0, sizeof *n);
    n->parent = parent;
}
static inline void node_absorb(sfo_node_t* n, const float* p, float m) {
    /* tree.c:166-169 : product rounded, then added, in arrival order */
    n->params[0] += m * p[0];
    n->params[1] += m * p[1];
    n->params[2] += m * p[2];
    n->params[3] += m;
}
static void normalise_moments(sfo_node_t* nodes, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {                                    /* tree.c:302-310 */
        const float M = nodes[i].params[3];
        const float inv = 0.0f < M ? 1.0f / M : 1.0f;
        for (int k = 0; k < 3; k++) nodes[i].params[k] *= inv;
    }
}

int sfo_tree_build_insert(const float bounds[6], uint32_t N, const float* pos, const float* mas,
                          int depth_max, uint32_t max_nodes, sfo_node_t* nodes, uint32_t* num_nodes) {
    uint32_t count = 1;
    node_clear(&nodes[0], 0);
    memcpy(nodes[0].bounds, bounds, 6 * sizeof(float));
    int rc = 0;
    for (uint32_t idx = 0; idx < N && rc == 0; idx++) {
        const float* p = &pos[3u*idx];
        if (!in_box(bounds, p)) continue;
        float lo[3] = {bounds[0], bounds[1], bounds[2]}, hi[3] = {bounds[3], bounds[4], bounds[5]};
        uint32_t cur = 0;
        for (int depth = 0; depth <= depth_max; depth++) {
            sfo_node_t* n = &nodes[cur];
            const uint32_t before = n->bodies++;
            const float mid[3] = {0.5f * (lo[0] + hi[0]), 0.5f * (lo[1] + hi[1]), 0.5f * (lo[2] + hi[2])};
            node_absorb(n, p, mas[idx]);
            if (before < 8u) { n->branch[before] = idx; break; }          /* tree.c:172-175 */
            if (before == 8u) {                                            /* split, tree.c:178-274 */
                if (max_nodes <= count + 8u || depth_max <= depth) { rc = -1; break; }
                sfo_node_t kid[8];
                for (int o = 0; o < 8; o++) {
                    node_clear(&kid[o], cur);
                    for (int k = 0; k < 3; k++) {
                        kid[o].bounds[k]     = (o >> k) & 1 ? mid[k] : lo[k];
                        kid[o].bounds[3 + k] = (o >> k) & 1 ? hi[k] : mid[k];
                    }
                }
                for (int b = 0; b < 8; b++) {                              /* residents: strict test, tree.c:244-247 */
                    const uint32_t r = n->branch[b];
                    const float* q = &pos[3u*r];
                    const int o = (mid[0] < q[0] ? 1 : 0) + (mid[1] < q[1] ? 2 : 0) + (mid[2] < q[2] ? 4 : 0);
                    kid[o].branch[kid[o].bodies++] = r;
                    node_absorb(&kid[o], q, mas[r]);
                }
                for (int o = 0; o < 8; o++) { nodes[count + o] = kid[o]; n->branch[o] = count + o; }
                count += 8u;
            }
            /* arriving particle: non-strict test, tree.c:282-285 */
            const int o = (mid[0] <= p[0] ? 1 : 0) + (mid[1] <= p[1] ? 2 : 0) + (mid[2] <= p[2] ? 4 : 0);
            for (int k = 0; k < 3; k++) { if ((o >> k) & 1) lo[k] = mid[k]; else hi[k] = mid[k]; }
            cur = n->branch[o];
        }
    }
    *num_nodes = count;
    if (rc != 0) return rc;              /* tree.c:180-189 return before the normalisation pass */
    normalise_moments(nodes, count);
    return 0;
}

/* ------------------------------------------------------------------ SURVEY App. B.4 */
int sfo_tree_build_sorted(const float bounds[6], uint32_t N, const float* pos, const float* mas,
                          int depth_max, uint32_t max_nodes, sfo_node_t* nodes, uint32_t* num_nodes,
                          uint64_t* keys, uint32_t* perm, uint32_t* num_inbox) {
    uint32_t* cur = (uint32_t*)malloc(sizeof(uint32_t) * (N ? N : 1));
    uint32_t* nxt = (uint32_t*)malloc(sizeof(uint32_t) * (N ? N : 1));
    uint8_t*  oct = (uint8_t*)malloc(N ? N : 1);
    /* per-level node lists: (node id, segment start, segment end) */
    uint32_t cap = 1024, nlev = 0;
    uint32_t* lev = (uint32_t*)malloc(sizeof(uint32_t) * 3 * cap);
    uint32_t* lev2 = (uint32_t*)malloc(sizeof(uint32_t) * 3 * cap);
    uint32_t cap2 = cap;

    uint32_t M = 0;
    for (uint32_t i = 0; i < N; i++) {
        if (in_box(bounds, &pos[3u*i])) { cur[M++] = i; keys[i] = 0; } else keys[i] = ~(uint64_t)0;
    }
    *num_inbox = M;
    uint32_t count = 1;
    node_clear(&nodes[0], 0);
    memcpy(nodes[0].bounds, bounds, 6 * sizeof(float));
    lev[0] = 0; lev[1] = 0; lev[2] = M; nlev = 1;
    int rc = 0;

    for (int depth = 0; depth <= depth_max && nlev > 0 && rc == 0; depth++) {
        uint32_t nnext = 0;
        for (uint32_t t = 0; t < nlev && rc == 0; t++) {
            const uint32_t id = lev[3*t], s = lev[3*t+1], e = lev[3*t+2];
            sfo_node_t* n = &nodes[id];
            n->bodies = e - s;
            for (uint32_t k = s; k < e; k++) node_absorb(n, &pos[3u*cur[k]], mas[cur[k]]); /* ascending index */
            if (e - s <= 8u) {
                for (uint32_t k = s; k < e; k++) { n->branch[k - s] = cur[k]; nxt[k] = cur[k]; }
                continue;
            }
            if (max_nodes <= count + 8u || depth_max <= depth) { rc = -1; break; }
            const float* b = n->bounds;
            const float mid[3] = {0.5f * (b[0] + b[3]), 0.5f * (b[1] + b[4]), 0.5f * (b[2] + b[5])};
            uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (uint32_t k = s; k < e; k++) {
                const float* q = &pos[3u*cur[k]];
                int o;
                if (k - s < 8u) o = (mid[0] <  q[0] ? 1 : 0) + (mid[1] <  q[1] ? 2 : 0) + (mid[2] <  q[2] ? 4 : 0);
                else            o = (mid[0] <= q[0] ? 1 : 0) + (mid[1] <= q[1] ? 2 : 0) + (mid[2] <= q[2] ? 4 : 0);
                oct[k] = (uint8_t)o; cnt[o]++;
                keys[cur[k]] |= (uint64_t)o << (45 - 3 * depth);
            }
            uint32_t off[8], run = s;
            for (int o = 0; o < 8; o++) { off[o] = run; run += cnt[o]; }
            if (nnext + 8 > cap2) { cap2 *= 2; lev2 = (uint32_t*)realloc(lev2, sizeof(uint32_t) * 3 * cap2); }
            for (int o = 0; o < 8; o++) {
                sfo_node_t* c = &nodes[count + o];
                node_clear(c, id);
                for (int k = 0; k < 3; k++) {
                    c->bounds[k]     = (o >> k) & 1 ? mid[k] : b[k];
                    c->bounds[3 + k] = (o >> k) & 1 ? b[3 + k] : mid[k];
                }
                n->branch[o] = count + o;
                lev2[3*nnext] = count + o; lev2[3*nnext+1] = off[o]; lev2[3*nnext+2] = off[o] + cnt[o];
                nnext++;
            }
            count += 8u;
            for (uint32_t k = s; k < e; k++) nxt[off[oct[k]]++] = cur[k];   /* stable 8-way partition */
        }
        { uint32_t* t = cur; cur = nxt; nxt = t; }
        /* leaf segments were copied through; internal segments were partitioned; both live in `cur` now.
         * nxt still holds the previous level, refresh it so untouched (leaf) slots stay valid next round. */
        memcpy(nxt, cur, sizeof(uint32_t) * M);
        { uint32_t* t = lev; lev = lev2; lev2 = t; uint32_t c = cap; cap = cap2; cap2 = c; }
        nlev = nnext;
    }
    memcpy(perm, cur, sizeof(uint32_t) * M);
    uint32_t tail = M;
    for (uint32_t i = 0; i < N; i++) if (keys[i] == ~(uint64_t)0) perm[tail++] = i;
    normalise_moments(nodes, count);
    *num_nodes = count;
    free(cur); free(nxt); free(oct); free(lev); free(lev2);
    return rc;
}

uint32_t sfo_tree_canonicalise(const sfo_node_t* in, sfo_node_t* out, uint32_t max_out) {
    /* breadth-first renumbering; queue holds old ids, position in the queue is the new id */
    uint32_t* q = (uint32_t*)malloc(sizeof(uint32_t) * max_out);
    uint32_t head = 0, tail = 0;
    q[tail++] = 0;
    out[0] = in[0];
    out[0].parent = 0;
    while (head < tail) {
        const uint32_t nid = head, oid = q[head++];
        if (in[oid].bodies > 8u) {
            if (tail + 8 > max_out) { free(q); return 0; }
            for (int o = 0; o < 8; o++) {
                const uint32_t child_old = in[oid].branch[o];
                out[tail] = in[child_old];
                out[tail].parent = nid;
                out[nid].branch[o] = tail;
                q[tail++] = child_old;
            }
        }
    }
    free(q);
    return tail;
}

/* ------------------------------------------------------------------ gravity.c:245-663 */
void sfo_bh_walk(const sfo_node_t* nodes, uint32_t N, const float* pos, const float* mas,
                 float G, float eps2, float theta, int depth_max,
                 uint32_t i0, uint32_t i1, float* acc, float* pot, sfo_walk_stats_t* stats) {
    double s_iter = 0, s_open = 0, s_eval = 0, s_leaf = 0, s_empty = 0, s_self = 0;
    (void)N;
    #pragma omp parallel for schedule(dynamic, 256) reduction(+: s_iter, s_open, s_eval, s_leaf, s_empty, s_self)
    for (uint32_t i = i0; i < i1; i++) {
        kahan_t a[3] = {{0, 0}, {0, 0}, {0, 0}}, p = {0, 0};
        const float ri[3] = {pos[3u*i+0], pos[3u*i+1], pos[3u*i+2]};
        const float mi = mas[i];
        int path[64];
        for (int d = 0; d < depth_max && d < 64; d++) path[d] = -1;
        uint32_t cur = 0;
        int depth = 0;
        while (depth >= 0) {
            s_iter += 1;
            const sfo_node_t* n = &nodes[cur];
            if (n->bodies == 0u) { s_empty += 1; cur = n->parent; depth--; continue; }   /* :324-328 */
            if (n->bodies > 8u) {
                const int o = path[depth];
                if (o >= 7) { path[depth] = -1; cur = n->parent; depth--; continue; }  /* :343-348 */
                if (o >= 0) { path[depth] = o + 1; cur = n->branch[o + 1]; depth++; continue; } /* :454-460 */
                const float* b = n->bounds;
                const float off[3] = {0.5f * (b[0] + b[3]) - ri[0], 0.5f * (b[1] + b[4]) - ri[1],
                                      0.5f * (b[2] + b[5]) - ri[2]};
                const float d2 = (off[0]*off[0]) + (off[1]*off[1]) + (off[2]*off[2]);
                const float w = (float)(1.0 / 3.0) * ((b[3] - b[0]) + (b[4] - b[1]) + (b[5] - b[2]));
                const float ratio = 0.0f < d2 ? (w * w) / d2 : 0.0f;                     /* :377 */
                if (theta <= ratio) { s_open += 1; path[depth] = 0; cur = n->branch[0]; depth++; continue; }
            } else s_leaf += 1;
            s_eval += 1;
            float rj[3] = {n->params[0], n->params[1], n->params[2]};
            float mj = n->params[3];
            const float* b = n->bounds;
            if (b[0] <= ri[0] && ri[0] <= b[3] && b[1] <= ri[1] && ri[1] <= b[4] &&
                b[2] <= ri[2] && ri[2] <= b[5] && 1u < n->bodies) {                      /* :473-490 */
                s_self += 1;
                for (int k = 0; k < 3; k++) rj[k] *= mj;
                for (int k = 0; k < 3; k++) rj[k] -= mi * ri[k];
                mj -= mi;
                for (int k = 0; k < 3; k++) rj[k] *= 0.0f < mj ? 1.0f / mj : 1.0f;
            }
            grav_term(ri, rj, mj, eps2, a, &p);
            cur = n->parent; depth--;
        }
        for (int k = 0; k < 3; k++) acc[3u*i+k] = (-G) * mi * a[k].s;
        pot[i] = (-G) * p.s;
    }
    if (stats) {
        stats->iterations = s_iter; stats->opened = s_open; stats->evals = s_eval;
        stats->leaf_evals = s_leaf; stats->empty = s_empty; stats->self_sub = s_self;
    }
}

/* ------------------------------------------------------------------ gravity.c:16-111 (SOLVER_VIS rasteriser)
 * One 3-D segment: both ends through the row-major 4x4 matrix (w ignored, no perspective divide), pixel
 * coordinates as gravity.c:61-65, then the all-octant error-term line walk of gravity.c:79-107 with `+=` into the
 * RGBA float buffer for pixels with 0 < x < sizex and 0 < y < sizey. */
static void overlay_segment(float* rb, uint32_t sx_, uint32_t sy_, const float* mvp, const float rgba[4],
                            const float a[3], const float b[3]) {
    int px[2], py[2];
    for (int e = 0; e < 2; e++) {
        const float* q = e ? b : a;
        float h[2];
        for (int r = 0; r < 2; r++)
            h[r] = (mvp[4*r+0] * q[0]) + (mvp[4*r+1] * q[1]) + (mvp[4*r+2] * q[2]) + (mvp[4*r+3] * 1.0f);
        px[e] = (int)(-0.5f + 0.5f * (float)sy_ * h[0] + 0.5f * (float)sx_);
        py[e] = (int)(-0.5f + 0.5f * (float)sy_ * h[1] + 0.5f * (float)sy_);
    }
    const int dx = abs(px[1] - px[0]), dy = -abs(py[1] - py[0]);
    const int stepx = px[0] < px[1] ? 1 : -1, stepy = py[0] < py[1] ? 1 : -1;
    int x = px[0], y = py[0], err = dx + dy;
    for (;;) {
        if (0 < x && x < (int)sx_ && 0 < y && y < (int)sy_)
            for (int c = 0; c < 4; c++) rb[4u * (uint32_t)((int)sx_ * y + x) + c] += rgba[c];
        const int e2 = 2 * err;
        if (e2 >= dy) { if (x == px[1]) break; err += dy; x += stepx; }
        if (e2 <= dx) { if (y == py[1]) break; err += dx; y += stepy; }
    }
}

/* gravity.c:380-445 (cell outline of every internal node when it is first looked at, colour by depth) and
 * gravity.c:550-574 (one segment from the particle to every monopole it evaluates), for particle `idx` only. */
void sfo_bh_overlay(const sfo_node_t* nodes, uint32_t N, const float* pos, const float* mas, float theta, int depth_max,
                    uint32_t idx, const float* mvp, float* rb, uint32_t sizex, uint32_t sizey) {
    if (idx >= N) return;
    const float ri[3] = {pos[3u*idx+0], pos[3u*idx+1], pos[3u*idx+2]};
    const float mi = mas[idx];
    int path[64];
    for (int d = 0; d < depth_max && d < 64; d++) path[d] = -1;
    uint32_t cur = 0;
    int depth = 0;
    while (depth >= 0) {
        const sfo_node_t* n = &nodes[cur];
        const float* b = n->bounds;
        if (n->bodies == 0u) { cur = n->parent; depth--; continue; }
        if (n->bodies > 8u) {
            const int o = path[depth];
            if (o >= 7) { path[depth] = -1; cur = n->parent; depth--; continue; }
            if (o >= 0) { path[depth] = o + 1; cur = n->branch[o + 1]; depth++; continue; }
            const float off[3] = {0.5f * (b[0] + b[3]) - ri[0], 0.5f * (b[1] + b[4]) - ri[1], 0.5f * (b[2] + b[5]) - ri[2]};
            const float d2 = (off[0]*off[0]) + (off[1]*off[1]) + (off[2]*off[2]);
            const float w = (float)(1.0 / 3.0) * ((b[3] - b[0]) + (b[4] - b[1]) + (b[5] - b[2]));
            const float ratio = 0.0f < d2 ? (w * w) / d2 : 0.0f;
            const float f = 1.0f / (float)(depth + 1);
            const float rgba[4] = {f * (0 == depth % 3 ? 0.8f : 0.1f), f * (1 == depth % 3 ? 0.8f : 0.1f),
                                   f * (2 == depth % 3 ? 0.8f : 0.1f), 1.0f};
            for (int ax = 0; ax < 3; ax++) {                 /* four edges along each axis, gravity.c:390-442 */
                const int u = ax == 0 ? 1 : 0, v = ax == 2 ? 1 : 2;
                for (int q = 0; q < 4; q++) {
                    float p0[3], p1[3];
                    p0[u] = p1[u] = (q & 1) ? b[3 + u] : b[u];
                    p0[v] = p1[v] = (q & 2) ? b[3 + v] : b[v];
                    p0[ax] = b[ax]; p1[ax] = b[3 + ax];
                    overlay_segment(rb, sizex, sizey, mvp, rgba, p0, p1);
                }
            }
            if (theta <= ratio) { path[depth] = 0; cur = n->branch[0]; depth++; continue; }
        }
        float rj[3] = {n->params[0], n->params[1], n->params[2]};
        float mj = n->params[3];
        if (b[0] <= ri[0] && ri[0] <= b[3] && b[1] <= ri[1] && ri[1] <= b[4] && b[2] <= ri[2] && ri[2] <= b[5] && 1u < n->bodies) {
            for (int k = 0; k < 3; k++) rj[k] *= mj;
            for (int k = 0; k < 3; k++) rj[k] -= mi * ri[k];
            mj -= mi;
            for (int k = 0; k < 3; k++) rj[k] *= 0.0f < mj ? 1.0f / mj : 1.0f;
        }
        const float magenta[4] = {0.8f, 0.1f, 0.8f, 1.0f};
        overlay_segment(rb, sizex, sizey, mvp, magenta, ri, rj);
        cur = n->parent; depth--;
    }
}

/* ------------------------------------------------------------------ sph.c:32-145 */
void sfo_sph_density(uint32_t N, const float* pos, const float* mas, const float* rad,
                     uint32_t i0, uint32_t i1, float* rho) {
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = i0; i < i1; i++) {
        kahan_t r = {0, 0};
        const float hi = rad[i];
        const float ri[3] = {pos[3u*i+0], pos[3u*i+1], pos[3u*i+2]};
        for (uint32_t j = 0; j < N; j++) {
            const float hj = rad[j];
            const float d[3] = {ri[0] - pos[3u*j+0], ri[1] - pos[3u*j+1], ri[2] - pos[3u*j+2]};
            const float r2 = (d[0]*d[0]) + (d[1]*d[1]) + (d[2]*d[2]);
            const float ai = (float)SFO_INV_SQRT_PI / hi, aj = (float)SFO_INV_SQRT_PI / hj;
            const float Wi = (ai * ai * ai) * expf(-r2 / (hi * hi));
            const float Wj = (aj * aj * aj) * expf(-r2 / (hj * hj));
            kahan_add(&r, 0.5f * (Wi + Wj) * mas[j]);                                    /* sph.c:123 */
        }
        rho[i] = r.s;
    }
}

/* ------------------------------------------------------------------ sph.c:147-159 */
void sfo_sph_eos(uint32_t N, const float* rho, float eos_k, float eos_n, float* prs) {
    const float ex = (float)((1.0 / (double)eos_n) + 1.0);
    /* With the reference's constants the exponent is the literal 2.0f and gcc (-O1 and above, no
     * fast-math needed) folds powf(x, 2.0f) into x*x in the reference binary; glibc's powf differs from
     * that by 1 ulp for ~1 in 2000 inputs.  The oracle follows the compiled reference. */
    if (ex == 2.0f) { for (uint32_t i = 0; i < N; i++) { const float x = eos_k * rho[i]; prs[i] = x * x; } return; }
    for (uint32_t i = 0; i < N; i++) prs[i] = powf(eos_k * rho[i], ex);
}

/* ------------------------------------------------------------------ sph.c:169-277 */
void sfo_sph_force(uint32_t N, const float* pos, const float* mas, const float* rad,
                   const float* rho, const float* prs, float cutoff2,
                   uint32_t i0, uint32_t i1, float* acc) {
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = i0; i < i1; i++) {
        kahan_t a[3] = {{0, 0}, {0, 0}, {0, 0}};
        const float hi = rad[i];
        const float ri[3] = {pos[3u*i+0], pos[3u*i+1], pos[3u*i+2]};
        const float rhoi2 = rho[i] * rho[i];
        const float ci = 0.0f < rhoi2 ? prs[i] / rhoi2 : 0.0f;
        for (uint32_t j = 0; j < N; j++) {
            const float hj = rad[j];
            const float d[3] = {ri[0] - pos[3u*j+0], ri[1] - pos[3u*j+1], ri[2] - pos[3u*j+2]};
            const float rhoj2 = rho[j] * rho[j];
            const float cj = 0.0f < rhoj2 ? prs[j] / rhoj2 : 0.0f;
            const float r2 = (d[0]*d[0]) + (d[1]*d[1]) + (d[2]*d[2]);
            if (r2 > cutoff2) continue;                                                   /* sph.c:231 */
            const float ni = -2.0f * expf(-r2 / (hi * hi)) / ((float)SFO_POW_PI_3_2 * hi * hi * hi * hi * hi);
            const float nj = -2.0f * expf(-r2 / (hj * hj)) / ((float)SFO_POW_PI_3_2 * hj * hj * hj * hj * hj);
            for (int k = 0; k < 3; k++) kahan_add(&a[k], (ci + cj) * (0.5f * (ni + nj) * d[k]));
        }
        const float mi = mas[i];
        for (int k = 0; k < 3; k++) acc[3u*i+k] += -mi * a[k].s;                         /* sph.c:274-276 */
    }
}

/* ------------------------------------------------------------------ simulation.c:381-451 */
void sfo_kick(uint32_t N, float* vel, const float* acc, float dt) {
    for (uint32_t i = 0; i < 3u * N; i++) vel[i] += 0.5f * dt * acc[i];
}
void sfo_drift(uint32_t N, float* pos, const float* vel, float dt) {
    for (uint32_t i = 0; i < 3u * N; i++) pos[i] += dt * vel[i];
}
void sfo_energies(uint32_t N, const float* vel, const float* mas, const float* pot,
                  float* ken, float* pen, double* T, double* V) {
    double ts = 0, tc = 0, vs = 0, vc = 0;
    for (uint32_t i = 0; i < N; i++) {
        const float v2 = (vel[3u*i]*vel[3u*i]) + (vel[3u*i+1]*vel[3u*i+1]) + (vel[3u*i+2]*vel[3u*i+2]);
        ken[i] = 0.5f * mas[i] * v2;
        pen[i] = mas[i] * pot[i];
    }
    for (uint32_t i = 0; i < N; i++) {
        const double y = (double)ken[i] - tc; volatile double t = ts + y; volatile double z = t - ts; tc = z - y; ts = t;
    }
    for (uint32_t i = 0; i < N; i++) {
        const double y = (double)pen[i] - vc; volatile double t = vs + y; volatile double z = t - vs; vc = z - y; vs = t;
    }
    *T = ts; *V = vs;
}

/* ================================================================== visualization.c:253-739 (vis_draw)
 * SURVEY 8(f) row N4, second half: the particle splat.  SHUTTER_SPEED is off in the reference (config.h:215), so
 * time_offset is 0 and the velocity term of visualization.c:593-606 is `p += 0 * v`. */

/* visualization.c:274-307: mean |pos| in double, serial Kahan */
double sfo_vis_mean_radius(uint32_t N, const float* pos) {
    double sum = 0.0, com = 0.0;
    for (uint32_t i = 0; i < N; i++) {
        const double x = pos[3u*i+0], y = pos[3u*i+1], z = pos[3u*i+2];
        const double r1 = sqrt((x*x) + (y*y) + (z*z));
        const double yy = r1 - com;
        volatile double t = sum + yy;
        volatile double zz = t - sum;
        com = zz - yy;
        sum = t;
    }
    return sum * (1.0 / (double)N);
}

static void mat4_mul(const float* A, const float* B, float* C) {   /* visualization.c:432-446 (term order 3,2,1,0) */
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            C[4*i+j] = (A[4*i+3] * B[4*3+j]) + (A[4*i+2] * B[4*2+j]) + (A[4*i+1] * B[4*1+j]) + (A[4*i+0] * B[4*0+j]);
}

/* visualization.c:327,344-489: model (rotation), view (identity), projection (ortho_scale on the 3x3 part).
 * ortho_exp2 = ORTHO_SCALE (config.h:209, -1.0), autoscale = AUTOSCALING (config.h:212), dt = TIMESTEP_SIZE. */
void sfo_vis_matrices(double average_radius, uint32_t step_number, double dt, double ortho_exp2, int autoscale,
                      float mat_vm[16], float mvp[16]) {
    const float time = (float)(dt * (double)step_number);
    float M[16], V[16], P[16], VM[16];
    for (int i = 0; i < 16; i++) M[i] = V[i] = P[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    float alp = 0.0f, bet = (float)SFO_TAU * 0.1250f, gam = (float)SFO_TAU * 0.0625f;
    alp += (float)SFO_TAU * 1.000e-4f * time;
    bet += (float)SFO_TAU * 1.000e-4f * time;
    gam += (float)SFO_TAU * 1.000e-4f * time;
    const float ca = cosf(alp), sa = sinf(alp), cb = cosf(bet), sb = sinf(bet), cg = cosf(gam), sg = sinf(gam);
    const float rot[16] = {
        ca*cb, ca*sb*sg - sa*cg, ca*sb*cg + sa*sg, 0.0f,
        sa*cb, sa*sb*sg + ca*cg, sa*sb*cg - ca*sg, 0.0f,
          -sb,            cb*sg,            cb*cg, 0.0f,
         0.0f,             0.0f,             0.0f, 1.0f};
    for (int i = 0; i < 16; i++) M[i] = rot[i];
    float ortho = (float)exp2(ortho_exp2);
    if (autoscale) ortho *= (float)(1.0 / average_radius);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) P[4*i+j] *= ortho;
    mat4_mul(V, M, VM);
    mat4_mul(P, VM, mvp);
    if (mat_vm) memcpy(mat_vm, VM, sizeof VM);
}

/* visualization.c:334-342 (clear) + :503-683 (atomic accumulation rasterisation).  Integer adds commute, so the
 * buffer does not depend on the thread schedule. */
void sfo_vis_splat(uint32_t N, const float* pos, const float* pot, uint32_t step_number, const float mvp[16],
                   uint32_t sizex, uint32_t sizey, uint32_t samples, int32_t* atomic_buffer) {
    memset(atomic_buffer, 0, sizeof(int32_t) * 4u * (size_t)sizex * sizey);
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t idx = 0; idx < N; idx++) {
        const float x[4] = {pos[3u*idx+0], pos[3u*idx+1], pos[3u*idx+2], 1.0f};
        float p[4];
        for (int i = 0; i < 4; i++) p[i] = (mvp[4*i+0]*x[0]) + (mvp[4*i+1]*x[1]) + (mvp[4*i+2]*x[2]) + (mvp[4*i+3]*x[3]);
        const float phase = 1.000e-1f * pot[idx];
        const float color[3] = {
            0.5f * cosf((float)SFO_TAU * phase - (float)(SFO_TAU * (0.0 / 3.0))) + 0.5f,
            0.5f * cosf((float)SFO_TAU * phase - (float)(SFO_TAU * (1.0 / 3.0))) + 0.5f,
            0.5f * cosf((float)SFO_TAU * phase - (float)(SFO_TAU * (2.0 / 3.0))) + 0.5f};
        const int32_t ac[4] = {(int32_t)(100.0f * color[0]), (int32_t)(100.0f * color[1]), (int32_t)(100.0f * color[2]), 1};
        for (uint32_t samp = 0; samp < samples; samp++) {
            uint32_t s[4] = {0xB79ABC95u + samp, 0xE0DA3F84u + idx, 0xAB75F07Bu + step_number, 0xCF52CA01u};
            sfo_pcg4d(s);
            const float r2 = (float)SFO_INV_U32 * (float)s[2], r3 = (float)SFO_INV_U32 * (float)s[3];
            const float n0 = sqrtf(-2.0f * logf(r2)) * cosf((float)SFO_TAU * r3);
            const float n1 = sqrtf(-2.0f * logf(r2)) * sinf((float)SFO_TAU * r3);
            const float so0 = 0.750f * n0, so1 = 0.750f * n1;
            const float fx = so0 + -0.5f + 0.5f * (float)sizey * p[0] + 0.5f * (float)sizex;
            const float fy = so1 + -0.5f + 0.5f * (float)sizey * p[1] + 0.5f * (float)sizey;
            /* (int) of a NaN / out-of-range float is INT_MIN on x86-64 (cvttss2si) and so fails the bounds test */
            if (!(fx > -2147483648.0f && fx < 2147483648.0f && fy > -2147483648.0f && fy < 2147483648.0f)) continue;
            const int cx = (int)fx, cy = (int)fy;
            if (0 <= cx && cx < (int)sizex && 0 <= cy && cy < (int)sizey) {
                const uint32_t pix = sizex * (uint32_t)cy + (uint32_t)cx;
                for (int i = 0; i < 4; i++) {
                    #pragma omp atomic update
                    atomic_buffer[4u*pix+i] += ac[i];
                }
            }
        }
    }
}

/* visualization.c:691-735: pixel_value_scale = 0.01 / SPATIAL_SAMPLES * exp2(EXPOSURE); tanh tonemap when
 * VIS_FILE_FORMAT > 0; alpha = 1 */
void sfo_vis_finish(const int32_t* atomic_buffer, uint32_t sizex, uint32_t sizey, double pixel_value_scale, int tonemap,
                    float* render_buffer) {
    for (size_t i = 0; i < (size_t)sizex * sizey; i++) {
        double px[4];
        for (int j = 0; j < 4; j++) px[j] = pixel_value_scale * (double)atomic_buffer[4u*i+j];
        if (tonemap) for (int j = 0; j < 4; j++) px[j] = tanh(px[j]);
        px[3] = 1.000;
        for (int j = 0; j < 4; j++) render_buffer[4u*i+j] = (float)px[j];
    }
}
