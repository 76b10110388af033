/*
 * CPU ORACLE (test infrastructure, NOT product code): C99 + OpenMP twin of oracle/d3q19_ref.py:step_physical.
 *
 * compat = physical is this repository's own consistent-lattice step (the reference has none), so the NumPy function
 * DEFINES its arithmetic: every operation is an explicit IEEE f32 add / sub / mul / fused multiply-add, a correctly
 * rounded reciprocal or square root (csrc/lbm_phys.cuh mirrors it operation by operation and is bit-exact against it).
 * NumPy needs minutes per step beyond 64^3; this file states the same operations cell by cell so that the CUDA kernels
 * can be checked at the sizes BASELINE.json names (256^3 periodic, V60 512^3): tests/test_gpu_parity_at_scale.py.
 * It is pinned to the NumPy function bit for bit on small boxes with every feature combination
 * (tests/test_oracle_phys_c_vs_numpy.py); the NumPy function in turn is what the CUDA kernels are tested against.
 *
 * Layout-agnostic: a cell (x, y, z) lives at x*sx + y*sy + z*sz of a scalar volume, population q and vector component c
 * at multiples of their own strides, so the same code walks the oracle's [q][i][j][k] arrays (k fastest) and a download
 * of the device's [q][z][y][x] buffers (x fastest) without a transpose.
 * Compiled with -ffp-contract=off: a*b+c is two roundings unless written fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define Q 19
static const int CX[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int CY[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int CZ[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
static const int OPP[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
/* opposite-direction pairs (p, m), e_m = -e_p */
static const int PAIR_P[9] = {1, 3, 5, 7, 9, 11, 13, 15, 17};
static const int PAIR_M[9] = {2, 4, 6, 10, 8, 14, 12, 18, 16};

typedef struct {
    int nx, ny, nz;
    long long sx, sy, sz;            /* scalar-volume strides of x, y, z */
    long long sq;                    /* stride between populations (= cells of the volume) */
    long long v_cell, v_comp;        /* vector fields: element (cell c, component k) at c * v_cell + k * v_comp */
    int per_x, per_y, per_z;
    int use_force, use_phase, les, porous;
    float tau_water, tau_air, gravity_lu, cs_smag, tau_min, tau_max, porous_darcy, porous_forch;
} phys_params;

typedef struct {
    const float *g;                  /* post-collision populations in */
    float *g_next;                   /* out (solid cells: copy of g) */
    float *rho, *u;                  /* out: moments of the streamed state (0 on solid cells) */
    const uint8_t *solid;            /* may be NULL */
    const float *body_force;         /* may be NULL */
    const float *phase;              /* may be NULL */
    const int32_t *filter_zone;      /* may be NULL unless porous */
    const int32_t *les_mask;         /* may be NULL (= 1 everywhere) */
} phys_fields;

static inline float vedot(int ex, int ey, int ez, float vx, float vy, float vz) {
    /* e . v for the + member of a pair: x, y, z order, one rounding per add / sub (lbm_phys.cuh:vedot) */
    float acc = 0.0f; int have = 0;
    const int e[3] = {ex, ey, ez}; const float v[3] = {vx, vy, vz};
    for (int a = 0; a < 3; ++a) {
        if (e[a] == 0) continue;
        if (!have) { acc = v[a]; have = 1; }
        else acc = e[a] > 0 ? acc + v[a] : acc - v[a];
    }
    return acc;
}

void phys_step(const phys_params *p, const phys_fields *F) {
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const float W0 = (float)(1.0 / 3.0), W1 = (float)(1.0 / 18.0), W2 = (float)(1.0 / 36.0);
    const float Wq[Q] = {W0, W1, W1, W1, W1, W1, W1, W2, W2, W2, W2, W2, W2, W2, W2, W2, W2, W2, W2};
    const float W1X2 = 2.0f * W1, W2X2 = 2.0f * W2, W1X6 = 6.0f * W1, W2X6 = 6.0f * W2, W1X18 = 18.0f * W1, W2X18 = 18.0f * W2;
    const float one = 1.0f, half = 0.5f;
    const int has_phase = F->phase != NULL && p->use_phase;
    const int has_force = (p->use_force && F->body_force != NULL) || (has_phase && p->gravity_lu != 0.0f);
    const double cs = (double)p->cs_smag;
    const float kk = (float)(18.0 * sqrt(2.0) * cs * cs);
    const long long ncell = (long long)nx * ny * nz;
#pragma omp parallel for schedule(static)
    for (long long lin = 0; lin < ncell; ++lin) {
        /* walk in the order of the unit stride: x fastest for a device download, z fastest for the oracle's own arrays */
        int x, y, z;
        if (p->sx == 1) { x = (int)(lin % nx); y = (int)((lin / nx) % ny); z = (int)(lin / ((long long)nx * ny)); }
        else { z = (int)(lin % nz); y = (int)((lin / nz) % ny); x = (int)(lin / ((long long)nz * ny)); }
        const long long c = x * p->sx + y * p->sy + z * p->sz;
        if (F->solid && F->solid[c]) {
            for (int q = 0; q < Q; ++q) F->g_next[q * p->sq + c] = F->g[q * p->sq + c];
            F->rho[c] = 0.0f;
            for (int k = 0; k < 3; ++k) F->u[c * p->v_cell + k * p->v_comp] = 0.0f;
            continue;
        }
        float f[Q];
        for (int q = 0; q < Q; ++q) {            /* d3q19_ref.py:_pull */
            int xs = x - CX[q], ys = y - CY[q], zs = z - CZ[q], oob = 0;
            if (xs < 0) { xs = nx - 1; oob |= !p->per_x; } else if (xs >= nx) { xs = 0; oob |= !p->per_x; }
            if (ys < 0) { ys = ny - 1; oob |= !p->per_y; } else if (ys >= ny) { ys = 0; oob |= !p->per_y; }
            if (zs < 0) { zs = nz - 1; oob |= !p->per_z; } else if (zs >= nz) { zs = 0; oob |= !p->per_z; }
            const long long cs_ = xs * p->sx + ys * p->sy + zs * p->sz;
            if (oob) f[q] = Wq[q];
            else if (F->solid && F->solid[cs_]) f[q] = F->g[OPP[q] * p->sq + c];
            else f[q] = F->g[q * p->sq + cs_];
        }
        float s[9], d[9];
        for (int k = 0; k < 9; ++k) { s[k] = f[PAIR_P[k]] + f[PAIR_M[k]]; d[k] = f[PAIR_P[k]] - f[PAIR_M[k]]; }
        float rho = f[0];
        for (int k = 0; k < 9; ++k) rho = rho + s[k];
        const float mx = (((d[0] + d[3]) + d[4]) + d[5]) + d[6];
        const float my = (((d[1] + d[3]) - d[4]) + d[7]) + d[8];
        const float mz = (((d[2] + d[5]) - d[6]) + d[7]) - d[8];
        const float inv_rho = one / rho;
        int forced = 0;
        float Fx = 0.0f, Fy = 0.0f, Fz = 0.0f, ux, uy, uz;
        const float ph = has_phase ? F->phase[c] : 0.0f;
        if (has_force) {
            forced = 1;
            if (p->use_force && F->body_force) {
                Fx = F->body_force[c * p->v_cell]; Fy = F->body_force[c * p->v_cell + p->v_comp]; Fz = F->body_force[c * p->v_cell + 2 * p->v_comp];
            }
            if (has_phase && p->gravity_lu != 0.0f) Fz = fmaf(-p->gravity_lu, ph, Fz);
            ux = fmaf(half, Fx, mx) * inv_rho; uy = fmaf(half, Fy, my) * inv_rho; uz = fmaf(half, Fz, mz) * inv_rho;
        } else { ux = mx * inv_rho; uy = my * inv_rho; uz = mz * inv_rho; }
        if (p->porous) {
            if (F->filter_zone[c] != 0) {
                const float darcy = p->porous_darcy, forch = p->porous_forch;
                const float vmag = sqrtf(fmaf(uz, uz, fmaf(uy, uy, ux * ux)));
                const float c0 = half * fmaf(half, darcy, one);
                const float c1 = half * forch;
                const float den = c0 + sqrtf(fmaf(c1, vmag, c0 * c0));
                const float sc = one / den;
                const float vx = ux * sc, vy = uy * sc, vz = uz * sc;
                const float umag = vmag * sc;
                const float cdrag = fmaf(forch, umag, darcy);
                const float cr = -(cdrag * rho);
                Fx = fmaf(cr, vx, Fx); Fy = fmaf(cr, vy, Fy); Fz = fmaf(cr, vz, Fz);
                ux = vx; uy = vy; uz = vz;
            }
            forced = 1;
        }
        const float tau0 = has_phase ? (ph > half ? p->tau_water : p->tau_air) : p->tau_water;
        float tau = tau0;
        if (p->les) {
            const float Mxx = (((s[0] + s[3]) + s[4]) + s[5]) + s[6];
            const float Myy = (((s[1] + s[3]) + s[4]) + s[7]) + s[8];
            const float Mzz = (((s[2] + s[5]) + s[6]) + s[7]) + s[8];
            const float Mxy = s[3] - s[4], Mxz = s[5] - s[6], Myz = s[7] - s[8];
            const float nr = -1.0f * rho;
            const float nrux = nr * ux, nruy = nr * uy;
            const float pxx = fmaf(nr, fmaf(ux, ux, W0), Mxx);
            const float pyy = fmaf(nr, fmaf(uy, uy, W0), Myy);
            const float pzz = fmaf(nr, fmaf(uz, uz, W0), Mzz);
            const float pxy = fmaf(nrux, uy, Mxy), pxz = fmaf(nrux, uz, Mxz), pyz = fmaf(nruy, uz, Myz);
            const float qa = fmaf(pzz, pzz, fmaf(pyy, pyy, pxx * pxx));
            const float qb = fmaf(pyz, pyz, fmaf(pxz, pxz, pxy * pxy));
            const float qsum = fmaf(2.0f, qb, qa);
            const float qn = sqrtf(qsum);
            const float arg = fmaf(kk * qn, inv_rho, tau0 * tau0);
            float tles = half * (tau0 + sqrtf(arg));
            if (F->les_mask && F->les_mask[c] == 0) tles = tau0;
            /* np.maximum(tau_min, np.minimum(tau_max, tles)) */
            const float t = (tles < p->tau_max || tles != tles) ? tles : p->tau_max;
            tau = (t > p->tau_min || t != t) ? t : p->tau_min;
        }
        const float omega = one / tau;
        const float nom = -1.0f * omega;
        const float u_sq = fmaf(uz, uz, fmaf(uy, uy, ux * ux));
        const float base = fmaf(-1.5f, u_sq, one);
        const float nws1 = (-W1X2) * rho, nws2 = (-W2X2) * rho, nwd1 = (-W1X6) * rho, nwd2 = (-W2X6) * rho;
        float out[Q];
        float f0 = fmaf(nom, fmaf((-W0) * rho, base, f[0]), f[0]);
        float pref = 0.0f, uF3 = 0.0f, c18[2] = {0, 0}, c6[2] = {0, 0}, nc2[2] = {0, 0};
        if (forced) {
            pref = fmaf(-0.5f, omega, one);
            uF3 = 3.0f * fmaf(uz, Fz, fmaf(uy, Fy, ux * Fx));
            f0 = fmaf((-W0) * pref, uF3, f0);
            c18[0] = W1X18 * pref; c18[1] = W2X18 * pref;
            c6[0] = W1X6 * pref; c6[1] = W2X6 * pref;
            nc2[0] = (-W1X2) * pref; nc2[1] = (-W2X2) * pref;
        }
        out[0] = f0;
        for (int k = 0; k < 9; ++k) {
            const int pp = PAIR_P[k], pm = PAIR_M[k], a = k < 3 ? 0 : 1;
            const float eu = vedot(CX[pp], CY[pp], CZ[pp], ux, uy, uz);
            const float A = fmaf(4.5f * eu, eu, base);
            const float ns = fmaf(a == 0 ? nws1 : nws2, A, s[k]);
            const float nd = fmaf(a == 0 ? nwd1 : nwd2, eu, d[k]);
            float sp = fmaf(nom, ns, s[k]);
            float dp = fmaf(nom, nd, d[k]);
            if (forced) {
                const float eF = vedot(CX[pp], CY[pp], CZ[pp], Fx, Fy, Fz);
                sp = fmaf(eu * eF, c18[a], sp);
                sp = fmaf(nc2[a], uF3, sp);
                dp = fmaf(eF, c6[a], dp);
            }
            const float hs = half * sp;
            out[pp] = fmaf(half, dp, hs);
            out[pm] = fmaf(-0.5f, dp, hs);
        }
        for (int q = 0; q < Q; ++q) F->g_next[q * p->sq + c] = out[q];
        F->rho[c] = rho;
        F->u[c * p->v_cell] = ux; F->u[c * p->v_cell + p->v_comp] = uy; F->u[c * p->v_cell + 2 * p->v_comp] = uz;
    }
}

/* number of 32-bit words that differ between a and b over `nfields` volumes of `vol` cells, fluid cells only
 * (solid == NULL: every cell); NaNs compare by bit pattern */
long long phys_count_mismatch(const uint32_t *a, const uint32_t *b, const uint8_t *solid, long long vol, int nfields) {
    long long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long long c = 0; c < vol; ++c) {
        if (solid && solid[c]) continue;
        for (int k = 0; k < nfields; ++k) bad += a[k * vol + c] != b[k * vol + c];
    }
    return bad;
}
