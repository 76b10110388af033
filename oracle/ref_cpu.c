/*
 * CPU ORACLE (test infrastructure, NOT product code).
 *
 * C99 + OpenMP restatement of the reference's Taichi kernels for one call of
 * LBMSolver.step() (src/core/legacy/lbm_solver.py:817-867), in the reference's
 * own structure: LES pre-pass, macroscopic kernel, collide + push-stream kernel,
 * copy-swap kernel, filter damping, face density writes.  It is (a) the fast
 * checker for the CUDA path at sizes NumPy is too slow for and (b) the timed
 * "C restatement of the Taichi ti.cpu kernels" baseline of bench.py
 * (cpu_baseline.kind = "port"; Taichi is not installable here).
 *
 * PARITY STATUS: pinned against the reference's own source code -- recorded runs of
 * the unmodified reference modules under a pure-Python Taichi stand-in
 * (tests/golden/reference_run_step_*.npz, see the header of oracle/d3q19_ref.py); this
 * file reproduces their rho, u and f bit for bit (tests/test_oracle_vs_reference_run.py)
 * and agrees bit for bit with the NumPy restatement (tests/test_oracle_c_vs_numpy.py).
 *
 * Layout: the reference's Taichi dense layout, f[q][i][j][k] with k (z) fastest,
 * u[i][j][k][3] AoS.  All arithmetic f32, left-to-right as written in the
 * reference, compiled with -ffp-contract=off (no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Q 19
/* config/core.py:36-38 */
static const int CXc[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int CYc[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int CZc[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
/* src/core/lbm_algorithms.py:158-164 (equilibrium-only table, quirk Q1) */
static const int EXc[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int EYc[Q] = {0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int EZc[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1};
/* legacy/lbm_solver.py:431-439 evaluated on the config table */
static const int OPPc[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};

typedef struct {
    int nx, ny, nz;
    int use_les;
    int apply_filter;   /* boundary manager has a filter system */
    int apply_faces;    /* top/bottom/outlet strategies */
    float tau_water, tau_air, gravity_lu, les_cs;
    float K_lu, beta_lu, c_darcy, c_forch;
} ref_params;

typedef struct {
    float *f, *f_new;            /* [19][nx][ny][nz] */
    float *rho, *u, *u_sq;       /* [nx][ny][nz], [nx][ny][nz][3], [nx][ny][nz] */
    float *phase, *body_force;   /* [nx][ny][nz], [..][3] */
    float *nu_sgs;
    const uint8_t *solid;
    const int32_t *les_mask;
    const int32_t *filter_zone;  /* may be NULL */
    const float *filter_blockage;/* may be NULL */
} ref_fields;

static inline float wq(int q) { return q == 0 ? (float)(1.0 / 3.0) : (q < 7 ? (float)(1.0 / 18.0) : (float)(1.0 / 36.0)); }

static inline float edot(int ex, int ey, int ez, float vx, float vy, float vz) {
    /* exact products with {0,+-1}; zeros add exactly -> sum of the non-zero terms in x,y,z order */
    float acc = 0.0f; int have = 0;
    if (ex) { acc = ex > 0 ? vx : -vx; have = 1; }
    if (ey) { float t = ey > 0 ? vy : -vy; acc = have ? acc + t : t; have = 1; }
    if (ez) { float t = ez > 0 ? vz : -vz; acc = have ? acc + t : t; have = 1; }
    return acc;
}

static inline float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return (ax * bx + ay * by) + az * bz;
}

#define IDX(i, j, k) (((size_t)(i) * ny + (j)) * nz + (k))

/* les_turbulence.py:318-380 */
void ref_les_update(const ref_params *p, ref_fields *F) {
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const float cs = p->les_cs;
    const float csd = (cs * 1.0f) * (cs * 1.0f);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                size_t c = IDX(i, j, k);
                if (i == 0 || j == 0 || k == 0 || i == nx - 1 || j == ny - 1 || k == nz - 1) { F->nu_sgs[c] = 0.0f; continue; }
                if (F->les_mask[c] == 0) { F->nu_sgs[c] = 0.0f; continue; }
                const float *uip = F->u + 3 * IDX(i + 1, j, k), *uim = F->u + 3 * IDX(i - 1, j, k);
                const float *ujp = F->u + 3 * IDX(i, j + 1, k), *ujm = F->u + 3 * IDX(i, j - 1, k);
                const float *ukp = F->u + 3 * IDX(i, j, k + 1), *ukm = F->u + 3 * IDX(i, j, k - 1);
                float dudx = (uip[0] - uim[0]) * 0.5f, dudy = (ujp[0] - ujm[0]) * 0.5f, dudz = (ukp[0] - ukm[0]) * 0.5f;
                float dvdx = (uip[1] - uim[1]) * 0.5f, dvdy = (ujp[1] - ujm[1]) * 0.5f, dvdz = (ukp[1] - ukm[1]) * 0.5f;
                float dwdx = (uip[2] - uim[2]) * 0.5f, dwdy = (ujp[2] - ujm[2]) * 0.5f, dwdz = (ukp[2] - ukm[2]) * 0.5f;
                float S11 = dudx, S22 = dvdy, S33 = dwdz;
                float S12 = 0.5f * (dudy + dvdx), S13 = 0.5f * (dudz + dwdx), S23 = 0.5f * (dvdz + dwdy);
                float mag = sqrtf(2.0f * (((S11 * S11 + S22 * S22) + S33 * S33) + 2.0f * ((S12 * S12 + S13 * S13) + S23 * S23)));
                if (mag < 1e-3f) { F->nu_sgs[c] = 0.0f; continue; }
                if (fabsf(F->phase[c]) < 0.9f) { F->nu_sgs[c] = 0.0f; continue; }
                float nu = csd * mag;
                F->nu_sgs[c] = nu < 0.1f ? nu : 0.1f;
            }
}

/* legacy/lbm_solver.py:488-535 */
void ref_macroscopic(const ref_params *p, ref_fields *F) {
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const size_t n = (size_t)nx * ny * nz;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < n; ++c) {
        if (F->solid[c] != 0) continue;
        float rho = 0.0f;
        for (int q = 0; q < Q; ++q) rho += F->f[q * n + c];
        float mx = 0.0f, my = 0.0f, mz = 0.0f;
        for (int q = 0; q < Q; ++q) {
            float fq = F->f[q * n + c];
            if (CXc[q]) mx += fq * (float)CXc[q];
            if (CYc[q]) my += fq * (float)CYc[q];
            if (CZc[q]) mz += fq * (float)CZc[q];
        }
        float ph = F->phase[c];
        float gz = ph > 0.001f ? -(p->gravity_lu * ph) : 0.0f;
        float Fx = 0.0f + F->body_force[3 * c], Fy = 0.0f + F->body_force[3 * c + 1], Fz = gz + F->body_force[3 * c + 2];
        float ux = 0.0f, uy = 0.0f, uz = 0.0f;
        if (rho > 1e-12f) {
            ux = (mx + 0.5f * Fx) / rho; uy = (my + 0.5f * Fy) / rho; uz = (mz + 0.5f * Fz) / rho;
        }
        F->rho[c] = rho;
        F->u[3 * c] = ux; F->u[3 * c + 1] = uy; F->u[3 * c + 2] = uz;
        F->u_sq[c] = dot3(ux, uy, uz, ux, uy, uz);
    }
}

/* legacy/lbm_solver.py:537-628, 688-764; lbm_algorithms.py:183-218 */
void ref_collide_stream(const ref_params *p, ref_fields *F) {
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const size_t n = (size_t)nx * ny * nz;
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                size_t c = IDX(i, j, k);
                if (F->solid[c] != 0) continue;
                float rho = F->rho[c];
                float ux = F->u[3 * c], uy = F->u[3 * c + 1], uz = F->u[3 * c + 2];
                float ph = F->phase[c];
                float gz = ph > 0.001f ? -(p->gravity_lu * ph) : 0.0f;
                float Fx = 0.0f + F->body_force[3 * c], Fy = 0.0f + F->body_force[3 * c + 1], Fz = gz + F->body_force[3 * c + 2];
                float tau = ph > 0.5f ? p->tau_water : p->tau_air;
                if (p->use_les) tau = tau + 3.0f * F->nu_sgs[c];
                tau = fmaxf(0.55f, fminf(1.90f, tau));
                float omega = 1.0f / tau;
                /* forcing prerequisites (per cell, identical for all q) */
                float fnorm = sqrtf(dot3(Fx, Fy, Fz, Fx, Fy, Fz));
                int forced = fnorm > 1e-15f;
                float tau_safe = fminf(fmaxf(tau, 0.6f), 1.5f);
                float sf = fnorm > 10.0f ? 10.0f / fnorm : 1.0f;
                float fsx = Fx * sf, fsy = Fy * sf, fsz = Fz * sf;
                float unorm = sqrtf(dot3(ux, uy, uz, ux, uy, uz));
                float usx = ux, usy = uy, usz = uz;
                if (unorm > 0.2f) { float s = 0.2f / unorm; usx = ux * s; usy = uy * s; usz = uz * s; }
                float uf = dot3(usx, usy, usz, fsx, fsy, fsz);
                float u_sq = dot3(ux, uy, uz, ux, uy, uz);
                for (int q = 0; q < Q; ++q) {
                    float w = wq(q);
                    float eu = edot(EXc[q], EYc[q], EZc[q], ux, uy, uz);
                    float feq = (w * rho) * (((1.0f + 3.0f * eu) + (4.5f * eu) * eu) - 1.5f * u_sq);
                    float Fq = 0.0f;
                    if (forced) {
                        float eus = edot(CXc[q], CYc[q], CZc[q], usx, usy, usz);
                        float ef = edot(CXc[q], CYc[q], CZc[q], fsx, fsy, fsz);
                        float coeff = w * (1.0f - 0.5f / tau_safe);
                        Fq = coeff * (3.0f * ef + (9.0f * eus) * uf);
                        Fq = fmaxf(-0.5f, fminf(0.5f, Fq));
                    }
                    float fq = F->f[q * n + c];
                    float fpost = (fq - omega * (fq - feq)) + Fq;
                    int ni = i + CXc[q], nj = j + CYc[q], nk = k + CZc[q];
                    if (ni >= 0 && ni < nx && nj >= 0 && nj < ny && nk >= 0 && nk < nz) {
                        size_t t = IDX(ni, nj, nk);
                        if (F->solid[t] == 0) F->f_new[q * n + t] = fpost;
                        else F->f_new[OPPc[q] * n + c] = fpost;
                    }
                }
            }
}

/* legacy/lbm_solver.py:630-654: element-wise exchange over all Q*N^3 entries */
void ref_swap_copy(const ref_params *p, ref_fields *F) {
    const size_t n = (size_t)Q * p->nx * p->ny * p->nz;
    float *a = F->f, *b = F->f_new;
#pragma omp parallel for schedule(static)
    for (size_t t = 0; t < n; ++t) { float x = a[t]; a[t] = b[t]; b[t] = x; }
}

/* filter_paper.py:538-614 */
void ref_apply_filter_effects(const ref_params *p, ref_fields *F) {
    if (!F->filter_zone) return;
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const float K = p->K_lu, beta = p->beta_lu;
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 1; i < nx - 1; ++i)
        for (int j = 1; j < ny - 1; ++j)
            for (int k = 1; k < nz - 1; ++k) {
                size_t c = IDX(i, j, k);
                if (F->filter_zone[c] != 1 || F->solid[c] != 0) continue;
                float ux = F->u[3 * c], uy = F->u[3 * c + 1], uz = F->u[3 * c + 2];
                float umag = sqrtf(dot3(ux, uy, uz, ux, uy, uz));
                if (umag > 1e-8f && K > 1e-12f) {
                    float darcy = p->c_darcy / K;
                    float forch = ((p->c_forch * beta) * umag) / sqrtf(K);
                    float blk = F->filter_blockage ? F->filter_blockage[c] : 0.0f;
                    float total = (darcy + forch) * (1.0f + blk);
                    float r = expf((-total) * 0.5f);
                    r = fmaxf(0.1f, r);
                    float hf = (r + 1.0f) * 0.5f;
                    F->u[3 * c + 2] = uz * r; F->u[3 * c] = ux * hf; F->u[3 * c + 1] = uy * hf;
                }
            }
}

/* boundary_conditions.py:178-324 (SoA branch: only the rho writes are observable) */
void ref_face_bcs(const ref_params *p, ref_fields *F) {
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    float *rho = F->rho; const uint8_t *s = F->solid;
    for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) if (s[IDX(i, j, nz - 1)] == 0) rho[IDX(i, j, nz - 1)] = 1.0f;
    for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) if (s[IDX(i, j, 0)] == 0) rho[IDX(i, j, 0)] = rho[IDX(i, j, 1)];
    for (int j = 0; j < ny; ++j) for (int k = 0; k < nz; ++k) {
        if (s[IDX(0, j, k)] == 0) rho[IDX(0, j, k)] = rho[IDX(1, j, k)];
        if (s[IDX(nx - 1, j, k)] == 0) rho[IDX(nx - 1, j, k)] = rho[IDX(nx - 2, j, k)];
    }
    for (int i = 0; i < nx; ++i) for (int k = 0; k < nz; ++k) {
        if (s[IDX(i, 0, k)] == 0) rho[IDX(i, 0, k)] = rho[IDX(i, 1, k)];
        if (s[IDX(i, ny - 1, k)] == 0) rho[IDX(i, ny - 1, k)] = rho[IDX(i, ny - 2, k)];
    }
    for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) if (s[IDX(i, j, 0)] == 0) rho[IDX(i, j, 0)] = rho[IDX(i, j, 1)];
}

/* legacy/lbm_solver.py:817-867.  `nsteps` calls of step(); the copy-swap really copies. */
void ref_step(const ref_params *p, ref_fields *F, int nsteps) {
    for (int s = 0; s < nsteps; ++s) {
        if (p->use_les) ref_les_update(p, F);
        ref_macroscopic(p, F);
        ref_collide_stream(p, F);
        ref_swap_copy(p, F);
        if (p->apply_filter) ref_apply_filter_effects(p, F);
        if (p->apply_faces) ref_face_bcs(p, F);
    }
}

/* legacy/lbm_solver.py:1067-1112 */
void ref_init_fields(const ref_params *p, ref_fields *F) {
    const size_t n = (size_t)p->nx * p->ny * p->nz;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < n; ++c) {
        F->rho[c] = 1.0f; F->phase[c] = 0.0f; F->u_sq[c] = 0.0f; F->nu_sgs[c] = 0.0f;
        for (int d = 0; d < 3; ++d) { F->u[3 * c + d] = 0.0f; F->body_force[3 * c + d] = 0.0f; }
        for (int q = 0; q < Q; ++q) { F->f[q * n + c] = wq(q) * 1.0f; F->f_new[q * n + c] = F->f[q * n + c]; }
    }
}

/* filter_paper.py:206-286.  Constants arrive pre-rounded to f32 (they are Python-scope f64 folds). */
void ref_v60_solid(int nx, int ny, int nz, float top_r, float bot_r, float cup_h, float gap, uint8_t *solid) {
    const float cx = (float)(nx * 0.5), cy = (float)(ny * 0.5);
    const float bottom_z = 5.0f, wall = 2.0f;
    const float top_z = bottom_z + cup_h;
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) {
            float dx = (float)i - cx, dy = (float)j - cy;
            float r = sqrtf(dx * dx + dy * dy);
            for (int k = 0; k < nz; ++k) {
                float z = (float)k; int is = 0;
                if (z <= bottom_z) { if (r > bot_r) is = 1; }
                else if (z <= top_z) {
                    float hr = (z - bottom_z) / cup_h;
                    float inner = bot_r + (top_r - bot_r) * hr;
                    if (r > (inner + gap) + wall) is = 1;
                } else { if (r > top_r + wall) is = 1; }
                if (i <= 2 || i >= nx - 3 || j <= 2 || j >= ny - 3 || k <= 2 || k >= nz - 3) is = 1;
                solid[IDX(i, j, k)] = (uint8_t)is;
            }
        }
}

/* ------------------------------------------------------------------------
 * "optimised CPU" variant reported for fairness: same arithmetic,
 * macroscopic fused into the collide kernel and pointer swap instead of the copy.
 * ---------------------------------------------------------------------- */
void ref_step_fused(const ref_params *p, ref_fields *F, int nsteps) {
    for (int s = 0; s < nsteps; ++s) {
        if (p->use_les) ref_les_update(p, F);
        ref_macroscopic(p, F);
        ref_collide_stream(p, F);
        float *t = F->f; F->f = F->f_new; F->f_new = t;
        if (p->apply_filter) ref_apply_filter_effects(p, F);
        if (p->apply_faces) ref_face_bcs(p, F);
    }
}

int ref_num_threads(void) {
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
#endif
    return n;
}

/* Exact f32 fused multiply-add on arrays: out[i] = round(a[i]*b[i] + c[i]) with a single rounding (C99 fmaf).
 * oracle/d3q19_ref.py:step_physical calls it for every FMA of the compat=physical arithmetic contract (NumPy has no
 * fused operation; emulating it in f64 double-rounds). */
void ref_fmaf_array(const float *a, const float *b, const float *c, float *out, long n) {
#pragma omp parallel for schedule(static) if (n > 65536)
    for (long i = 0; i < n; ++i) out[i] = fmaf(a[i], b[i], c[i]);
}
