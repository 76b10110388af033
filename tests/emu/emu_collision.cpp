// TEST INFRASTRUCTURE -- not product code.  The collision operator of compat = physical, collide_phys<float, ...>
// (pour_over_coffee_lbm_b200/csrc/lbm_phys.cuh: the arithmetic contract of the headline kernel), compiled by the HOST
// compiler and applied cell by cell after a periodic pull, so that the operator's source is checked against
// oracle/d3q19_ref.py:step_physical on every CPU test run.  The explicitly rounded intrinsics map to their IEEE host
// equivalents (one rounding each: +, -, *, fmaf, 1/x, sqrtf); the packed f32x2 instantiation (inline PTX) is not
// compiled here -- on the GPU lbm_selftest_math() proves it equal to this scalar one.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline unsigned __float_as_uint(float a) { unsigned u; __builtin_memcpy(&u, &a, 4); return u; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
#define __launch_bounds__(...)
#define LBM_EMULATE_ON_HOST 1
#include "../../pour_over_coffee_lbm_b200/csrc/lbm_common.cuh"
#define LBM_PHYS_COLLISION_ONLY 1
#include "../../pour_over_coffee_lbm_b200/csrc/lbm_phys.cuh"

using namespace lbm;

template <bool FORCED, bool LES, bool POROUS>
static void run(int nx, int ny, int nz, const float *g, float *g_out, float *rho, float *u, const float *force, const float *phase,
                const uint8_t *flags, const StepArgs &P) {
    const long long plane = (long long)nx * ny, vol = plane * nz;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const long long c = ((long long)z * ny + y) * nx + x;
                float f[Q];
                for (int q = 0; q < Q; ++q) {
                    const int sx = (x - cx(q) + nx) % nx, sy = (y - cy(q) + ny) % ny, sz = (z - cz(q) + nz) % nz;
                    f[q] = g[q * vol + ((long long)sz * ny + sy) * nx + sx];
                }
                CellIn<float> in{};
                in.Fx = force ? force[c] : 0.0f; in.Fy = force ? force[vol + c] : 0.0f; in.Fz = force ? force[2 * vol + c] : 0.0f;
                in.phase = phase ? phase[c] : 0.0f;
                in.flag[0] = flags ? flags[c] : (unsigned)LBM_FLAG_LES;
                CellMacro<float> m;
                const bool has_phase = FORCED && phase != nullptr;
                const bool has_force = FORCED && (force != nullptr || (has_phase && P.gravity_lu != 0.0f));
                collide_phys<float, FORCED, LES, POROUS, true, false, true>(f, in, m, P, has_phase, has_force);
                for (int q = 0; q < Q; ++q) g_out[q * vol + c] = f[q];
                rho[c] = m.rho; u[c] = m.ux; u[vol + c] = m.uy; u[2 * vol + c] = m.uz;
            }
}

extern "C" int emu_collide_periodic(int nx, int ny, int nz, const float *g, float *g_out, float *rho, float *u, const float *force,
                                    const float *phase, const uint8_t *flags, int les, int porous, float tau_water, float tau_air, float gravity_lu,
                                    float cs_smag, float tau_min, float tau_max, float porous_darcy, float porous_forch, float mrt_magic) {
    StepArgs P{};
    P.tau_water = tau_water; P.tau_air = tau_air; P.gravity_lu = gravity_lu; P.tau_min = tau_min; P.tau_max = tau_max; P.mrt_magic = mrt_magic;
    P.les_k = (float)(18.0 * sqrt(2.0) * (double)cs_smag * (double)cs_smag);          // as lbm_api.cu forms it
    P.porous_darcy = porous_darcy; P.porous_forch = porous_forch;
    const bool forced = force != nullptr || phase != nullptr;
    if (forced && les && porous) run<true, true, true>(nx, ny, nz, g, g_out, rho, u, force, phase, flags, P);
    else if (forced && les) run<true, true, false>(nx, ny, nz, g, g_out, rho, u, force, phase, flags, P);
    else if (forced) run<true, false, false>(nx, ny, nz, g, g_out, rho, u, force, phase, flags, P);
    else if (les) run<false, true, false>(nx, ny, nz, g, g_out, rho, u, force, phase, flags, P);
    else run<false, false, false>(nx, ny, nz, g, g_out, rho, u, force, phase, flags, P);
    return 0;
}
