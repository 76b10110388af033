// TEST INFRASTRUCTURE -- not product code.  The legacy-compatible step of the product, step_cells<LBM_COMPAT_REFERENCE,
// MODE_BULK, ..., VEC = 1> from pour_over_coffee_lbm_b200/csrc/lbm_step_kernel.cuh (pull, halfway bounce-back / open-face
// inflow from the neighbour masks, FD-LES on the lagged u, moments, clamped Guo term, BGK, filter damping, write-back),
// compiled by the HOST compiler and executed cell by cell.  The one-cell-per-thread path has no warp intrinsics; the
// 4-cells-per-thread path (shuffles, predicated PTX loads) and the packed compat = physical kernels are GPU-only.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

struct EmuIdx { unsigned x, y, z; };
static EmuIdx emu_block_idx, emu_thread_idx, emu_block_dim;
#define blockIdx emu_block_idx
#define threadIdx emu_thread_idx
#define blockDim emu_block_dim
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
static inline float __shfl_up_sync(unsigned, float v, int) { return v; }       // only named by the VEC > 1 branches (discarded)
static inline float __shfl_down_sync(unsigned, float v, int) { return v; }
#define __launch_bounds__(...)
#define LBM_EMULATE_ON_HOST 1
#define LBM_PHYS_COLLISION_ONLY 1
#include "../../pour_over_coffee_lbm_b200/csrc/lbm_step_kernel.cuh"

using namespace lbm;

// EMU_VEC (cells per thread of the legacy-compatible step; emu_step_reference_vec2.cpp sets 2: the packed collision
// collide_reference_t<P2> on host stand-ins of the f32x2 primitives, x -+ 1 words fetched by every thread itself)
#ifndef EMU_VEC
#define EMU_VEC 1
#endif
template <bool LES, bool POROUS>
static void run(const StepArgs &P) {
    const Grid &G = P.g;
    for (int z = 0; z < G.nz; ++z)
        for (int y = 0; y < G.ny; ++y)
            for (int x = 0; x < G.nx; x += EMU_VEC)
                step_cells<LBM_COMPAT_REFERENCE, MODE_BULK, true, LES, POROUS, EMU_VEC, true>(P, x, y, z, true, (unsigned)((x / EMU_VEC) & 31));
}

extern "C" int emu_step_reference_slab(int nx, int ny, int nz, int z0, int nz_global, const float *src, float *dst, float *rho, const float *u_src,
                                       float *u_dst, const float *force, const float *phase, const float *blockage, const uint8_t *flags,
                                       const unsigned long long *nbr, int les, int porous, float tau_water, float tau_air, float gravity_lu,
                                       float cs_smag, float tau_min, float tau_max, float K_lu, float beta_lu, float c_darcy, float c_forch) {
    StepArgs P{};
    P.g.nx = nx; P.g.ny = ny; P.g.nz = nz; P.g.zg = 1; P.g.nz_global = nz_global; P.g.z0 = z0;
    P.g.per_x = P.g.per_y = P.g.per_z = 0;
    P.g.plane = (long long)nx * ny; P.g.vol = P.g.plane * (nz + 2);
    P.src = src; P.dst = dst; P.rho = rho; P.u_src = u_src; P.u_dst = u_dst; P.force = force; P.phase = phase; P.blockage = blockage;
    P.flags = flags; P.nbr = nbr; P.write_macro = 1;
    P.tau_water = tau_water; P.tau_air = tau_air; P.gravity_lu = gravity_lu; P.tau_min = tau_min; P.tau_max = tau_max;
    P.les_k = (cs_smag * 1.0f) * (cs_smag * 1.0f);
    P.K_lu = K_lu; P.beta_lu = beta_lu; P.c_darcy = c_darcy; P.c_forch = c_forch;
    if (les && porous) run<true, true>(P);
    else if (les) run<true, false>(P);
    else if (porous) run<false, true>(P);
    else run<false, false>(P);
    return 0;
}

extern "C" int emu_step_reference(int nx, int ny, int nz, const float *src, float *dst, float *rho, const float *u_src, float *u_dst,
                                  const float *force, const float *phase, const float *blockage, const uint8_t *flags,
                                  const unsigned long long *nbr, int les, int porous, float tau_water, float tau_air, float gravity_lu,
                                  float cs_smag, float tau_min, float tau_max, float K_lu, float beta_lu, float c_darcy, float c_forch) {
    StepArgs P{};
    P.g.nx = nx; P.g.ny = ny; P.g.nz = nz; P.g.zg = 0; P.g.nz_global = nz; P.g.z0 = 0;
    P.g.per_x = P.g.per_y = P.g.per_z = 0;
    P.g.plane = (long long)nx * ny; P.g.vol = P.g.plane * nz;
    P.src = src; P.dst = dst; P.rho = rho; P.u_src = u_src; P.u_dst = u_dst; P.force = force; P.phase = phase; P.blockage = blockage;
    P.flags = flags; P.nbr = nbr; P.write_macro = 1;
    P.tau_water = tau_water; P.tau_air = tau_air; P.gravity_lu = gravity_lu; P.tau_min = tau_min; P.tau_max = tau_max;
    P.les_k = (cs_smag * 1.0f) * (cs_smag * 1.0f);                              // as lbm_api.cu forms it for compat = reference
    P.K_lu = K_lu; P.beta_lu = beta_lu; P.c_darcy = c_darcy; P.c_forch = c_forch;
    if (les && porous) run<true, true>(P);
    else if (les) run<true, false>(P);
    else if (porous) run<false, true>(P);
    else run<false, false>(P);
    return 0;
}

// compat = physical, fully periodic box (the headline configuration): step_cells<LBM_COMPAT_PHYSICAL, MODE_DENSE, ..., VEC = 1>
// -- pull with in-kernel wrap, collide_phys<float>, write-back.  The GPU runs the VEC = 4 instantiation of the same function
// (128-bit loads, shuffles, packed f32x2 collision); the arithmetic contract makes both produce the same bits.
template <bool LES>
static void run_dense(const StepArgs &P) {
    const Grid &G = P.g;
    for (int z = 0; z < G.nz; ++z)
        for (int y = 0; y < G.ny; ++y)
            for (int x = 0; x < G.nx; ++x)
                step_cells<LBM_COMPAT_PHYSICAL, MODE_DENSE, false, LES, false, 1, true>(P, x, y, z, true, (unsigned)(x & 31));
}
extern "C" int emu_step_physical_dense(int nx, int ny, int nz, int steps, float *g0, float *g1, float *rho, float *u, int les, float tau,
                                       float cs_smag, float tau_min, float tau_max, int write_macro_last_only) {
    StepArgs P{};
    P.g.nx = nx; P.g.ny = ny; P.g.nz = nz; P.g.zg = 0; P.g.nz_global = nz; P.g.z0 = 0;
    P.g.per_x = P.g.per_y = P.g.per_z = 1;
    P.g.plane = (long long)nx * ny; P.g.vol = P.g.plane * nz;
    P.rho = rho; P.u_dst = u; P.u_src = u;
    P.tau_water = tau; P.tau_air = tau; P.tau_min = tau_min; P.tau_max = tau_max;
    P.les_k = (float)(18.0 * sqrt(2.0) * (double)cs_smag * (double)cs_smag);
    float *buf[2] = {g0, g1};
    for (int s = 0; s < steps; ++s) {
        P.src = buf[s & 1]; P.dst = buf[(s + 1) & 1];
        P.write_macro = (!write_macro_last_only || s == steps - 1) ? 1 : 0;
        if (les) run_dense<true>(P); else run_dense<false>(P);
    }
    return steps & 1;      // index of the buffer that holds the newest populations
}
