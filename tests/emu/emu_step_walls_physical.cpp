// TEST INFRASTRUCTURE -- not product code.  The walls path of compat = physical, phys_walls_kernel<..., VEC = 1, ...> +
// phys_finish of pour_over_coffee_lbm_b200/csrc/lbm_phys.cuh (pure pull, open-face inflow, collide_phys<float>, write-back,
// WRITE-SIDE halfway bounce-back into the solid neighbours' slots), compiled by the HOST compiler and executed warp-tile by
// warp-tile, lane by lane.  The GPU default is the two-cell instantiation (packed f32x2, inline PTX) of the same template;
// the arithmetic contract makes both produce the same bits.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <vector>

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
static inline unsigned long long __cvta_generic_to_shared(const void *p) { return (unsigned long long)p; }   // named by an unused cp.async helper
#define __launch_bounds__(...)
#define LBM_EMULATE_ON_HOST 1
#include "../../pour_over_coffee_lbm_b200/csrc/lbm_common.cuh"
#include "../../pour_over_coffee_lbm_b200/csrc/lbm_phys.cuh"

using namespace lbm;

template <bool FORCED, bool LES, bool POROUS>
static void run(const StepArgs &P) {
    emu_block_dim = {32, 1, 1};
    for (int w = 0; w < P.n_items; ++w)
        for (unsigned lane = 0; lane < 32; ++lane) {
            emu_block_idx = {(unsigned)w, 0, 0}; emu_thread_idx = {lane, 0, 0};
            phys_walls_kernel<FORCED, LES, POROUS, 1, 32, true, 1>(P);
        }
}

extern "C" int emu_step_walls_physical(int nx, int ny, int nz, int periodic, int steps, float *g0, float *g1, float *rho, float *u,
                                       const float *force, const float *phase, const uint8_t *flags, const unsigned long long *nbr, int les,
                                       int porous, float tau_water, float tau_air, float gravity_lu, float cs_smag, float tau_min, float tau_max,
                                       float porous_darcy, float porous_forch) {
    StepArgs P{};
    P.g.nx = nx; P.g.ny = ny; P.g.nz = nz; P.g.zg = 0; P.g.nz_global = nz; P.g.z0 = 0;
    P.g.per_x = periodic & 1; P.g.per_y = (periodic >> 1) & 1; P.g.per_z = (periodic >> 2) & 1;
    P.g.plane = (long long)nx * ny; P.g.vol = P.g.plane * nz;
    P.rho = rho; P.u_dst = u; P.u_src = u; P.force = force; P.phase = phase; P.flags = flags; P.nbr = nbr; P.write_macro = 1;
    P.tau_water = tau_water; P.tau_air = tau_air; P.gravity_lu = gravity_lu; P.tau_min = tau_min; P.tau_max = tau_max;
    P.les_k = (float)(18.0 * sqrt(2.0) * (double)cs_smag * (double)cs_smag);
    P.porous_darcy = porous_darcy; P.porous_forch = porous_forch;
    // the active warp-tile list as lbm_pack_flags builds it: x_segment | y << 8 | z << 20 for every 32-cell segment with a fluid cell
    std::vector<unsigned> items;
    const int segs = (nx + 31) / 32;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int s = 0; s < segs; ++s) {
                bool any = false;
                for (int x = 32 * s; x < nx && x < 32 * (s + 1); ++x) any |= !(flags[((long long)z * ny + y) * nx + x] & LBM_FLAG_SOLID);
                if (any) items.push_back((unsigned)s | ((unsigned)y << 8) | ((unsigned)z << 20));
            }
    P.items = items.data(); P.item_begin = 0; P.n_items = (int)items.size();
    float *buf[2] = {g0, g1};
    const bool forced = force != nullptr || phase != nullptr;
    for (int s = 0; s < steps; ++s) {
        P.src = buf[s & 1]; P.dst = buf[(s + 1) & 1];
        if (forced && les && porous) run<true, true, true>(P);
        else if (forced && les) run<true, true, false>(P);
        else if (forced) run<true, false, false>(P);
        else if (les) run<false, true, false>(P);
        else run<false, false, false>(P);
    }
    return steps & 1;
}
