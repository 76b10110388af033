// TEST INFRASTRUCTURE -- not product code.  The particle kernels of pour_over_coffee_lbm_b200/csrc/lbm_particles.cu compiled by
// the HOST compiler and executed thread by thread (see emu_producers.cpp).  Warp intrinsics are modelled for a warp whose
// lanes run one after the other: __match_any_sync finds no peer, so every particle scatters its own eight corners -- the
// aggregation is an optimisation of the scatter's ORDER, which the reference leaves unspecified (atomics); everything else
// (gather, drag law, under-relaxation, integrator, force producer) is the product's statement sequence.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

struct EmuIdx { unsigned x, y, z; };
static EmuIdx emu_block_idx, emu_thread_idx, emu_block_dim;
#define blockIdx emu_block_idx
#define threadIdx emu_thread_idx
#define blockDim emu_block_dim
static inline float atomicAdd(float *p, float v) { const float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __match_any_sync(unsigned, long long) { return 1u << (emu_thread_idx.x & 31u); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline float __shfl_sync(unsigned, float v, int) { return v; }
#define __launch_bounds__(...)
#define LBM_EMULATE_ON_HOST 1
#include "../../pour_over_coffee_lbm_b200/csrc/lbm_particles.cu"

using namespace lbm;

template <class F>
static void run(unsigned n, unsigned block, F &&kernel) {
    emu_block_dim = {block, 1, 1};
    for (unsigned b = 0; b < (n + block - 1) / block; ++b)
        for (unsigned t = 0; t < block; ++t) { emu_block_idx = {b, 0, 0}; emu_thread_idx = {t, 0, 0}; kernel(); }
}
static Grid make_grid(int nx, int ny, int nz, int zg = 0, int z0 = 0, int nz_global = 0) {
    Grid G{};
    G.nx = nx; G.ny = ny; G.nz = nz; G.zg = zg; G.nz_global = nz_global > 0 ? nz_global : nz; G.z0 = z0;
    G.plane = (long long)nx * ny; G.vol = G.plane * (nz + 2 * zg);
    return G;
}

extern "C" {
int emu_particles_couple(int nx, int ny, int nz, const float *u, float *reaction, lbm_particles *ps, float rho_w, float mu_w, float relax) {
    ParticleArgs A{make_grid(nx, ny, nz), u, reaction, *ps, rho_w, mu_w, relax};
    for (long long i = 0; i < A.g.vol * 3; ++i) reaction[i] = 0.0f;          // lbm_particles_couple clears the field first (lbm_api.cu)
    run((unsigned)ps->n, 256, [&] { particles_couple_kernel(A); });
    return 0;
}
// z-slab: nz owned planes from global plane z0, one ghost plane per side (u and reaction are [3][nz + 2][ny][nx])
int emu_particles_couple_slab(int nx, int ny, int nz, int z0, int nz_global, const float *u, float *reaction, lbm_particles *ps, float rho_w,
                              float mu_w, float relax) {
    ParticleArgs A{make_grid(nx, ny, nz, 1, z0, nz_global), u, reaction, *ps, rho_w, mu_w, relax};
    for (long long i = 0; i < A.g.vol * 3; ++i) reaction[i] = 0.0f;
    run((unsigned)ps->n, 256, [&] { particles_couple_kernel(A); });
    return 0;
}
int emu_particles_under_relax(lbm_particles *ps, float relax) {
    run((unsigned)ps->n, 256, [&] { particles_under_relax_kernel(*ps, relax); });
    return 0;
}
int emu_particles_advance(lbm_particles *ps, float *force, const lbm_particle_bounds *b, float dt, int *counters) {
    run((unsigned)ps->n, 256, [&] { particles_advance_kernel(*ps, force, *b, dt, counters); });
    return 0;
}
int emu_particles_fluid_forces(int nx, int ny, int nz, const float *u, lbm_particles *ps, float *force, double water_density, double water_viscosity,
                               double gravity, int *counters) {
    const Grid G = make_grid(nx, ny, nz);
    const float max_coord = (float)std::max(nx, std::max(ny, nz));
    const float mu_safe = (float)std::max(1e-8, water_viscosity), vol_k = (float)((4.0 / 3.0) * 3.14159);   // as lbm_api.cu folds them
    run((unsigned)ps->n, 256, [&] { particles_fluid_forces_kernel(G, u, *ps, force, (float)water_density, mu_safe, (float)gravity, vol_k, max_coord, counters); });
    return 0;
}
// z-slab: nz owned planes from global plane z0, one ghost plane per side (u is [3][nz + 2][ny][nx])
int emu_particles_fluid_forces_slab(int nx, int ny, int nz, int z0, int nz_global, const float *u, lbm_particles *ps, float *force, double water_density,
                                    double water_viscosity, double gravity, int *counters) {
    const Grid G = make_grid(nx, ny, nz, 1, z0, nz_global);
    const float max_coord = (float)std::max(nx, std::max(ny, nz_global));
    const float mu_safe = (float)std::max(1e-8, water_viscosity), vol_k = (float)((4.0 / 3.0) * 3.14159);
    run((unsigned)ps->n, 256, [&] { particles_fluid_forces_kernel(G, u, *ps, force, (float)water_density, mu_safe, (float)gravity, vol_k, max_coord, counters); });
    return 0;
}
}  // extern "C"
