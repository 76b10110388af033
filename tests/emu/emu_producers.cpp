// TEST INFRASTRUCTURE -- not product code.  Compiles the product's CUDA kernel bodies of
// pour_over_coffee_lbm_b200/csrc/lbm_producers.cu with the HOST compiler (the CUDA qualifiers expand to nothing under g++,
// blockIdx / threadIdx / blockDim become plain variables) and executes a launch as nested loops over blocks and threads.
// Purpose: the authoring container has no GPU, so this is how the kernel source itself -- VEC = 4 and VEC = 1 paths, masks,
// edge lanes, launch geometry helpers -- is checked against the recorded reference runs before it is sent to a B200.
// The kernels here have no intra-launch dependencies between threads (disjoint writes, atomics of one constant), so
// sequential execution is a valid schedule.  Only tests/test_producers_emulated.py builds and loads this file.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

struct EmuIdx { unsigned x, y, z; };
static EmuIdx emu_block_idx, emu_thread_idx, emu_block_dim;
#define blockIdx emu_block_idx
#define threadIdx emu_thread_idx
#define blockDim emu_block_dim
static inline float atomicAdd(float *p, float v) { const float o = *p; *p = o + v; return o; }

#define __launch_bounds__(...)
#define LBM_EMULATE_ON_HOST 1
#include "../../pour_over_coffee_lbm_b200/csrc/lbm_producers.cu"

using namespace lbm;

template <class F>
static void run(dim3 grid, unsigned block, F &&kernel) {
    emu_block_dim = {block, 1, 1};
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx)
                for (unsigned t = 0; t < block; ++t) {
                    emu_block_idx = {bx, by, bz};
                    emu_thread_idx = {t, 0, 0};
                    kernel();
                }
}
static Grid make_grid(int nx, int ny, int nz, int zg = 0, int z0 = 0, int nz_global = 0) {
    Grid G{};
    G.nx = nx; G.ny = ny; G.nz = nz; G.zg = zg; G.nz_global = nz_global > 0 ? nz_global : nz; G.z0 = z0;
    G.plane = (long long)nx * ny; G.vol = G.plane * (nz + 2 * zg);
    return G;
}
#define RUN_CELLS(kernel, vec, ...)                                                                          \
    do {                                                                                                     \
        if ((vec) == 4) { const int b_ = cell_block(G, 4); run(cell_grid(G, 4, b_), b_, [&] { kernel<4>(__VA_ARGS__); }); } \
        else { const int b_ = cell_block(G, 1); run(cell_grid(G, 1, b_), b_, [&] { kernel<1>(__VA_ARGS__); }); }            \
    } while (0)

extern "C" {

int emu_chemical_potential(int vec, int nx, int ny, int nz, const float *phi, float *lap, float *mu, float kappa) {
    const Grid G = make_grid(nx, ny, nz);
    RUN_CELLS(mp_chemical_potential_kernel, vec, G, phi, lap, mu, kappa);
    return 0;
}
int emu_surface_tension(int vec, int nx, int ny, int nz, const float *phi, const float *mu, const float *rho, const uint8_t *flags, float *grad_phi,
                        float *grad_mu, float *normal, float *curvature, float *surface_force, float *body_force, float sigma) {
    const Grid G = make_grid(nx, ny, nz);
    RUN_CELLS(mp_gradients_kernel, vec, G, phi, mu, grad_phi, grad_mu, normal);
    RUN_CELLS(mp_curvature_force_kernel, vec, G, phi, rho, flags, grad_phi, normal, curvature, surface_force, body_force, sigma);
    return 0;
}
// z-slab variants: nz owned planes starting at global plane z0, one ghost plane per side (fields are [nz + 2][ny][nx])
int emu_slab_gradients(int vec, int nx, int ny, int nz, int z0, int nz_global, const float *phi, const float *mu, float *grad_phi, float *grad_mu,
                       float *normal) {
    const Grid G = make_grid(nx, ny, nz, 1, z0, nz_global);
    RUN_CELLS(mp_gradients_kernel, vec, G, phi, mu, grad_phi, grad_mu, normal);
    return 0;
}
int emu_slab_curvature_force(int vec, int nx, int ny, int nz, int z0, int nz_global, const float *phi, const float *rho, const uint8_t *flags,
                             const float *grad_phi, const float *normal, float *curvature, float *surface_force, float *body_force, float sigma) {
    const Grid G = make_grid(nx, ny, nz, 1, z0, nz_global);
    RUN_CELLS(mp_curvature_force_kernel, vec, G, phi, rho, flags, grad_phi, normal, curvature, surface_force, body_force, sigma);
    return 0;
}
int emu_slab_phase_field_step(int vec, int nx, int ny, int nz, int z0, int nz_global, float *phi, float *phi_new, const float *mu, const float *u,
                              float *rho, float *phase, float mobility, float dt, double rho_water, double rho_air) {
    const Grid G = make_grid(nx, ny, nz, 1, z0, nz_global);
    RUN_CELLS(mp_phase_update_kernel, vec, G, phi, mu, u, phi_new, mobility, dt);
    RUN_CELLS(mp_copy_density_kernel, vec, G, phi_new, phi, rho, phase, (float)rho_air, (float)(rho_water - rho_air));
    return 0;
}
int emu_surface_tension_lean(int vec, int nx, int ny, int nz, const float *phi, const float *rho, const uint8_t *flags, const float *normal_outer,
                             const float *force_outer, float *body_force, float sigma) {
    const Grid G = make_grid(nx, ny, nz);
    RUN_CELLS(mp_surface_tension_lean_kernel, vec, G, phi, rho, flags, normal_outer, force_outer, body_force, sigma);
    return 0;
}
int emu_apply_surface_tension(int vec, int nx, int ny, int nz, const float *surface_force, const float *rho, const uint8_t *flags, float *body_force) {
    const Grid G = make_grid(nx, ny, nz);
    RUN_CELLS(mp_apply_surface_tension_kernel, vec, G, surface_force, rho, flags, body_force);
    return 0;
}
int emu_phase_field_step(int vec, int nx, int ny, int nz, float *phi, float *phi_new, const float *mu, const float *u, float *rho, float *phase,
                         float mobility, float dt, double rho_water, double rho_air) {
    const Grid G = make_grid(nx, ny, nz);
    RUN_CELLS(mp_phase_update_kernel, vec, G, phi, mu, u, phi_new, mobility, dt);
    RUN_CELLS(mp_copy_density_kernel, vec, G, phi_new, phi, rho, phase, (float)rho_air, (float)(rho_water - rho_air));
    return 0;
}
int emu_density_from_phase(int vec, int nx, int ny, int nz, const float *phi, float *rho, float *phase, double rho_water, double rho_air) {
    const Grid G = make_grid(nx, ny, nz);
    RUN_CELLS(mp_copy_density_kernel, vec, G, phi, const_cast<float *>(phi), rho, phase, (float)rho_air, (float)(rho_water - rho_air));
    return 0;
}
int emu_dynamic_resistance(int cells, int nx, int ny, int nz, const uint8_t *flags, float *blockage, float *accumulated) {
    const Grid G = make_grid(nx, ny, nz);
    const long long begin = (long long)G.zg * G.plane, count = (long long)G.nz * G.plane;
    if (cells == 0) cells = dynamic_resistance_cells(begin, count, flags, false);      // what the launcher would pick
    const long long threads = (count + cells - 1) / cells;
    const dim3 grid((unsigned)((threads + 255) / 256), 1, 1);
    if (cells == 16) run(grid, 256, [&] { dynamic_resistance_kernel<16>(begin, count, flags, blockage, accumulated); });
    else if (cells == 4) run(grid, 256, [&] { dynamic_resistance_kernel<4>(begin, count, flags, blockage, accumulated); });
    else run(grid, 256, [&] { dynamic_resistance_kernel<1>(begin, count, flags, blockage, accumulated); });
    return cells;
}
int emu_particles_block_at_filter(int nx, int ny, int nz, int n, float *pos, float *vel, int32_t *active, const uint8_t *flags, float *accumulated,
                                  float scale_length, float noise, unsigned seed) {
    const Grid G = make_grid(nx, ny, nz);
    lbm_particles P{};
    P.pos = pos; P.vel = vel; P.active = active; P.n = n;
    const unsigned b = 256;
    run(dim3((n + b - 1) / b, 1, 1), b, [&] { particles_block_at_filter_kernel(G, P, flags, accumulated, scale_length, noise, seed); });
    return 0;
}
// z-slab: nz owned planes from global plane z0, one ghost plane per side (flags and accumulated are [nz + 2][ny][nx])
int emu_particles_block_at_filter_slab(int nx, int ny, int nz, int z0, int nz_global, lbm_particles *ps, const uint8_t *flags, float *accumulated,
                                       float scale_length, float noise, unsigned seed) {
    const Grid G = make_grid(nx, ny, nz, 1, z0, nz_global);
    const unsigned b = 256;
    run(dim3((ps->n + b - 1) / b, 1, 1), b, [&] { particles_block_at_filter_kernel(G, *ps, flags, accumulated, scale_length, noise, seed); });
    return 0;
}
int emu_pour(int mode, int nx, int ny, int nz, float pour_x, float pour_y, float radius, int pour_z, float velocity, float flow_rate, float dt,
             const uint8_t *flags, float *field) {
    const Grid G = make_grid(nx, ny, nz);
    PourArgs P{};
    P.pour_x = pour_x; P.pour_y = pour_y; P.radius = radius; P.pour_z = pour_z; P.velocity = velocity; P.flow_rate = flow_rate; P.dt = dt;
    for (int d = 0; d < 5; ++d) P.decay[d] = (float)exp(-(double)d / 2.0);
    const int x0 = std::max(0, (int)floorf(pour_x - radius)), x1 = std::min(nx - 1, (int)ceilf(pour_x + radius));
    const int y0 = std::max(0, (int)floorf(pour_y - radius)), y1 = std::min(ny - 1, (int)ceilf(pour_y + radius));
    const int k0 = std::max(0, pour_z - 4), k1 = std::min(nz - 1, pour_z);
    if (x1 < x0 || y1 < y0 || k1 < k0) return 0;
    P.x0 = x0; P.y0 = y0; P.k0 = k0; P.wx = x1 - x0 + 1; P.wy = y1 - y0 + 1; P.wk = k1 - k0 + 1;
    const unsigned b = 128, cells = (unsigned)(P.wx * P.wy * P.wk);
    if (mode == 0) run(dim3((cells + b - 1) / b, 1, 1), b, [&] { pour_kernel<0>(G, P, flags, field); });
    else run(dim3((cells + b - 1) / b, 1, 1), b, [&] { pour_kernel<1>(G, P, flags, field); });
    return 0;
}

}  // extern "C"
