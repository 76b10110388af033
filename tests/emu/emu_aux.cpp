// TEST INFRASTRUCTURE -- not product code.  The plain kernels of pour_over_coffee_lbm_b200/csrc/lbm_aux.cu (V60 geometry,
// flag packing, neighbour masks, exact f <-> g conversion, face density writes, pressure-gradient and Forchheimer forces,
// reaction accumulation, equilibrium initialisation) compiled by the HOST compiler and executed thread by thread (see
// emu_producers.cpp).  Grid-stride kernels run as one block of one thread.  The cub-based work-list builder, the packed-math
// self-test and the shuffle-based statistics stay GPU-only.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

struct EmuIdx { unsigned x, y, z; };
static EmuIdx emu_block_idx, emu_thread_idx, emu_block_dim, emu_grid_dim;
#define blockIdx emu_block_idx
#define threadIdx emu_thread_idx
#define blockDim emu_block_dim
#define gridDim emu_grid_dim
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
static inline float atomicAdd(float *p, float v) { const float o = *p; *p = o + v; return o; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int min(int a, int b) { return a < b ? a : b; }               // CUDA's integer min / max device functions
static inline int max(int a, int b) { return a > b ? a : b; }
#define __launch_bounds__(...)
#define LBM_EMULATE_ON_HOST 1
#define LBM_PHYS_COLLISION_ONLY 1
#include "../../pour_over_coffee_lbm_b200/csrc/lbm_aux.cu"

using namespace lbm;

template <class F>
static void run(dim3 grid, unsigned block, F &&kernel) {
    emu_block_dim = {block, 1, 1}; emu_grid_dim = {grid.x, grid.y, grid.z};
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx)
                for (unsigned t = 0; t < block; ++t) { emu_block_idx = {bx, by, bz}; emu_thread_idx = {t, 0, 0}; kernel(); }
}
template <class F> static void run_stride(F &&kernel) { run(dim3(1, 1, 1), 1, kernel); }     // grid-stride loop: one thread walks everything
static Grid make_grid(int nx, int ny, int nz, int periodic = 0, int zg = 0, int z0 = 0, int nz_global = 0) {
    Grid G{};
    G.nx = nx; G.ny = ny; G.nz = nz; G.zg = zg; G.nz_global = nz_global > 0 ? nz_global : nz; G.z0 = z0;
    G.per_x = periodic & 1; G.per_y = (periodic >> 1) & 1; G.per_z = (periodic >> 2) & 1;
    G.plane = (long long)nx * ny; G.vol = G.plane * (nz + 2 * zg);
    return G;
}

extern "C" {
int emu_v60_geometry(int nx, int ny, int nz, uint8_t *solid, int32_t *zone, const float *geom5) {
    const Grid G = make_grid(nx, ny, nz);
    run_stride([&] { v60_geometry_kernel(G, solid, zone, geom5[0], geom5[1], geom5[2], geom5[3], geom5[4]); });
    return 0;
}
int emu_pack_flags_and_masks(int nx, int ny, int nz, int periodic, uint8_t *flags, const uint8_t *solid, const int32_t *zone, const int32_t *les,
                             unsigned long long *nbr) {
    const Grid G = make_grid(nx, ny, nz, periodic);
    run_stride([&] { pack_flags_kernel(G, flags, solid, zone, les); });
    run_stride([&] { neighbour_mask_kernel(G, flags, nbr); });
    return 0;
}
int emu_convert_f(int nx, int ny, int nz, int to_reference_f, const float *in, const uint8_t *flags, float *out) {
    const Grid G = make_grid(nx, ny, nz);
    if (to_reference_f) run_stride([&] { convert_f_kernel<true>(G, in, flags, out); });
    else run_stride([&] { convert_f_kernel<false>(G, in, flags, out); });
    return 0;
}
// z-slab variants (one ghost plane per side; fields are [nz + 2][ny][nx]; the ghost planes of `solid` hold the neighbours' mask)
int emu_slab_pack_flags_and_masks(int nx, int ny, int nz, int z0, int nz_global, uint8_t *flags, const uint8_t *solid, const int32_t *zone,
                                  const int32_t *les, unsigned long long *nbr) {
    const Grid G = make_grid(nx, ny, nz, 0, 1, z0, nz_global);
    run_stride([&] { pack_flags_kernel(G, flags, solid, zone, les); });
    run_stride([&] { neighbour_mask_kernel(G, flags, nbr); });
    return 0;
}
int emu_slab_convert_f(int nx, int ny, int nz, int z0, int nz_global, int to_reference_f, const float *in, const uint8_t *flags, float *out) {
    const Grid G = make_grid(nx, ny, nz, 0, 1, z0, nz_global);
    if (to_reference_f) run_stride([&] { convert_f_kernel<true>(G, in, flags, out); });
    else run_stride([&] { convert_f_kernel<false>(G, in, flags, out); });
    return 0;
}
int emu_face_bc(int nx, int ny, int nz, float *rho, const uint8_t *flags) {
    const Grid G = make_grid(nx, ny, nz);
    const unsigned b = 128;
    for (int pass = 0; pass < 5; ++pass) {                                    // launch geometry of launch_face_bc
        dim3 grid;
        if (pass == 0 || pass == 1 || pass == 4) grid = dim3((nx + b - 1) / b, ny);
        else if (pass == 2) grid = dim3((ny + b - 1) / b, nz);
        else grid = dim3((nx + b - 1) / b, nz);
        run(grid, b, [&] { face_bc_kernel(G, rho, flags, pass); });
    }
    return 0;
}
int emu_pressure_gradient(int nx, int ny, int nz, const float *rho, const uint8_t *flags, float *bf, float max_force, float scale, int accumulate) {
    const Grid G = make_grid(nx, ny, nz);
    const unsigned b = nx >= 128 ? 128 : 64;
    run(dim3((nx + b - 1) / b, ny, nz), b, [&] { pressure_gradient_kernel(G, rho, flags, bf, max_force, scale, accumulate); });
    return 0;
}
// the tile-list variant the library uses when the caller's flag field is the one its work lists were built for: one warp per
// active warp-tile (32 * vec x-consecutive cells of a row holding at least one fluid cell)
int emu_pressure_gradient_tiles(int nx, int ny, int nz, int vec, const float *rho, const uint8_t *flags, float *bf, float max_force, float scale,
                                int accumulate) {
    const Grid G = make_grid(nx, ny, nz);
    std::vector<unsigned> items;
    const int span = 32 * vec, segs = (nx + span - 1) / span;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int sgm = 0; sgm < segs; ++sgm) {
                bool any = false;
                for (int x = span * sgm; x < nx && x < span * (sgm + 1); ++x) any |= !(flags[((long long)z * ny + y) * nx + x] & LBM_FLAG_SOLID);
                if (any) items.push_back((unsigned)sgm | ((unsigned)y << 8) | ((unsigned)z << 20));
            }
    const int n_items = (int)items.size();
    run(dim3((n_items + 3) / 4, 1, 1), 128, [&] { pressure_gradient_tiles_kernel(G, rho, flags, bf, max_force, scale, accumulate, items.data(), n_items, vec); });
    return n_items;
}
int emu_forchheimer(int nx, int ny, int nz, const float *u, const uint8_t *flags, float *bf, float K, float beta, float c_darcy, float c_forch,
                    float fmax) {
    const Grid G = make_grid(nx, ny, nz);
    const unsigned b = 64;             // the launch geometry of launch_forchheimer_force: (x-chunk, y, z), 4 cells per thread when nx % 4 == 0
    if (nx % 4 == 0) run(dim3((unsigned)((nx / 4 + b - 1) / b), (unsigned)ny, (unsigned)nz), b, [&] { forchheimer_force_kernel<4>(G, u, flags, bf, K, beta, c_darcy, c_forch, fmax); });
    else run(dim3((unsigned)((nx + b - 1) / b), (unsigned)ny, (unsigned)nz), b, [&] { forchheimer_force_kernel<1>(G, u, flags, bf, K, beta, c_darcy, c_forch, fmax); });
    return 0;
}
int emu_bounce_slots(int nx, int ny, int nz, int periodic, float *g, const uint8_t *flags, const unsigned long long *nbr) {
    const Grid G = make_grid(nx, ny, nz, periodic);
    run_stride([&] { bounce_slots_kernel(G, g, flags, nbr, 0, nz); });
    return 0;
}
// packed quad list + wall links of the four-cell walls kernel (build_chord_lists without cub: the exclusive sums and the
// per-plane padding run here on the host).  Returns the number of tiles; quads_out = 32 u64 per tile, tile_links_out = 2 u32 per
// tile, links_out up to max_links u32, tile_off_out = nz + 1 ints.
int emu_chord_lists(int nx, int ny, int nz, int periodic, const uint8_t *flags, const unsigned long long *nbr, unsigned long long *quads_out,
                    int max_tiles, unsigned *tile_links_out, unsigned *links_out, int max_links, int *n_links_out, int *tile_off_out) {
    const Grid G = make_grid(nx, ny, nz, periodic);
    const int rows = nz * ny;
    std::vector<int> cnt(rows + 1, 0), off(rows + 1, 0), plane_base(nz + 1, 0);
    run(dim3((rows + 127) / 128, 1, 1), 128, [&] { quad_count_kernel(G, flags, cnt.data()); });
    for (int r = 0; r < rows; ++r) off[r + 1] = off[r] + cnt[r];
    tile_off_out[0] = 0;
    for (int z = 0; z < nz; ++z) {
        const int nq = off[(z + 1) * ny] - off[z * ny];
        tile_off_out[z + 1] = tile_off_out[z] + (nq + 31) / 32;
        plane_base[z + 1] = tile_off_out[z + 1] * 32;
    }
    const int n_t = tile_off_out[nz];
    if (n_t > max_tiles) return -1;
    for (long long i = 0; i < (long long)n_t * 32; ++i) quads_out[i] = 0;
    run(dim3((rows + 127) / 128, 1, 1), 128, [&] { quad_fill_kernel(G, flags, off.data(), plane_base.data(), quads_out); });
    std::vector<int> tc(n_t + 1, 0), lo(n_t + 1, 0);
    run(dim3((n_t + 127) / 128, 1, 1), 128, [&] { quad_links_kernel(G, flags, nbr, quads_out, n_t, 0, tc.data(), nullptr, nullptr, nullptr); });
    for (int t = 0; t < n_t; ++t) lo[t + 1] = lo[t] + tc[t];
    if (lo[n_t] > max_links) return -2;
    run(dim3((n_t + 127) / 128, 1, 1), 128, [&] { quad_links_kernel(G, flags, nbr, quads_out, n_t, 1, nullptr, lo.data(), reinterpret_cast<uint2 *>(tile_links_out), links_out); });
    *n_links_out = lo[n_t];
    return n_t;
}
// the value waiting on every wall link, taken from the populations (wall_values_kernel)
int emu_wall_values(int nx, int ny, int nz, int periodic, const float *g, const unsigned long long *quads, const unsigned *tile_links, const unsigned *links,
                    float *wall, int n_tiles) {
    const Grid G = make_grid(nx, ny, nz, periodic);
    run(dim3((n_tiles + 127) / 128, 1, 1), 128, [&] { wall_values_kernel(G, g, quads, reinterpret_cast<const uint2 *>(tile_links), links, wall, n_tiles); });
    return 0;
}
// pressure-gradient producer over the packed quad list
int emu_pressure_gradient_chord(int nx, int ny, int nz, const float *rho, const uint8_t *flags, float *bf, float max_force, float scale, int accumulate,
                                const unsigned long long *quads, int n_tiles) {
    const Grid G = make_grid(nx, ny, nz);
    run(dim3((n_tiles + 3) / 4, 1, 1), 128, [&] { pressure_gradient_chord_kernel(G, rho, flags, bf, max_force, scale, accumulate, quads, n_tiles); });
    return 0;
}
int emu_add_reaction(int nx, int ny, int nz, const float *reaction, const uint8_t *flags, float *bf) {
    const Grid G = make_grid(nx, ny, nz);
    if (G.vol % 4 == 0) run_stride([&] { add_reaction_kernel<4>(G, reaction, flags, bf); });
    else run_stride([&] { add_reaction_kernel<1>(G, reaction, flags, bf); });
    return 0;
}
}  // extern "C"
