// Initialisation, geometry, flag packing, f<->g conversion, face BCs and the small
// stencil kernels that feed body_force.  Compiled with -fmad=false so that every value is
// bit-identical to oracle/d3q19_ref.py (these kernels are not on the roofline path).
#ifndef LBM_EMULATE_ON_HOST   /* tests/emu compiles the plain kernels of this file with the host compiler */
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#endif

#include <vector>

#include "lbm_common.cuh"
#include "lbm_phys.cuh"

namespace lbm {

// ---- equilibrium initialisation ---------------------------------------------------------
// legacy/lbm_solver.py:1067-1112, lbm_unified.py:236-248 (mode's own equilibrium table).
template <int COMPAT>
__global__ void init_equilibrium_kernel(Grid G, float *g, const float *rho, const float *u, float rho0,
                                        float u0x, float u0y, float u0z) {
    const long long n = G.vol;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const float r = rho ? rho[c] : rho0;
        const float ux = u ? u[c] : u0x, uy = u ? u[n + c] : u0y, uz = u ? u[2 * n + c] : u0z;
        const float u_sq = dot3(ux, uy, uz, ux, uy, uz);
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            float eu;
            if constexpr (COMPAT == LBM_COMPAT_REFERENCE) eu = edot<ex(q), ey(q), ez(q)>(ux, uy, uz);
            else eu = edot<cx(q), cy(q), cz(q)>(ux, uy, uz);
            g[(long long)q * n + c] = (wq(q) * r) * (((1.0f + 3.0f * eu) + (4.5f * eu) * eu) - 1.5f * u_sq);
        });
    }
}

// ---- V60 geometry ---------------------------------------------------------------------------
// FilterPaperSystem._setup_v60_geometry / _setup_filter_zones, filter_paper.py:206-364.
// Predicates in f32 without contraction; constants arrive f32-rounded from the host.
__global__ void v60_geometry_kernel(Grid G, uint8_t *solid, int32_t *zone, float top_r, float bot_r, float cup_h,
                                    float gap, float thick) {
    const long long n = G.vol;
    const float cxf = (float)(G.nx * 0.5), cyf = (float)(G.ny * 0.5);
    const float bottom_z = 5.0f, wall = 2.0f;
    const float top_z = bottom_z + cup_h;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(c % G.nx);
        const int y = (int)((c / G.nx) % G.ny);
        const int zp = (int)(c / G.plane);
        const int k = G.z0 + zp - G.zg;                 // global z (ghost planes included)
        const float dx = (float)x - cxf, dy = (float)y - cyf;
        const float r = sqrtf(dx * dx + dy * dy);
        const float z = (float)k;
        if (solid) {
            bool is = false;
            if (z <= bottom_z) { if (r > bot_r) is = true; }
            else if (z <= top_z) {
                const float hr = (z - bottom_z) / cup_h;
                const float inner = bot_r + (top_r - bot_r) * hr;
                if (r > (inner + gap) + wall) is = true;
            } else { if (r > top_r + wall) is = true; }
            if (x <= 2 || x >= G.nx - 3 || y <= 2 || y >= G.ny - 3 || k <= 2 || k >= G.nz_global - 3) is = true;
            if (k < 0 || k >= G.nz_global) is = true;   // ghost planes outside the global box
            solid[c] = is ? 1 : 0;
        }
        if (zone) {
            int zn = 0;
            const float f_bot = 5.0f, f_top = 5.0f + cup_h;
            if (z >= f_bot && z <= f_top) {
                float hr = (z - f_bot) / cup_h;
                hr = fmaxf(0.0f, fminf(1.0f, hr));
                const float inner_r = bot_r + (top_r - bot_r) * hr;
                const float f_out = inner_r - gap;
                const float f_in = f_out - thick;
                if (f_in <= r && r <= f_out) zn = 1;
            } else if (z >= f_bot - thick && z < f_bot) {
                if (r <= bot_r - gap) zn = 1;
            }
            zone[c] = zn;
        }
    }
}

// ---- flag packing ---------------------------------------------------------------------------
__device__ __forceinline__ bool wrap_or_oob(int &v, int n, int per) {
    if (v < 0) { v = n - 1; return !per; }
    if (v >= n) { v = 0; return !per; }
    return false;
}

__global__ void pack_flags_kernel(Grid G, uint8_t *flags, const uint8_t *solid, const int32_t *zone, const int32_t *les) {
    const long long n = G.vol;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(c % G.nx);
        const int y = (int)((c / G.nx) % G.ny);
        const int zp = (int)(c / G.plane);
        const int z = zp - G.zg;
        unsigned f = 0;
        if (solid[c]) f |= LBM_FLAG_SOLID;
        if (zone && zone[c] == 1) f |= LBM_FLAG_FILTER;
        if (!les || les[c] != 0) f |= LBM_FLAG_LES;
        bool near = false;
        if (z >= 0 && z < G.nz) {
            for (int q = 1; q < Q; ++q) {
                int xs = x - cx(q), ys = y - cy(q), zs = z - cz(q);
                bool oob = wrap_or_oob(xs, G.nx, G.per_x) | wrap_or_oob(ys, G.ny, G.per_y);
                const int zs_g = G.z0 + zs;
                if (zs_g < 0 || zs_g >= G.nz_global) oob |= !G.per_z;
                int zsp = zs + G.zg;
                if (!G.zg) { if (zs < 0) zsp = G.nz - 1; else if (zs >= G.nz) zsp = 0; }
                if (oob) { near = true; break; }
                if (solid[((long long)zsp * G.ny + ys) * G.nx + xs]) { near = true; break; }
            }
        }
        if (near) f |= LBM_FLAG_NEAR;
        flags[c] = (uint8_t)f;
    }
}

// ---- exact f <-> g conversion ---------------------------------------------------------------
// export: f[q,x] = g[q,x-e] (fluid source) | g[opp q,x] (solid source) | w_q (source outside an open face)
// import: g[q,x] = f[q,x+e] (fluid target) | f[opp q,x] (solid target) | f[q,x] (target outside: dropped later)
template <bool EXPORT>
__global__ void convert_f_kernel(Grid G, const float *in, const uint8_t *flags, float *out) {
    const long long n = G.vol;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(c % G.nx);
        const int y = (int)((c / G.nx) % G.ny);
        const int zp = (int)(c / G.plane);
        const int z = zp - G.zg;
        const bool ghost = z < 0 || z >= G.nz;
        const bool self_solid = flags && (flags[c] & LBM_FLAG_SOLID);
        for (int q = 0; q < Q; ++q) {
            float v = in[(long long)q * n + c];
            if (q > 0 && !ghost && !self_solid) {
                const int s = EXPORT ? -1 : 1;
                int xs = x + s * cx(q), ys = y + s * cy(q), zs = z + s * cz(q);
                bool oob = wrap_or_oob(xs, G.nx, G.per_x) | wrap_or_oob(ys, G.ny, G.per_y);
                const int zs_g = G.z0 + zs;
                if (zs_g < 0 || zs_g >= G.nz_global) oob |= !G.per_z;
                int zsp = zs + G.zg;
                if (!G.zg) { if (zs < 0) zsp = G.nz - 1; else if (zs >= G.nz) zsp = 0; }
                if (oob) { if (EXPORT) v = wq(q); }
                else {
                    const long long nb = ((long long)zsp * G.ny + ys) * G.nx + xs;
                    if (flags && (flags[nb] & LBM_FLAG_SOLID)) v = in[(long long)opp(q) * n + c];
                    else v = in[(long long)q * n + nb];
                }
            }
            out[(long long)q * n + c] = v;
        }
    }
}

// ---- face density writes ----------------------------------------------------------------------
// boundary_conditions.py:178-324 (SoA branch): only rho is observable (quirk Q5).  One launch
// per serial Taichi offload: top, bottom, x faces, y faces, z=0 again.
__global__ void face_bc_kernel(Grid G, float *rho, const uint8_t *flags, int pass) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    auto idx = [&](int x, int y, int zglob) { return ((long long)(zglob - G.z0 + G.zg) * G.ny + y) * G.nx + x; };
    auto fluid = [&](long long c) { return !(flags && (flags[c] & LBM_FLAG_SOLID)); };
    auto own = [&](int zglob) { return zglob >= G.z0 && zglob < G.z0 + G.nz; };
    const int NZ = G.nz_global;
    if (pass == 0) {            // top
        if (a < G.nx && b < G.ny && own(NZ - 1)) { long long c = idx(a, b, NZ - 1); if (fluid(c)) rho[c] = 1.0f; }
    } else if (pass == 1 || pass == 4) {   // bottom / outlet bottom
        if (a < G.nx && b < G.ny && own(0)) { long long c = idx(a, b, 0); if (fluid(c)) rho[c] = rho[idx(a, b, 1)]; }
    } else if (pass == 2) {     // x faces: a = y, b = local z
        if (a < G.ny && b < G.nz) {
            const int zg_ = G.z0 + b;
            long long c0 = idx(0, a, zg_), c1 = idx(G.nx - 1, a, zg_);
            if (fluid(c0)) rho[c0] = rho[idx(1, a, zg_)];
            if (fluid(c1)) rho[c1] = rho[idx(G.nx - 2, a, zg_)];
        }
    } else if (pass == 3) {     // y faces: a = x, b = local z
        if (a < G.nx && b < G.nz) {
            const int zg_ = G.z0 + b;
            long long c0 = idx(a, 0, zg_), c1 = idx(a, G.ny - 1, zg_);
            if (fluid(c0)) rho[c0] = rho[idx(a, 1, zg_)];
            if (fluid(c1)) rho[c1] = rho[idx(a, G.ny - 2, zg_)];
        }
    }
}

// ---- neighbours feeding body_force ------------------------------------------------------------
// pressure_gradient_drive.py:124-193, 274-279
// One thread per cell on a (x-chunk, y, owned z) grid: coordinates come from the block index (the first version derived
// them from a linear index with three 64-bit divisions per thread and took 1.14 ms on a 512^3 V60 box).
// accumulate = 0 writes body_force = F on fluid cells instead of adding to it (saves the caller's clear pass).
__device__ __forceinline__ void pressure_gradient_cell(const Grid &G, const float *rho, const uint8_t *flags, float *bf, float max_force,
                                                       float scale, int accumulate, int x, int y, int z) {
    const long long n = G.vol;
    const long long c = ((long long)(z + G.zg) * G.ny + y) * G.nx + x;
    if (flags && (flags[c] & LBM_FLAG_SOLID)) return;
    const int k = G.z0 + z;
    const float r0 = rho[c];
    const float gx = pressure_gradient_diff(r0, x > 0 ? rho[c - 1] : r0, x < G.nx - 1 ? rho[c + 1] : r0, x == 0 ? -1 : (x == G.nx - 1 ? 1 : 0));
    const float gy = pressure_gradient_diff(r0, y > 0 ? rho[c - G.nx] : r0, y < G.ny - 1 ? rho[c + G.nx] : r0, y == 0 ? -1 : (y == G.ny - 1 ? 1 : 0));
    const float gz = pressure_gradient_diff(r0, k > 0 ? rho[c - G.plane] : r0, k < G.nz_global - 1 ? rho[c + G.plane] : r0,
                                            k == 0 ? -1 : (k == G.nz_global - 1 ? 1 : 0));
    float fx, fy, fz;
    pressure_gradient_value(r0, gx, gy, gz, max_force, scale, fx, fy, fz);
    if (accumulate) { bf[c] = bf[c] + fx; bf[n + c] = bf[n + c] + fy; bf[2 * n + c] = bf[2 * n + c] + fz; }
    else { bf[c] = fx; bf[n + c] = fy; bf[2 * n + c] = fz; }
}
// One thread per cell on a (x-chunk, y, owned z) grid: coordinates come from the block index (the first version derived
// them from a linear index with three 64-bit divisions per thread and took 1.14 ms on a 512^3 V60 box).
// accumulate = 0 writes body_force = F on fluid cells instead of adding to it (saves the caller's clear pass).
__global__ void pressure_gradient_kernel(Grid G, const float *rho, const uint8_t *flags, float *bf, float max_force, float scale, int accumulate) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= G.nx) return;
    pressure_gradient_cell(G, rho, flags, bf, max_force, scale, accumulate, x, blockIdx.y, blockIdx.z);
}
// Same over the step kernel's active warp-tile list (one warp per tile of 32*vec cells): the solid 65 % of a V60 box is
// never visited.
__global__ void pressure_gradient_tiles_kernel(Grid G, const float *rho, const uint8_t *flags, float *bf, float max_force, float scale,
                                               int accumulate, const unsigned *items, int n_items, int vec) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_items) return;
    const unsigned e = __ldg(items + w);
    const int y = (int)((e >> 8) & 0xfffu), z = (int)(e >> 20);
    const int xb = (int)(e & 0xffu) * 32 * vec + (int)(threadIdx.x & 31u);
    for (int i = 0; i < vec; ++i) {
        const int x = xb + 32 * i;
        if (x < G.nx) pressure_gradient_cell(G, rho, flags, bf, max_force, scale, accumulate, x, y, z);
    }
}

// Same over the packed quad list of the four-cell walls kernel: one thread per listed quad, the four flag bytes as one word, rho and
// its y / z neighbours as 128-bit vectors, the force of an all-fluid quad as three 128-bit stores (the cell-at-a-time form issued
// every load and store with a 16-byte lane stride: 0.42 ms at V60 512^3 for 0.9 GB).  Same statements per cell as
// pressure_gradient_cell, so the values are the same bits.
__global__ void pressure_gradient_chord_kernel(Grid G, const float *rho, const uint8_t *flags, float *bf, float max_force, float scale,
                                               int accumulate, const unsigned long long *quads, int n_items) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)n_items * 32) return;
    const unsigned long long e = quads[i];
    if (!(e & (1ull << 44))) return;
    const int xb = (int)(e & 0xfffu) * 4, y = (int)((e >> 12) & 0xffffu), z = (int)((e >> 28) & 0xffffu);
    const long long n = G.vol;
    const long long c = ((long long)(z + G.zg) * G.ny + y) * G.nx + xb;
    // every load is issued before anything branches on the flag word (a flags -> branch -> rho chain is one more DRAM round trip per
    // thread, and 99 % of the listed quads hold a fluid cell); a neighbour outside the box re-reads the quad itself
    const unsigned fw = flags ? __ldg(reinterpret_cast<const unsigned *>(flags + c)) : 0u;
    constexpr unsigned SOLID4 = 0x01010101u * LBM_FLAG_SOLID;
    const int k = G.z0 + z;
    const float *pr = rho + c;
    const float4 r = __ldg(reinterpret_cast<const float4 *>(pr));
    const float4 ym = __ldg(reinterpret_cast<const float4 *>(pr - (y > 0 ? G.nx : 0)));
    const float4 yq = __ldg(reinterpret_cast<const float4 *>(pr + (y < G.ny - 1 ? G.nx : 0)));
    const float4 zm = __ldg(reinterpret_cast<const float4 *>(pr - (k > 0 ? G.plane : 0)));
    const float4 zq = __ldg(reinterpret_cast<const float4 *>(pr + (k < G.nz_global - 1 ? G.plane : 0)));
    const float xm = __ldg(pr - (xb > 0 ? 1 : 0)), xq = __ldg(pr + (xb + 4 < G.nx ? 4 : 3));
    if ((fw & SOLID4) == SOLID4) return;
    const float r0[4] = {r.x, r.y, r.z, r.w}, lo[4] = {xm, r.x, r.y, r.z}, hi[4] = {r.y, r.z, r.w, xq};
    const float ylo[4] = {ym.x, ym.y, ym.z, ym.w}, yhi[4] = {yq.x, yq.y, yq.z, yq.w};
    const float zlo[4] = {zm.x, zm.y, zm.z, zm.w}, zhi[4] = {zq.x, zq.y, zq.z, zq.w};
    const int ypos = y == 0 ? -1 : (y == G.ny - 1 ? 1 : 0), zpos = k == 0 ? -1 : (k == G.nz_global - 1 ? 1 : 0);
    float f[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = xb + j;
        const float gx = pressure_gradient_diff(r0[j], lo[j], hi[j], x == 0 ? -1 : (x == G.nx - 1 ? 1 : 0));
        const float gy = pressure_gradient_diff(r0[j], ylo[j], yhi[j], ypos);
        const float gz = pressure_gradient_diff(r0[j], zlo[j], zhi[j], zpos);
        pressure_gradient_value(r0[j], gx, gy, gz, max_force, scale, f[0][j], f[1][j], f[2][j]);
    }
    if ((fw & SOLID4) == 0u) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float4 *p = reinterpret_cast<float4 *>(bf + d * n + c);
            float4 v = make_float4(f[d][0], f[d][1], f[d][2], f[d][3]);
            if (accumulate) { const float4 o = *p; v = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w); }
            *p = v;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((fw >> (8 * j)) & LBM_FLAG_SOLID) continue;
#pragma unroll
            for (int d = 0; d < 3; ++d) bf[d * n + c + j] = accumulate ? bf[d * n + c + j] + f[d][j] : f[d][j];
        }
    }
}

// filter_paper.py:471-536.  (x-chunk, y, owned z) launch grid, VEC x-consecutive cells per thread: the coordinates come from the
// block index (the first version derived them from a linear index over the whole volume with 64-bit divisions) and the flag bytes
// of a quad are one 32-bit load -- the filter zone is a thin shell, so nearly every thread ends after that word.
__device__ __forceinline__ void forchheimer_cell(const long long n, const long long c, const float *u, float *bf, float K, float beta, float c_darcy,
                                                 float c_forch, float fmax) {
    const float ux = u[c], uy = u[n + c], uz = u[2 * n + c];
    const float umag = sqrtf(dot3(ux, uy, uz, ux, uy, uz));
    if (!(umag > 1e-8f) || !(K > 1e-12f)) return;
    const float coeff = c_darcy / K + ((c_forch * beta) * umag) / sqrtf(K);
    float rx = (-coeff) * ux, ry = (-coeff) * uy, rz = (-coeff) * uz;
    const float mag = sqrtf(dot3(rx, ry, rz, rx, ry, rz));
    if (mag > fmax) { const float s = fmax / mag; rx = rx * s; ry = ry * s; rz = rz * s; }
    bf[c] = bf[c] + rx; bf[n + c] = bf[n + c] + ry; bf[2 * n + c] = bf[2 * n + c] + rz;
}
template <int VEC>
__global__ void forchheimer_force_kernel(Grid G, const float *u, const uint8_t *flags, float *bf, float K, float beta,
                                         float c_darcy, float c_forch, float fmax) {
    const int x0 = (int)(blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (x0 >= G.nx) return;
    const int y = (int)blockIdx.y, z = (int)blockIdx.z, k = G.z0 + z;
    if (y < 1 || y > G.ny - 2 || k < 1 || k > G.nz_global - 2) return;
    const long long c0 = ((long long)(z + G.zg) * G.ny + y) * G.nx + x0;
    unsigned fw;
    if constexpr (VEC == 4) fw = *reinterpret_cast<const unsigned *>(flags + c0);
    else fw = flags[c0];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int x = x0 + j;
        const unsigned f = (fw >> (8 * j)) & 0xffu;
        if (x < 1 || x > G.nx - 2 || !(f & LBM_FLAG_FILTER) || (f & LBM_FLAG_SOLID)) continue;
        forchheimer_cell(G.vol, c0 + j, u, bf, K, beta, c_darcy, c_forch, fmax);
    }
}

// legacy/lbm_solver.py:1478-1483.  VEC = 4: one quad per trip, flags as one word, all-fluid quads as 128-bit vectors.
template <int VEC>
__global__ void add_reaction_kernel(Grid G, const float *reaction, const uint8_t *flags, float *bf) {
    const long long n = G.vol;
    if constexpr (VEC == 4) {
        for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n / 4; q += (long long)gridDim.x * blockDim.x) {
            const long long c = 4 * q;
            const unsigned fw = flags ? *reinterpret_cast<const unsigned *>(flags + c) : 0u;
            constexpr unsigned SOLID4 = 0x01010101u * LBM_FLAG_SOLID;
            if ((fw & SOLID4) == SOLID4) continue;
            if ((fw & SOLID4) == 0u) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float4 *p = reinterpret_cast<float4 *>(bf + d * n + c);
                    const float4 a = *p, r = *reinterpret_cast<const float4 *>(reaction + d * n + c);
                    *p = make_float4(a.x + r.x, a.y + r.y, a.z + r.z, a.w + r.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((fw >> (8 * j)) & LBM_FLAG_SOLID) continue;
#pragma unroll
                    for (int d = 0; d < 3; ++d) bf[d * n + c + j] = bf[d * n + c + j] + reaction[d * n + c + j];
                }
            }
        }
    } else {
        for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
            if (flags && (flags[c] & LBM_FLAG_SOLID)) continue;
#pragma unroll
            for (int d = 0; d < 3; ++d) bf[d * n + c] = bf[d * n + c] + reaction[d * n + c];
        }
    }
}

// ---- work list + neighbour masks for the step kernel ---------------------------------------------
// tile = 32*vec x-consecutive cells of `ty` consecutive rows of an owned plane (ty = 1: one warp-tile of the
// register-staged kernels; ty > 1: one CTA tile of the TMA-staged kernel); active when at least one cell is fluid.
// id = (z*rows + y/ty)*segs + seg, ascending ids follow memory order.
__global__ void tile_flags_kernel(Grid G, const uint8_t *flags, int vec, int ty, int rows, int segs, uint8_t *tile_flag) {
    const long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n = (long long)G.nz * rows * segs;
    if (id >= n) return;
    const int seg = (int)(id % segs);
    const long long row = id / segs;               // z*rows + y/ty
    const int y0 = (int)(row % rows) * ty, z = (int)(row / rows);
    const int xs = seg * 32 * vec, xe = min(G.nx, xs + 32 * vec);
    bool fluid = false;
    for (int y = y0; y < min(G.ny, y0 + ty); ++y) {
        const uint8_t *f = flags + ((long long)(z + G.zg) * G.ny + y) * G.nx;
        for (int x = xs; x < xe; ++x) fluid |= !(f[x] & LBM_FLAG_SOLID);
    }
    tile_flag[id] = fluid ? 1 : 0;
}
#ifndef LBM_EMULATE_ON_HOST   /* tests/emu compiles the plain kernels of this file with the host compiler */

__global__ void count_flagged_per_plane_kernel(const uint8_t *tile_flag, int per_plane, int nz, int *plane_count) {
    const int z = blockIdx.x;
    int n = 0;
    for (int i = threadIdx.x; i < per_plane; i += blockDim.x) n += tile_flag[(long long)z * per_plane + i] ? 1 : 0;
    typedef cub::BlockReduce<int, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const int tot = BR(tmp).Sum(n);
    if (threadIdx.x == 0 && z < nz) plane_count[z] = tot;
}
#endif

__global__ void expand_tiles_kernel(const int *ids, int n, int rows, int ty, int segs, unsigned *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids[i];
    const int seg = id % segs, row = id / segs;
    out[i] = (unsigned)seg | ((unsigned)((row % rows) * ty) << 8) | ((unsigned)(row / rows) << 20);
}

// per list entry (ty = 1): bit l = lane l of the warp must load, i.e. some cell of [x0 - 1, x0 + vec] in this row is fluid
// (its own cells, or the adjacent cell of a neighbour lane that takes a shifted population from it by shuffle)
__global__ void tile_lane_mask_kernel(Grid G, const uint8_t *flags, int vec, const unsigned *items, int n, unsigned *mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned e = items[i];
    const int seg = (int)(e & 0xffu), y = (int)((e >> 8) & 0xfffu), z = (int)(e >> 20);
    const uint8_t *row = flags + ((long long)(z + G.zg) * G.ny + y) * G.nx;
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) {
        const int xs = seg * 32 * vec + l * vec;
        if (xs >= G.nx) break;
        bool any = false;
        for (int x = max(0, xs - 1); x <= min(G.nx - 1, xs + vec); ++x) any |= !(row[x] & LBM_FLAG_SOLID);
        if (any) m |= 1u << l;
    }
    mask[i] = m;
}

// per near-wall fluid cell: bit q of the low word = the source cell x - e_q is solid (bounce-back), bit q of the high
// word = the source lies outside an open face (stale inflow w_q).  Replaces 18 neighbour-flag loads per cell per step;
// the array is dense ([vol] u64) but the step kernel reads it only where the NEAR flag is set.
__global__ void neighbour_mask_kernel(Grid G, const uint8_t *flags, unsigned long long *nbr) {
    const long long n = G.vol;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const unsigned f = flags[c];
        unsigned solid_bits = 0, oob_bits = 0;
        if ((f & LBM_FLAG_NEAR) && !(f & LBM_FLAG_SOLID)) {
            const int x = (int)(c % G.nx);
            const int y = (int)((c / G.nx) % G.ny);
            const int z = (int)(c / G.plane) - G.zg;
            for (int q = 1; q < Q; ++q) {
                int xs = x - cx(q), ys = y - cy(q), zs = z - cz(q);
                bool oob = wrap_or_oob(xs, G.nx, G.per_x) | wrap_or_oob(ys, G.ny, G.per_y);
                const int zs_g = G.z0 + zs;
                if (zs_g < 0 || zs_g >= G.nz_global) oob |= !G.per_z;
                int zsp = zs + G.zg;
                if (!G.zg) { if (zs < 0) zsp = G.nz - 1; else if (zs >= G.nz) zsp = 0; }
                if (oob) oob_bits |= 1u << q;
                else if (flags[((long long)zsp * G.ny + ys) * G.nx + xs] & LBM_FLAG_SOLID) solid_bits |= 1u << q;
            }
        }
        nbr[c] = (unsigned long long)solid_bits | ((unsigned long long)oob_bits << 32);
    }
}

// Builds the active warp-tile list (device array allocated here, owned by the caller = lbm_ctx), its per-plane
// offsets (host vector of nz+1 entries) and the neighbour masks.  Synchronises the stream: geometry changes are rare.
#ifndef LBM_EMULATE_ON_HOST   /* tests/emu compiles the plain kernels of this file with the host compiler */
cudaError_t build_work_lists(const Grid &G, const uint8_t *flags, int vec, int ty, unsigned **d_tiles, unsigned **d_tile_mask,
                             std::vector<int> &tile_off, unsigned long long **d_nbr, cudaStream_t s) {
    cudaError_t e;
    const int segs = (G.nx + 32 * vec - 1) / (32 * vec);
    if (segs > 256 || G.ny > 4096 || G.nz > 4096 || ty < 1) return cudaErrorInvalidValue;      // packing limits of a list entry
    const int rows = (G.ny + ty - 1) / ty;
    const int per_plane = rows * segs;
    const long long ntiles = (long long)per_plane * G.nz;
    uint8_t *tile_flag = nullptr; int *d_count = nullptr, *d_num = nullptr, *d_ids = nullptr; void *tmp = nullptr; size_t tmp_bytes = 0;
    if (!*d_nbr) { if ((e = cudaMalloc(d_nbr, sizeof(unsigned long long) * (size_t)G.vol)) != cudaSuccess) return e; }
    neighbour_mask_kernel<<<148 * 16, 256, 0, s>>>(G, flags, *d_nbr);
    if ((e = cudaMalloc(&tile_flag, (size_t)ntiles)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d_count, sizeof(int) * (size_t)G.nz)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d_num, sizeof(int))) != cudaSuccess) return e;
    tile_flags_kernel<<<(unsigned)((ntiles + 255) / 256), 256, 0, s>>>(G, flags, vec, ty, rows, segs, tile_flag);
    count_flagged_per_plane_kernel<<<G.nz, 256, 0, s>>>(tile_flag, per_plane, G.nz, d_count);
    std::vector<int> counts((size_t)G.nz);
    if ((e = cudaMemcpyAsync(counts.data(), d_count, sizeof(int) * counts.size(), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    tile_off.assign(G.nz + 1, 0);
    for (int z = 0; z < G.nz; ++z) tile_off[z + 1] = tile_off[z] + counts[z];
    if (*d_tiles) { cudaFree(*d_tiles); *d_tiles = nullptr; }
    if (*d_tile_mask) { cudaFree(*d_tile_mask); *d_tile_mask = nullptr; }
    const int n_t = tile_off[G.nz];
    // slabs: room for a copy of the first and the last owned plane's entries behind the list (one launch for both, lbm_api.cu)
    const int n_b = (G.zg && G.nz >= 2) ? (tile_off[1] - tile_off[0]) + (tile_off[G.nz] - tile_off[G.nz - 1]) : 0;
    if ((e = cudaMalloc(d_tiles, sizeof(unsigned) * (size_t)(n_t + n_b > 0 ? n_t + n_b : 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc(d_tile_mask, sizeof(unsigned) * (size_t)(n_t + n_b > 0 ? n_t + n_b : 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d_ids, sizeof(int) * (size_t)(n_t > 0 ? n_t : 1))) != cudaSuccess) return e;
    thrust::counting_iterator<int> idx(0);
    cub::DeviceSelect::Flagged(nullptr, tmp_bytes, idx, tile_flag, d_ids, d_num, (int)ntiles, s);
    if ((e = cudaMalloc(&tmp, tmp_bytes)) != cudaSuccess) return e;
    if (n_t > 0) {
        cub::DeviceSelect::Flagged(tmp, tmp_bytes, idx, tile_flag, d_ids, d_num, (int)ntiles, s);
        expand_tiles_kernel<<<(n_t + 255) / 256, 256, 0, s>>>(d_ids, n_t, rows, ty, segs, *d_tiles);
        if (ty == 1) tile_lane_mask_kernel<<<(n_t + 127) / 128, 128, 0, s>>>(G, flags, vec, *d_tiles, n_t, *d_tile_mask);
        if (n_b > 0) {
            const int n0 = tile_off[1] - tile_off[0], n1 = tile_off[G.nz] - tile_off[G.nz - 1];
            cudaMemcpyAsync(*d_tiles + n_t, *d_tiles + tile_off[0], sizeof(unsigned) * (size_t)n0, cudaMemcpyDeviceToDevice, s);
            cudaMemcpyAsync(*d_tiles + n_t + n0, *d_tiles + tile_off[G.nz - 1], sizeof(unsigned) * (size_t)n1, cudaMemcpyDeviceToDevice, s);
            cudaMemcpyAsync(*d_tile_mask + n_t, *d_tile_mask + tile_off[0], sizeof(unsigned) * (size_t)n0, cudaMemcpyDeviceToDevice, s);
            cudaMemcpyAsync(*d_tile_mask + n_t + n0, *d_tile_mask + tile_off[G.nz - 1], sizeof(unsigned) * (size_t)n1, cudaMemcpyDeviceToDevice, s);
        }
    }
    e = cudaStreamSynchronize(s);
    cudaFree(tmp); cudaFree(tile_flag); cudaFree(d_count); cudaFree(d_num); cudaFree(d_ids);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}
#endif

// ---- PressureGradientDrive.apply_density_drive (pressure_gradient_drive.py:95-122, "method A") ---------------------------
// rho of every fluid cell moves towards the target profile of its z plane by rate * (target - rho), at most max_adjust per
// call, and is clamped to [rho_min, rho_max].  f32, the reference's statement order.  Four x-consecutive cells per thread on
// 128-bit loads / stores when rows are 16-byte multiples (VEC = 4), else one.  target_z: one value per plane, ghost planes included.
template <int VEC>
__global__ void density_drive_kernel(Grid G, float *rho, const uint8_t *flags, const float *target_z, float rate, float max_adjust,
                                     float rho_min, float rho_max) {
    const long long n = (long long)G.nz * G.plane / VEC;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long c = (long long)G.zg * G.plane + i * VEC;
        const float target = target_z[c / G.plane];
        float r[VEC]; uint8_t fl[VEC];
        if (VEC == 4) {
            const float4 v = *reinterpret_cast<const float4 *>(rho + c);
            const uchar4 f = *reinterpret_cast<const uchar4 *>(flags + c);
            r[0] = v.x; r[VEC > 1 ? 1 : 0] = v.y; r[VEC > 2 ? 2 : 0] = v.z; r[VEC > 3 ? 3 : 0] = v.w;
            fl[0] = f.x; fl[VEC > 1 ? 1 : 0] = f.y; fl[VEC > 2 ? 2 : 0] = f.z; fl[VEC > 3 ? 3 : 0] = f.w;
        } else { r[0] = rho[c]; fl[0] = flags[c]; }
        bool any = false;
        for (int k = 0; k < VEC; ++k) {
            if (fl[k] & LBM_FLAG_SOLID) continue;
            const float diff = target - r[k];
            float adj = diff * rate;
            if (fabsf(adj) > max_adjust) adj = adj > 0.0f ? max_adjust : -max_adjust;
            r[k] = fmaxf(rho_min, fminf(rho_max, r[k] + adj));
            any = true;
        }
        if (!any) continue;
        if (VEC == 4) *reinterpret_cast<float4 *>(rho + c) = make_float4(r[0], r[VEC > 1 ? 1 : 0], r[VEC > 2 ? 2 : 0], r[VEC > 3 ? 3 : 0]);
        else rho[c] = r[0];
    }
}

// ---- packed quad list and wall links of the four-cell walls kernel (lbm_phys_chord.cuh) --------------------------
// A "quad" is 4 x-consecutive cells on a 16-byte boundary.  It is active when it or the other quad of its 32-byte sector holds a
// fluid cell: the kernel stores whole quads, so both halves of every sector it touches are written and no sector reaches DRAM half
// filled (a partly written sector costs a 32-byte fill read; at the chord ends of the V60 mask that was 0.37 GB per step).  The
// kernel's work list is the sequence of ALL active quads of a plane in memory order (y, then x), cut into tiles of 32 -- one warp
// per tile, one lane per quad, tiles run across row ends.  A first version cut tiles per chord (<= 32 quads of ONE row): on the
// V60 512^3 mask that launched 14.77 M lane slots for 12.05 M active quads (82 % of the lanes alive), and since a warp costs the same
// whether 11 or 32 of its lanes work, the step ran at 0.67 of the HBM peak where a periodic box (every lane alive) runs at
// 0.86 (profiles/r02_exp_structure_cost_periodic_box.log).  Only the last tile of a plane is padded (slab launches address plane
// ranges).  Per lane slot (u64): bits 0-11 quad index in the row, 12-27 y, 28-43 z, 44 live, 45 / 46 the previous / next lane
// of the SAME tile holds the quad to the left / right in the same row (else the lane fetches that neighbour itself), 47-54 one bit
// per neighbouring row (dz, dy) != (0, 0), bit (dz + 1) * 3 + (dy + 1) (minus one behind the centre): the quad at the same x in that
// row holds no fluid cell or lies outside an open face, so nothing the collision uses comes from it (a fluid cell that pulls from a
// solid cell takes the link value, a source outside the box is w_q) and the kernel does not read it from DRAM.
// Wall link (u32) = one (fluid cell, direction q) pair whose target x + e_q is solid.  Halfway bounce-back hands the cell's
// post-collision f_q back to the same cell as f_opp(q) one step later, so the value never has to visit the population arrays: it
// waits in a per-link buffer (`wall`, one float per link, read and written by the tile that owns the link, coalesced).  The link
// names two WORDS OF THE TILE'S STAGE (lbm_phys_chord.cuh: row r = 160 words: 4 per lane, then one edge word per lane):
//   bits 0-11  where the waiting value goes before the collision: row opp(q), the word the cell pulls opp(q) from (its solid
//              neighbour's place: word 4 lane + cell + cx(q), or the lane's edge word when that place lies outside the lane's quad
//              and the neighbouring lane does not hold the neighbouring quad);
//   bits 12-23 where the new value is taken from after the collision: row q, word 4 lane + cell.
// Per tile (uint2): first link, number of links.  A first version stored the value into the solid neighbour's slot of the population
// array instead (4-byte scattered stores, 8-byte links with the target index) and the fluid cells of chord-end quads one by one:
// timing the kernel without those two loops showed 0.14 of 1.75 ms in them (profiles/r02_exp_copy_nolinks.log).
// The kernels below are plain (one thread per row / per tile) so that tests/emu can run them on the CPU; they run once per
// geometry change.
#define LBM_QUAD_LIVE (1ull << 44)
#define LBM_QUAD_LEFT (1ull << 45)
#define LBM_QUAD_RIGHT (1ull << 46)
#define LBM_STAGE_ROW_WORDS 160
__device__ __forceinline__ bool quad_has_fluid(const uint8_t *row, int q) {
    return !((row[4 * q] & row[4 * q + 1] & row[4 * q + 2] & row[4 * q + 3]) & LBM_FLAG_SOLID);
}
__device__ __forceinline__ bool quad_active(const uint8_t *row, int q, int nquads) {
    return quad_has_fluid(row, q) || ((q ^ 1) < nquads && quad_has_fluid(row, q ^ 1));
}
__global__ void quad_count_kernel(Grid G, const uint8_t *flags, int *row_quads) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= G.nz * G.ny) return;
    const int z = r / G.ny, y = r - z * G.ny;
    const uint8_t *row = flags + ((long long)(z + G.zg) * G.ny + y) * G.nx;
    int n = 0;
    for (int q = 0; q < G.nx / 4; ++q) n += quad_active(row, q, G.nx / 4) ? 1 : 0;
    row_quads[r] = n;
}
// row_off: exclusive prefix of row_quads over all rows; plane_base[z]: first lane slot of plane z (a multiple of 32)
__global__ void quad_fill_kernel(Grid G, const uint8_t *flags, const int *row_off, const int *plane_base, unsigned long long *quads) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= G.nz * G.ny) return;
    const int z = r / G.ny, y = r - z * G.ny;
    const uint8_t *row = flags + ((long long)(z + G.zg) * G.ny + y) * G.nx;
    int slot = plane_base[z] + (row_off[r] - row_off[z * G.ny]);
    int prev = -2;
    for (int q = 0; q < G.nx / 4; ++q) {
        if (!quad_active(row, q, G.nx / 4)) continue;
        unsigned long long e = (unsigned long long)q | ((unsigned long long)y << 12) | ((unsigned long long)z << 28) | LBM_QUAD_LIVE;
        if (prev == q - 1 && (slot & 31) != 0) e |= LBM_QUAD_LEFT;
        if (q + 1 < G.nx / 4 && quad_active(row, q + 1, G.nx / 4) && ((slot + 1) & 31) != 0) e |= LBM_QUAD_RIGHT;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy) {
                if (dz == 0 && dy == 0) continue;
                int bit = (dz + 1) * 3 + (dy + 1); if (bit > 4) --bit;
                int ys = y + dy, zs = z + dz;
                bool dead = false;
                if (ys < 0 || ys >= G.ny) { if (G.per_y) ys = (ys + G.ny) % G.ny; else dead = true; }
                const int zglob = G.z0 + zs;
                if (zglob < 0 || zglob >= G.nz_global) { if (!G.per_z) dead = true; }
                if (!G.zg) { if (zs < 0) zs = G.nz - 1; else if (zs >= G.nz) zs = 0; }
                if (!dead) dead = !quad_has_fluid(flags + ((long long)(zs + G.zg) * G.ny + ys) * G.nx, q);
                if (dead) e |= 1ull << (47 + bit);
            }
        quads[slot++] = e;
        prev = q;
    }
}
// one thread per tile: number of wall links (fill == 0) or the links themselves
__global__ void quad_links_kernel(Grid G, const uint8_t *flags, const unsigned long long *nbr, const unsigned long long *quads, int n_tiles,
                                  int fill, int *tile_count, const int *link_off, uint2 *tile_links, unsigned *links) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    unsigned n = 0;
    const unsigned begin = fill ? (unsigned)link_off[t] : 0u;
    for (int l = 0; l < 32; ++l) {
        const unsigned long long e = quads[(long long)t * 32 + l];
        if (!(e & LBM_QUAD_LIVE)) continue;
        const int q0 = (int)(e & 0xfffu), y = (int)((e >> 12) & 0xffffu), z = (int)((e >> 28) & 0xffffu);
        const long long base = ((long long)(z + G.zg) * G.ny + y) * G.nx;
        for (int c = 0; c < 4; ++c) {
            const int x = 4 * q0 + c;
            const unsigned fl = flags[base + x];
            if ((fl & LBM_FLAG_SOLID) || !(fl & LBM_FLAG_NEAR)) continue;
            const unsigned m = (unsigned)nbr[base + x];
            for (int q = 1; q < Q; ++q) {
                if (!((m >> opp(q)) & 1u)) continue;
                if (fill) {
                    const int w = c + cx(q);                  // the solid neighbour's place in the stage row of opp(q)
                    unsigned pos = (unsigned)(4 * l + w);
                    if (w < 0 && !(e & LBM_QUAD_LEFT)) pos = 128u + (unsigned)l;
                    if (w > 3 && !(e & LBM_QUAD_RIGHT)) pos = 128u + (unsigned)l;
                    links[begin + n] = ((unsigned)opp(q) * LBM_STAGE_ROW_WORDS + pos) | (((unsigned)q * LBM_STAGE_ROW_WORDS + (unsigned)(4 * l + c)) << 12);
                }
                ++n;
            }
        }
    }
    if (fill) tile_links[t] = make_uint2(begin, n);
    else tile_count[t] = (int)n;
}
// wall[link] = the owning cell's own g[q] (g holds post-collision values): what the step kernel would have left there.  Runs after
// the populations were written from outside the step kernel (initialisation, lbm_import_f, a halo refresh, another kernel variant).
__global__ void wall_values_kernel(Grid G, const float *g, const unsigned long long *quads, const uint2 *tile_links, const unsigned *links,
                                   float *wall, int n_tiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint2 tl = tile_links[t];
    for (unsigned i = 0; i < tl.y; ++i) {
        const unsigned src = (links[tl.x + i] >> 12) & 0xfffu;
        const int q = (int)(src / LBM_STAGE_ROW_WORDS), w = (int)(src % LBM_STAGE_ROW_WORDS);
        const unsigned long long e = quads[(long long)t * 32 + (w >> 2)];
        const int x = 4 * (int)(e & 0xfffu) + (w & 3), y = (int)((e >> 12) & 0xffffu), z = (int)((e >> 28) & 0xffffu);
        wall[tl.x + i] = g[(long long)q * G.vol + ((long long)(z + G.zg) * G.ny + y) * G.nx + x];
    }
}
#ifndef LBM_EMULATE_ON_HOST
cudaError_t launch_wall_values(const Grid &G, const float *g, const unsigned long long *quads, const uint2 *tile_links, const unsigned *links,
                               float *wall, int n_tiles, cudaStream_t s) {
    if (n_tiles > 0) wall_values_kernel<<<(n_tiles + 127) / 128, 128, 0, s>>>(G, g, quads, tile_links, links, wall, n_tiles);
    return cudaGetLastError();
}
// Builds the packed quad list, its per-plane TILE offsets (host vector of nz + 1 entries), the wall links and the neighbour
// masks.  Synchronises the stream: geometry changes are rare.
cudaError_t build_chord_lists(const Grid &G, const uint8_t *flags, unsigned long long **d_quads, uint2 **d_tile_links, unsigned **d_links, float **d_wall,
                              std::vector<int> &tile_off, unsigned long long **d_nbr, long long *n_links_out, cudaStream_t s) {
    cudaError_t e;
    if (G.nx % 4 != 0 || G.nx > 16384 || G.ny > 65535 || G.nz + 2 * G.zg > 65535 || G.vol >= (1ll << 32)) return cudaErrorInvalidValue;   // packing limits
    const int rows = G.nz * G.ny;
    int *row_cnt = nullptr, *row_off = nullptr, *d_plane_base = nullptr, *tile_cnt = nullptr, *link_off = nullptr; void *tmp = nullptr; size_t tmp_bytes = 0;
    if (!*d_nbr) { if ((e = cudaMalloc(d_nbr, sizeof(unsigned long long) * (size_t)G.vol)) != cudaSuccess) return e; }
    neighbour_mask_kernel<<<148 * 16, 256, 0, s>>>(G, flags, *d_nbr);
    if ((e = cudaMalloc(&row_cnt, sizeof(int) * (size_t)(rows + 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&row_off, sizeof(int) * (size_t)(rows + 1))) != cudaSuccess) return e;
    cudaMemsetAsync(row_cnt + rows, 0, sizeof(int), s);
    quad_count_kernel<<<(rows + 127) / 128, 128, 0, s>>>(G, flags, row_cnt);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, row_cnt, row_off, rows + 1, s);
    if ((e = cudaMalloc(&tmp, tmp_bytes)) != cudaSuccess) return e;
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, row_cnt, row_off, rows + 1, s);
    std::vector<int> plane_start((size_t)G.nz + 1), plane_base((size_t)G.nz + 1, 0);
    if ((e = cudaMemcpy2DAsync(plane_start.data(), sizeof(int), row_off, sizeof(int) * (size_t)G.ny, sizeof(int), (size_t)G.nz + 1, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    cudaFree(tmp); tmp = nullptr;
    tile_off.assign(G.nz + 1, 0);
    for (int z = 0; z < G.nz; ++z) {
        const int nq = plane_start[z + 1] - plane_start[z];
        tile_off[z + 1] = tile_off[z] + (nq + 31) / 32;            // the last tile of a plane is padded with dead lanes
        plane_base[z + 1] = tile_off[z + 1] * 32;
    }
    const int n_t = tile_off[G.nz];
    // slabs: room for a copy of the first and the last owned plane's tiles behind the list (one launch for both, lbm_api.cu)
    const int n_b = (G.zg && G.nz >= 2) ? (tile_off[1] - tile_off[0]) + (tile_off[G.nz] - tile_off[G.nz - 1]) : 0;
    if (*d_quads) { cudaFree(*d_quads); *d_quads = nullptr; }
    if (*d_tile_links) { cudaFree(*d_tile_links); *d_tile_links = nullptr; }
    if (*d_links) { cudaFree(*d_links); *d_links = nullptr; }
    if (*d_wall) { cudaFree(*d_wall); *d_wall = nullptr; }
    const size_t nt_alloc = (size_t)(n_t + n_b > 0 ? n_t + n_b : 1);
    if ((e = cudaMalloc(d_quads, sizeof(unsigned long long) * 32 * nt_alloc)) != cudaSuccess) return e;
    if ((e = cudaMalloc(d_tile_links, sizeof(uint2) * nt_alloc)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d_plane_base, sizeof(int) * (size_t)(G.nz + 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&tile_cnt, sizeof(int) * (size_t)(n_t + 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&link_off, sizeof(int) * (size_t)(n_t + 1))) != cudaSuccess) return e;
    cudaMemsetAsync(*d_quads, 0, sizeof(unsigned long long) * 32 * nt_alloc, s);      // dead lanes: cell (0, 0, 0), never stored
    cudaMemsetAsync(*d_tile_links, 0, sizeof(uint2) * nt_alloc, s);
    cudaMemsetAsync(tile_cnt + n_t, 0, sizeof(int), s);
    cudaMemcpyAsync(d_plane_base, plane_base.data(), sizeof(int) * (size_t)(G.nz + 1), cudaMemcpyHostToDevice, s);
    int n_l = 0;
    if (n_t > 0) {
        quad_fill_kernel<<<(rows + 127) / 128, 128, 0, s>>>(G, flags, row_off, d_plane_base, *d_quads);
        quad_links_kernel<<<(n_t + 127) / 128, 128, 0, s>>>(G, flags, *d_nbr, *d_quads, n_t, 0, tile_cnt, nullptr, nullptr, nullptr);
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tile_cnt, link_off, n_t + 1, s);
        if ((e = cudaMalloc(&tmp, tmp_bytes)) != cudaSuccess) return e;
        cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, tile_cnt, link_off, n_t + 1, s);
        if ((e = cudaMemcpyAsync(&n_l, link_off + n_t, sizeof(int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    }
    if ((e = cudaMalloc(d_links, sizeof(unsigned) * (size_t)(n_l > 0 ? n_l : 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc(d_wall, sizeof(float) * (size_t)(n_l > 0 ? n_l : 1))) != cudaSuccess) return e;
    if (n_t > 0) quad_links_kernel<<<(n_t + 127) / 128, 128, 0, s>>>(G, flags, *d_nbr, *d_quads, n_t, 1, nullptr, link_off, *d_tile_links, *d_links);
    if (n_b > 0) {      // the finished entries of the two boundary planes, once more, contiguous
        const int n0 = tile_off[1] - tile_off[0], n1 = tile_off[G.nz] - tile_off[G.nz - 1];
        cudaMemcpyAsync(*d_quads + (size_t)n_t * 32, *d_quads + (size_t)tile_off[0] * 32, sizeof(unsigned long long) * 32 * (size_t)n0, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(*d_quads + (size_t)(n_t + n0) * 32, *d_quads + (size_t)tile_off[G.nz - 1] * 32, sizeof(unsigned long long) * 32 * (size_t)n1, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(*d_tile_links + n_t, *d_tile_links + tile_off[0], sizeof(uint2) * (size_t)n0, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(*d_tile_links + n_t + n0, *d_tile_links + tile_off[G.nz - 1], sizeof(uint2) * (size_t)n1, cudaMemcpyDeviceToDevice, s);
    }
    e = cudaStreamSynchronize(s);
    if (n_links_out) *n_links_out = n_l;
    cudaFree(tmp); cudaFree(row_cnt); cudaFree(row_off); cudaFree(d_plane_base); cudaFree(tile_cnt); cudaFree(link_off);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}
#endif

// ---- write-side bounce-back slots (compat = physical, walls path; see lbm_phys.cuh) ---------------------
// Re-creates the bounce-back slots of a population buffer from the cells' own values:
//   g[opp q][x + e_q] = g[q][x]   for every fluid cell x of planes [z_begin, z_end) whose neighbour x + e_q is solid.
// Needed once after the populations were written from outside the step kernel (initialisation, lbm_import_f, a
// geometry change) and, in slab mode, for the two boundary planes after each halo exchange (the incoming ghost
// planes overwrite the slots that live in them).
__global__ void bounce_slots_kernel(Grid G, float *g, const uint8_t *flags, const unsigned long long *nbr, int z_begin, int z_end) {
    const long long per = G.plane;
    const long long n = per * (z_end - z_begin);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = z_begin + (int)(i / per);
        const int rem = (int)(i % per);
        const int y = rem / G.nx, x = rem % G.nx;
        const int zp = z + G.zg;
        const long long c = ((long long)zp * G.ny + y) * G.nx + x;
        const unsigned fl = flags[c];
        if ((fl & LBM_FLAG_SOLID) || !(fl & LBM_FLAG_NEAR)) continue;
        const unsigned solid_src = (unsigned)nbr[c];
        if (!solid_src) continue;
        for (int q = 1; q < Q; ++q) {
            if (!(solid_src & (1u << opp(q)))) continue;
            int xt = x + cx(q), yt = y + cy(q), zt = zp + cz(q);
            if (xt < 0) xt = G.nx - 1; else if (xt >= G.nx) xt = 0;
            if (yt < 0) yt = G.ny - 1; else if (yt >= G.ny) yt = 0;
            if (!G.zg) { if (zt < 0) zt = G.nz - 1; else if (zt >= G.nz) zt = 0; }
            const long long t = ((long long)zt * G.ny + yt) * G.nx + xt;
            g[(long long)opp(q) * G.vol + t] = g[(long long)q * G.vol + c];
        }
    }
}


#ifndef LBM_EMULATE_ON_HOST   /* tests/emu compiles the plain kernels of this file with the host compiler */
// ---- self-test of the packed reciprocal / square root (lbm_phys.cuh) -----------------------------------
// Every f32 bit pattern goes through both lanes of Ops<P2>::rcp / ::sqrt and is compared, bit for bit, with the
// correctly rounded scalar intrinsics.  counts[0] = reciprocal mismatches, counts[1] = square-root mismatches.
__global__ void selftest_math_kernel(unsigned long long *counts) {
    unsigned long long bad_r = 0, bad_s = 0;
    const unsigned long long n = 1ull << 32;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned a = (unsigned)i, b = (unsigned)i * 2654435761u + 12345u;
        const float x0 = __uint_as_float(a), x1 = __uint_as_float(b);
        const P2 v = p2_make(x0, x1);
        const P2 r = Ops<P2>::rcp(v), s = Ops<P2>::sqrt(v);
        auto same = [](float p, float q) { return __float_as_uint(p) == __float_as_uint(q) || (p != p && q != q); };
        if (!same(p2_lo(r), __frcp_rn(x0)) || !same(p2_hi(r), __frcp_rn(x1))) ++bad_r;
        if (!same(p2_lo(s), __fsqrt_rn(x0)) || !same(p2_hi(s), __fsqrt_rn(x1))) ++bad_s;
    }
    if (bad_r) atomicAdd(counts, bad_r);
    if (bad_s) atomicAdd(counts + 1, bad_s);
}

// packed add / sub / mul / fma (all operand forms the collision uses: registers, broadcast scalars, literals) and the
// whole collision operator, packed against scalar, on pseudo-random near-equilibrium states.
// counts[2..5] = add, sub, mul, fma mismatches; counts[6] = cells whose packed collision differs from the scalar one.
__device__ __forceinline__ unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s; }
__device__ __forceinline__ float unit(unsigned &s) { return (float)(lcg(s) >> 8) * (1.0f / 16777216.0f); }
__global__ void selftest_packed_kernel(unsigned long long *counts, StepArgs P, int iters) {
    using O = Ops<P2>;
    unsigned seed = (blockIdx.x * blockDim.x + threadIdx.x) * 747796405u + 2891336453u;
    unsigned long long bad[5] = {0, 0, 0, 0, 0};
    auto same = [](float p, float q) { return __float_as_uint(p) == __float_as_uint(q) || (p != p && q != q); };
    for (int it = 0; it < iters; ++it) {
        const float a0 = unit(seed) - 0.5f, a1 = (unit(seed) - 0.5f) * 1e-3f, b0 = unit(seed) * 3.0f - 1.0f, b1 = unit(seed) - 0.3f;
        const float c0 = -a0 * b0 + (unit(seed) - 0.5f) * 1e-7f, c1 = unit(seed);
        const P2 a = p2_make(a0, a1), b = p2_make(b0, b1), c = p2_make(c0, c1);
        P2 r = O::add(a, b); if (!same(p2_lo(r), __fadd_rn(a0, b0)) || !same(p2_hi(r), __fadd_rn(a1, b1))) ++bad[0];
        r = O::sub(a, b); if (!same(p2_lo(r), __fsub_rn(a0, b0)) || !same(p2_hi(r), __fsub_rn(a1, b1))) ++bad[1];
        r = O::mul(a, b); if (!same(p2_lo(r), __fmul_rn(a0, b0)) || !same(p2_hi(r), __fmul_rn(a1, b1))) ++bad[2];
        r = O::mul(O::bc(4.5f), b); if (!same(p2_lo(r), __fmul_rn(4.5f, b0)) || !same(p2_hi(r), __fmul_rn(4.5f, b1))) ++bad[2];
        r = O::fma(a, b, c); if (!same(p2_lo(r), __fmaf_rn(a0, b0, c0)) || !same(p2_hi(r), __fmaf_rn(a1, b1, c1))) ++bad[3];
        r = O::fma(O::bc(-1.5f), b, O::bc(1.0f)); if (!same(p2_lo(r), __fmaf_rn(-1.5f, b0, 1.0f)) || !same(p2_hi(r), __fmaf_rn(-1.5f, b1, 1.0f))) ++bad[3];
        r = O::fma(O::bc(P.tau_water), b, c); if (!same(p2_lo(r), __fmaf_rn(P.tau_water, b0, c0)) || !same(p2_hi(r), __fmaf_rn(P.tau_water, b1, c1))) ++bad[3];
        // collision: two random near-equilibrium cells
        float fs[2][Q]; P2 fp[Q];
        CellIn<P2> ip; float F[2][3], ph[2];
        for (int l = 0; l < 2; ++l) {
            const float rho = 0.9f + 0.2f * unit(seed);
            const float ux = 0.1f * (unit(seed) - 0.5f), uy = 0.1f * (unit(seed) - 0.5f), uz = 0.1f * (unit(seed) - 0.5f);
            for (int q = 0; q < Q; ++q) {
                const float eu = cx(q) * ux + cy(q) * uy + cz(q) * uz;
                fs[l][q] = wq(q) * rho * (1.0f + 3.0f * eu + 4.5f * eu * eu - 1.5f * (ux * ux + uy * uy + uz * uz)) * (1.0f + 1e-3f * (unit(seed) - 0.5f));
            }
            for (int d = 0; d < 3; ++d) F[l][d] = 1e-4f * (unit(seed) - 0.5f);
            ph[l] = unit(seed);
            ip.flag[l] = (lcg(seed) >> 16) & (LBM_FLAG_FILTER | LBM_FLAG_LES);
        }
        for (int q = 0; q < Q; ++q) fp[q] = p2_make(fs[0][q], fs[1][q]);
        ip.Fx = p2_make(F[0][0], F[1][0]); ip.Fy = p2_make(F[0][1], F[1][1]); ip.Fz = p2_make(F[0][2], F[1][2]);
        ip.phase = p2_make(ph[0], ph[1]);
        CellMacro<P2> mp;
        collide_phys<P2, true, true, true, true>(fp, ip, mp, P, true, true);
        for (int l = 0; l < 2; ++l) {
            CellIn<float> is; is.Fx = F[l][0]; is.Fy = F[l][1]; is.Fz = F[l][2]; is.phase = ph[l]; is.flag[0] = ip.flag[l];
            CellMacro<float> ms;
            collide_phys<float, true, true, true, true>(fs[l], is, ms, P, true, true);
            bool ok = same(O::get(mp.rho, l), ms.rho) && same(O::get(mp.ux, l), ms.ux) && same(O::get(mp.uy, l), ms.uy) && same(O::get(mp.uz, l), ms.uz);
            for (int q = 0; q < Q; ++q) ok = ok && same(O::get(fp[q], l), fs[l][q]);
            if (!ok) ++bad[4];
#ifdef LBM_SELFTEST_DEBUG
            if (!ok && blockIdx.x == 0 && threadIdx.x == 0 && it < 2) {
                printf("it %d lane %d flag %u: rho %08x %08x ux %08x %08x uy %08x %08x uz %08x %08x\n", it, l, ip.flag[l],
                       __float_as_uint(O::get(mp.rho, l)), __float_as_uint(ms.rho), __float_as_uint(O::get(mp.ux, l)), __float_as_uint(ms.ux),
                       __float_as_uint(O::get(mp.uy, l)), __float_as_uint(ms.uy), __float_as_uint(O::get(mp.uz, l)), __float_as_uint(ms.uz));
                for (int q = 0; q < Q; ++q)
                    printf("   q %d packed %08x scalar %08x (%.9g %.9g)\n", q, __float_as_uint(O::get(fp[q], l)), __float_as_uint(fs[l][q]), O::get(fp[q], l), fs[l][q]);
            }
#endif
        }
    }
    for (int k = 0; k < 5; ++k) if (bad[k]) atomicAdd(counts + 2 + k, bad[k]);
}
cudaError_t run_selftest_math(unsigned long long out[7], const StepArgs &args, cudaStream_t s) {
    unsigned long long *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 7 * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(d, 0, 7 * sizeof(unsigned long long), s);
    selftest_math_kernel<<<148 * 16, 256, 0, s>>>(d);
    selftest_packed_kernel<<<148 * 4, 128, 0, s>>>(d, args, 64);
    e = cudaMemcpyAsync(out, d, 7 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ---- fused field statistics -------------------------------------------------------------------------
// One pass over rho and u of the owned fluid cells replaces the reference's statistics / stability scans:
// visualizer.compute_statistics (visualizer.py:130-183: max / mean speed, masses), NumericalStabilityMonitor.
// check_field_stability (numerical_stability.py:52-110: max |u|, min / max rho, NaN and Inf counts) and the
// .to_numpy() reductions of main.py:907-912 and lbm_diagnostics.py.  Deterministic: per-block partials in a fixed
// order, then one block folds them in index order (no float atomics).
//   out[0] max |u| (finite values)   out[1] min rho   out[2] max rho   out[3] sum rho (mass)
//   out[4] sum 0.5 rho |u|^2         out[5] NaN count (rho, |u|)       out[6] Inf count     out[7] fluid cells
struct StatPartial { double v[8]; };
__device__ __forceinline__ void stat_merge(StatPartial &a, const StatPartial &b) {
    a.v[0] = fmax(a.v[0], b.v[0]); a.v[1] = fmin(a.v[1], b.v[1]); a.v[2] = fmax(a.v[2], b.v[2]);
    a.v[3] += b.v[3]; a.v[4] += b.v[4]; a.v[5] += b.v[5]; a.v[6] += b.v[6]; a.v[7] += b.v[7];
}
__device__ __forceinline__ StatPartial stat_identity() {
    StatPartial s; s.v[0] = 0.0; s.v[1] = 1e300; s.v[2] = -1e300; s.v[3] = s.v[4] = s.v[5] = s.v[6] = s.v[7] = 0.0; return s;
}
__device__ __forceinline__ StatPartial stat_block_reduce(StatPartial s) {
    __shared__ StatPartial sh[32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        StatPartial o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = __shfl_down_sync(0xffffffffu, s.v[k], off);
        stat_merge(s, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = s;
    __syncthreads();
    if (warp == 0) {
        s = lane < (int)(blockDim.x >> 5) ? sh[lane] : stat_identity();
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            StatPartial o;
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = __shfl_down_sync(0xffffffffu, s.v[k], off);
            stat_merge(s, o);
        }
    }
    return s;
}
// Per-thread accumulator: the maxima / minima and the counters stay in 32-bit registers (a float max of floats is exact, the counters
// of one thread stay far below 2^32); only the two sums run in f64.  With all eight values in f64 the loop body was ~150 instructions
// per cell (SASS: DSETP + SEL pairs for every fmin / fmax, F2F conversions, 64-bit register moves) and the kernel was issue-bound.
struct StatLocal {
    float umax = 0.0f, rmin = INFINITY, rmax = -INFINITY;      // +-inf = "no finite value seen" (a non-finite rho never enters)
    unsigned nan = 0u, inf = 0u, cells = 0u;
    double mass = 0.0, ke = 0.0;
};
__device__ __forceinline__ void stat_cell(StatLocal &s, const float r, const float ux, const float uy, const float uz) {
    const float um = sqrtf(dot3(ux, uy, uz, ux, uy, uz));
    s.cells += 1u;
    const bool r_nan = r != r, r_inf = isinf(r), u_nan = um != um, u_inf = isinf(um);
    s.nan += (r_nan ? 1u : 0u) + (u_nan ? 1u : 0u);
    s.inf += (r_inf ? 1u : 0u) + (u_inf ? 1u : 0u);
    const bool r_ok = !r_nan && !r_inf;
    if (r_ok) { s.rmin = fminf(s.rmin, r); s.rmax = fmaxf(s.rmax, r); s.mass += (double)r; }
    if (!u_nan && !u_inf) {
        s.umax = fmaxf(s.umax, um);
        if (r_ok) s.ke += 0.5 * (double)r * ((double)ux * ux + (double)uy * uy + (double)uz * uz);
    }
}
__device__ __forceinline__ StatPartial stat_widen(const StatLocal &l) {
    StatPartial s;
    s.v[0] = (double)l.umax; s.v[1] = l.rmin == INFINITY ? 1e300 : (double)l.rmin; s.v[2] = l.rmax == -INFINITY ? -1e300 : (double)l.rmax;
    s.v[3] = l.mass; s.v[4] = l.ke; s.v[5] = (double)l.nan; s.v[6] = (double)l.inf; s.v[7] = (double)l.cells;
    return s;
}
// VEC = 4 (nx % 4 == 0, 16-byte aligned fields): one thread per quad of x-consecutive cells -- the four flag bytes as ONE 32-bit
// load, and only quads holding a fluid cell fetch rho and u (4 x 128 bit).  Four quads per loop trip, their flag words loaded
// first: the solid 65 % of a V60 box costs one byte per cell and the data loads of up to four quads are in flight together.
// The scalar form (ragged nx) read one flag BYTE per thread and then, dependent on it, four 4-byte words: 0.74 ms at V60 512^3
// (ncu launch list of the bench command, profiles/r02_final_launches_*), five times the 0.9 GB it has to move.
template <int VEC>
__global__ void __launch_bounds__(256) field_statistics_kernel(Grid G, const float *rho, const float *u, const uint8_t *flags, StatPartial *partials) {
    StatLocal s;
    const long long per = G.plane, n = per * G.nz, off = per * G.zg;
    if constexpr (VEC == 4) {
        constexpr int TRIP = 4;
        const long long nq = n >> 2, stride = (long long)gridDim.x * blockDim.x;
        const unsigned *fw4 = reinterpret_cast<const unsigned *>(flags ? flags + off : nullptr);
        for (long long q0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; q0 < nq; q0 += stride * TRIP) {
            unsigned fw[TRIP];
#pragma unroll
            for (int t = 0; t < TRIP; ++t) {
                const long long q = q0 + t * stride;
                fw[t] = q < nq ? (fw4 ? __ldg(fw4 + q) : 0u) : 0x01010101u;          // past the end: all solid
            }
#pragma unroll
            for (int t = 0; t < TRIP; ++t) {
                if ((fw[t] & 0x01010101u) == 0x01010101u) continue;                   // LBM_FLAG_SOLID in all four bytes
                const long long c = off + 4 * (q0 + t * stride);
                const float4 r = __ldcs(reinterpret_cast<const float4 *>(rho + c));
                const float4 x = __ldcs(reinterpret_cast<const float4 *>(u + c));
                const float4 y = __ldcs(reinterpret_cast<const float4 *>(u + G.vol + c));
                const float4 z = __ldcs(reinterpret_cast<const float4 *>(u + 2 * G.vol + c));
                if (!(fw[t] & 0x00000001u)) stat_cell(s, r.x, x.x, y.x, z.x);
                if (!(fw[t] & 0x00000100u)) stat_cell(s, r.y, x.y, y.y, z.y);
                if (!(fw[t] & 0x00010000u)) stat_cell(s, r.z, x.z, y.z, z.z);
                if (!(fw[t] & 0x01000000u)) stat_cell(s, r.w, x.w, y.w, z.w);
            }
        }
    } else {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            const long long c = off + i;
            if (flags && (flags[c] & LBM_FLAG_SOLID)) continue;
            stat_cell(s, rho[c], u[c], u[G.vol + c], u[2 * G.vol + c]);
        }
    }
    const StatPartial b = stat_block_reduce(stat_widen(s));
    if (threadIdx.x == 0) partials[blockIdx.x] = b;
}
// Over the packed quad list of the four-cell walls kernel (every quad whose 32-byte sector holds a fluid cell, in memory order; owned
// planes only): the list entry gives the address, so the flag word and the four data vectors of a quad are ONE round trip -- the dense
// scan above has to see a flag word before it may fetch the data behind it.  Two list entries per loop trip, loads of both first.
__global__ void __launch_bounds__(256) field_statistics_quads_kernel(Grid G, const float *rho, const float *u, const uint8_t *flags,
                                                                     const unsigned long long *quads, long long n_slots, StatPartial *partials) {
    StatLocal s;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < n_slots; i0 += 2 * stride) {
        unsigned long long e[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) e[t] = i0 + t * stride < n_slots ? __ldg(quads + i0 + t * stride) : 0ull;      // bit 44 clear: skipped
        unsigned fw[2]; float4 r[2], x[2], y[2], z[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {        // dead slots sit on cell (0, 0, 0): loaded, never counted
            const long long c = ((long long)((int)((e[t] >> 28) & 0xffffu) + G.zg) * G.ny + (int)((e[t] >> 12) & 0xffffu)) * G.nx + (int)(e[t] & 0xfffu) * 4;
            fw[t] = __ldg(reinterpret_cast<const unsigned *>(flags + c));
            r[t] = __ldcs(reinterpret_cast<const float4 *>(rho + c));
            x[t] = __ldcs(reinterpret_cast<const float4 *>(u + c));
            y[t] = __ldcs(reinterpret_cast<const float4 *>(u + G.vol + c));
            z[t] = __ldcs(reinterpret_cast<const float4 *>(u + 2 * G.vol + c));
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (!(e[t] & (1ull << 44))) continue;
            if (!(fw[t] & 0x00000001u)) stat_cell(s, r[t].x, x[t].x, y[t].x, z[t].x);
            if (!(fw[t] & 0x00000100u)) stat_cell(s, r[t].y, x[t].y, y[t].y, z[t].y);
            if (!(fw[t] & 0x00010000u)) stat_cell(s, r[t].z, x[t].z, y[t].z, z[t].z);
            if (!(fw[t] & 0x01000000u)) stat_cell(s, r[t].w, x[t].w, y[t].w, z[t].w);
        }
    }
    const StatPartial b = stat_block_reduce(stat_widen(s));
    if (threadIdx.x == 0) partials[blockIdx.x] = b;
}
__global__ void __launch_bounds__(256) field_statistics_fold_kernel(const StatPartial *partials, int n, double *out) {
    // thread t folds partials t, t + 256, ... in index order; the block reduction order is fixed as well
    StatPartial s = stat_identity();
    for (int i = threadIdx.x; i < n; i += blockDim.x) stat_merge(s, partials[i]);
    s = stat_block_reduce(s);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) out[k] = s.v[k];
    }
}
cudaError_t launch_field_statistics(const Grid &G, const float *rho, const float *u, const uint8_t *flags, const unsigned long long *quads, long long n_slots,
                                    void *scratch, int blocks, double *out, cudaStream_t s) {
    if (quads && flags && n_slots > 0 && (((uintptr_t)rho | (uintptr_t)u) & 15u) == 0 && ((uintptr_t)flags & 3u) == 0) {
        field_statistics_quads_kernel<<<blocks, 256, 0, s>>>(G, rho, u, flags, quads, n_slots, (StatPartial *)scratch);
        field_statistics_fold_kernel<<<1, 256, 0, s>>>((const StatPartial *)scratch, blocks, out);
        return cudaGetLastError();
    }
    const bool vec4 = G.nx % 4 == 0 && G.vol % 4 == 0 && (((uintptr_t)rho | (uintptr_t)u) & 15u) == 0 && ((uintptr_t)flags & 3u) == 0;
    if (vec4) field_statistics_kernel<4><<<blocks, 256, 0, s>>>(G, rho, u, flags, (StatPartial *)scratch);
    else field_statistics_kernel<1><<<blocks, 256, 0, s>>>(G, rho, u, flags, (StatPartial *)scratch);
    field_statistics_fold_kernel<<<1, 256, 0, s>>>((const StatPartial *)scratch, blocks, out);
    return cudaGetLastError();
}

// ---- host launchers (called from lbm_api.cu) -----------------------------------------------
static inline int grid_for(long long n, int block) { long long g = (n + block - 1) / block; return (int)(g > 148LL * 32 ? 148 * 32 : g); }

cudaError_t launch_init_equilibrium(const Grid &G, int compat, float *g, const float *rho, const float *u, float rho0,
                                    const float u0[3], cudaStream_t s) {
    const int b = 256, gr = grid_for(G.vol, b);
    if (compat == LBM_COMPAT_REFERENCE) init_equilibrium_kernel<LBM_COMPAT_REFERENCE><<<gr, b, 0, s>>>(G, g, rho, u, rho0, u0[0], u0[1], u0[2]);
    else init_equilibrium_kernel<LBM_COMPAT_PHYSICAL><<<gr, b, 0, s>>>(G, g, rho, u, rho0, u0[0], u0[1], u0[2]);
    return cudaGetLastError();
}
cudaError_t launch_v60_geometry(const Grid &G, uint8_t *solid, int32_t *zone, const float geom[5], cudaStream_t s) {
    const int b = 256, gr = grid_for(G.vol, b);
    v60_geometry_kernel<<<gr, b, 0, s>>>(G, solid, zone, geom[0], geom[1], geom[2], geom[3], geom[4]);
    return cudaGetLastError();
}
cudaError_t launch_pack_flags(const Grid &G, uint8_t *flags, const uint8_t *solid, const int32_t *zone, const int32_t *les, cudaStream_t s) {
    const int b = 256, gr = grid_for(G.vol, b);
    pack_flags_kernel<<<gr, b, 0, s>>>(G, flags, solid, zone, les);
    return cudaGetLastError();
}
cudaError_t launch_convert_f(const Grid &G, bool to_reference_f, const float *in, const uint8_t *flags, float *out, cudaStream_t s) {
    const int b = 256, gr = grid_for(G.vol, b);
    if (to_reference_f) convert_f_kernel<true><<<gr, b, 0, s>>>(G, in, flags, out);
    else convert_f_kernel<false><<<gr, b, 0, s>>>(G, in, flags, out);
    return cudaGetLastError();
}
// returns the number of launches through *count
cudaError_t launch_face_bc(const Grid &G, float *rho, const uint8_t *flags, cudaStream_t s, int *count) {
    const int b = 128;
    for (int pass = 0; pass < 5; ++pass) {
        dim3 grid;
        if (pass == 0 || pass == 1 || pass == 4) grid = dim3((G.nx + b - 1) / b, G.ny);
        else if (pass == 2) grid = dim3((G.ny + b - 1) / b, G.nz);
        else grid = dim3((G.nx + b - 1) / b, G.nz);
        face_bc_kernel<<<grid, b, 0, s>>>(G, rho, flags, pass);
    }
    *count = 5;
    return cudaGetLastError();
}
cudaError_t launch_pressure_gradient(const Grid &G, const float *rho, const uint8_t *flags, float *bf, float max_force, float scale, int accumulate,
                                     const unsigned *items, const unsigned long long *ctiles, int n_items, int vec, cudaStream_t s) {
    if (ctiles) {     // packed quad list of the four-cell walls kernel
        if (n_items > 0) pressure_gradient_chord_kernel<<<(n_items + 3) / 4, 128, 0, s>>>(G, rho, flags, bf, max_force, scale, accumulate, ctiles, n_items);
        return cudaGetLastError();
    }
    if (items) {      // the caller's tile list was built for exactly this flag field
        if (n_items > 0) pressure_gradient_tiles_kernel<<<(n_items + 3) / 4, 128, 0, s>>>(G, rho, flags, bf, max_force, scale, accumulate, items, n_items, vec);
        return cudaGetLastError();
    }
    if (G.ny > 65535 || G.nz > 65535) return cudaErrorInvalidValue;
    const int b = G.nx >= 128 ? 128 : 64;
    const dim3 grid((unsigned)((G.nx + b - 1) / b), (unsigned)G.ny, (unsigned)G.nz);
    pressure_gradient_kernel<<<grid, b, 0, s>>>(G, rho, flags, bf, max_force, scale, accumulate);
    return cudaGetLastError();
}
cudaError_t launch_density_drive(const Grid &G, float *rho, const uint8_t *flags, const float *target_z, float rate, float max_adjust,
                                 float rho_min, float rho_max, cudaStream_t s) {
    const int b = 256;
    if (G.nx % 4 == 0) density_drive_kernel<4><<<grid_for((long long)G.nz * G.plane / 4, b), b, 0, s>>>(G, rho, flags, target_z, rate, max_adjust, rho_min, rho_max);
    else density_drive_kernel<1><<<grid_for((long long)G.nz * G.plane, b), b, 0, s>>>(G, rho, flags, target_z, rate, max_adjust, rho_min, rho_max);
    return cudaGetLastError();
}
cudaError_t launch_forchheimer_force(const Grid &G, const float *u, const uint8_t *flags, float *bf, float K, float beta,
                                     float c_darcy, float c_forch, float fmax, cudaStream_t s) {
    if (G.ny > 65535 || G.nz > 65535) return cudaErrorInvalidValue;
    if (G.nx % 4 == 0 && ((uintptr_t)flags & 3u) == 0) {
        const int per_row = G.nx / 4, b = per_row >= 128 ? 128 : 64;
        forchheimer_force_kernel<4><<<dim3((unsigned)((per_row + b - 1) / b), (unsigned)G.ny, (unsigned)G.nz), b, 0, s>>>(G, u, flags, bf, K, beta, c_darcy, c_forch, fmax);
    } else {
        const int b = G.nx >= 128 ? 128 : 64;
        forchheimer_force_kernel<1><<<dim3((unsigned)((G.nx + b - 1) / b), (unsigned)G.ny, (unsigned)G.nz), b, 0, s>>>(G, u, flags, bf, K, beta, c_darcy, c_forch, fmax);
    }
    return cudaGetLastError();
}
cudaError_t launch_bounce_slots(const Grid &G, float *g, const uint8_t *flags, const unsigned long long *nbr, int z_begin, int z_end, cudaStream_t s) {
    if (z_end <= z_begin) return cudaSuccess;
    const int b = 256, gr = grid_for(G.plane * (z_end - z_begin), b);
    bounce_slots_kernel<<<gr, b, 0, s>>>(G, g, flags, nbr, z_begin, z_end);
    return cudaGetLastError();
}
cudaError_t launch_add_reaction(const Grid &G, const float *reaction, const uint8_t *flags, float *bf, cudaStream_t s) {
    const int b = 256;
    const bool vec4 = G.vol % 4 == 0 && (((uintptr_t)reaction | (uintptr_t)bf) & 15u) == 0 && ((uintptr_t)flags & 3u) == 0;
    if (vec4) add_reaction_kernel<4><<<grid_for(G.vol / 4, b), b, 0, s>>>(G, reaction, flags, bf);
    else add_reaction_kernel<1><<<grid_for(G.vol, b), b, 0, s>>>(G, reaction, flags, bf);
    return cudaGetLastError();
}

#endif
}  // namespace lbm
