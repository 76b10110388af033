// Shared device-side definitions for liblbm_b200 (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lbm_b200.h"

namespace lbm {

constexpr int Q = LBM_Q;

// config/core.py:36-38 -- the lattice used for moments, streaming, forcing, opposite table.
__host__ __device__ constexpr int cx(int q) { constexpr int t[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0}; return t[q]; }
__host__ __device__ constexpr int cy(int q) { constexpr int t[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1}; return t[q]; }
__host__ __device__ constexpr int cz(int q) { constexpr int t[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1}; return t[q]; }
// src/core/lbm_algorithms.py:158-164 -- the table the reference's equilibrium uses (quirk Q1).
__host__ __device__ constexpr int ex(int q) { return cx(q); }
__host__ __device__ constexpr int ey(int q) { constexpr int t[Q] = {0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1}; return t[q]; }
__host__ __device__ constexpr int ez(int q) { constexpr int t[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1}; return t[q]; }
// legacy/lbm_solver.py:431-439 evaluated on the config table.
__host__ __device__ constexpr int opp(int q) { constexpr int t[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15}; return t[q]; }
__host__ __device__ constexpr float wq(int q) { return q == 0 ? (float)(1.0 / 3.0) : (q < 7 ? (float)(1.0 / 18.0) : (float)(1.0 / 36.0)); }

// Geometry of one slab as the kernels see it.
struct Grid {
    int nx, ny, nz;        // owned extent
    int zg;                // ghost planes per z side
    int nz_global, z0;
    int per_x, per_y, per_z;
    long long plane;       // nx*ny
    long long vol;         // nx*ny*(nz+2*zg): stride between populations / vector components
};

struct StepArgs {
    Grid g;
    const float *src; float *dst;
    float *rho; const float *u_src; float *u_dst;
    const float *force; const float *phase; const float *blockage;
    const uint8_t *flags;
    int z_begin, z_end;    // owned planes processed by this launch (dense mode): z = z_begin + blockIdx.y * z_stride
    int z_stride;          // 1, or nz - 1 for the launch that takes the two boundary planes of a slab together
    const unsigned *items; // bulk mode: active warp-tiles (32*VEC x-consecutive cells): x_segment | y << 8 | z << 20
    int item_begin, n_items;
    const unsigned *item_mask;   // VEC = 4 walls kernel: per list entry, the lanes that must load (lbm_phys.cuh)
    const unsigned long long *nbr;     // per cell (valid where NEAR): solid-source bits | out-of-box bits << 32
    // packed quad list of the four-cell walls kernel (lbm_phys_chord.cuh, lbm_aux.cu): one u64 per lane slot, one uint2 (first link,
    // links) per tile, one u32 per wall link, one float per wall link (the bounced-back value waiting for the next step)
    const unsigned long long *quads; const uint2 *tile_links; const unsigned *links; float *wall;
    // fused pressure-gradient drive (LBM_FEAT_DRIVE): rho of the previous step, clamp and scale of the force
    const float *rho_src; float drive_max_force, drive_scale;
    int write_macro;
    float tau_water, tau_air, gravity_lu;
    float tau_min, tau_max;
    float mrt_magic;       // physical: 0 = BGK, else (tau - 1/2)(tau_odd - 1/2) of the two-rate MRT collision
    float les_k;           // physical: 18*sqrt(2)*Cs^2 ; reference: (Cs*1)*(Cs*1)
    float porous_darcy, porous_forch;
    float K_lu, beta_lu, c_darcy, c_forch;
};

// e . v for e components in {0,+1,-1}: sum of the non-zero terms in x,y,z order (one rounding per add).
template <int EX, int EY, int EZ>
__device__ __forceinline__ float edot(float vx, float vy, float vz) {
    float acc = 0.0f;
    if (EX != 0) acc = EX > 0 ? vx : -vx;
    if (EY != 0) { float t = EY > 0 ? vy : -vy; acc = (EX != 0) ? acc + t : t; }
    if (EZ != 0) { float t = EZ > 0 ? vz : -vz; acc = (EX != 0 || EY != 0) ? acc + t : t; }
    return acc;
}
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return (ax * bx + ay * by) + az * bz;
}

// PressureGradientDrive.compute_pressure_gradient + the force it accumulates (pressure_gradient_drive.py:124-193, 274-279),
// shared by the stand-alone producer (lbm_aux.cu) and the step kernel that fuses it (lbm_phys_chord.cuh, LBM_FEAT_DRIVE);
// both translation units are compiled with -fmad=false, so the two give the same bits.
// pos: -1 = first cell of the axis (one-sided forward difference), +1 = last cell (backward), 0 = interior (central).
__device__ __forceinline__ float pressure_gradient_diff(float r0, float lo, float hi, int pos) {
    return pos == 0 ? (hi - lo) * 0.5f : (pos < 0 ? hi - r0 : r0 - lo);
}
// Out of line on purpose: four inlined copies per thread (three IEEE divisions, a square root and their slow paths each) grew the
// fused-drive step kernel past the instruction cache -- ncu: 1.5 of 14.4 stall cycles per issue on instruction fetch.
#ifdef __CUDACC__
#define LBM_NOINLINE __noinline__
#else
#define LBM_NOINLINE
#endif
static __device__ LBM_NOINLINE float3 pressure_gradient_force(float r0, float gx, float gy, float gz, float max_force, float scale) {
    const float cs2 = (float)(1.0 / 3.0);
    float fx = 0.0f, fy = 0.0f, fz = 0.0f;
    if (r0 > 1e-12f) {
        fx = -(gx * cs2) / r0; fy = -(gy * cs2) / r0; fz = -(gz * cs2) / r0;
        const float mag = sqrtf(dot3(fx, fy, fz, fx, fy, fz));
        if (mag > max_force) { const float s = max_force / mag; fx = fx * s; fy = fy * s; fz = fz * s; }
    }
    if (scale != 1.0f) { fx = scale * fx; fy = scale * fy; fz = scale * fz; }
    return make_float3(fx, fy, fz);
}
__device__ __forceinline__ void pressure_gradient_value(float r0, float gx, float gy, float gz, float max_force, float scale,
                                                        float &fx, float &fy, float &fz) {
    const float cs2 = (float)(1.0 / 3.0);
    fx = 0.0f; fy = 0.0f; fz = 0.0f;
    if (r0 > 1e-12f) {
        fx = -(gx * cs2) / r0; fy = -(gy * cs2) / r0; fz = -(gz * cs2) / r0;
        const float mag = sqrtf(dot3(fx, fy, fz, fx, fy, fz));
        if (mag > max_force) { const float s = max_force / mag; fx = fx * s; fy = fy * s; fz = fz * s; }
    }
    if (scale != 1.0f) { fx = scale * fx; fy = scale * fy; fz = scale * fz; }
}

// tensor maps of the TMA-staged walls kernel (lbm_phys_tma.cuh), passed as a __grid_constant__ kernel parameter
struct alignas(64) TmaMaps {
    CUtensorMap pops;      // f32 [19][nzp][ny][nx], box 64 x TY
    CUtensorMap pops_wide; // same tensor, box 68 x TY (populations that move in x: one 16-byte halo)
    CUtensorMap force;     // f32 [3][nzp][ny][nx]
    CUtensorMap phase;     // f32 [nzp][ny][nx]
    CUtensorMap flags;     // u8  [nzp][ny][nx]
};

template <int N> struct IC { static constexpr int value = N; };
// compile-time loop: f(IC<0>{}), f(IC<1>{}), ...
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (B < E) { f(IC<B>{}); static_for<B + 1, E>(f); }
}

}  // namespace lbm
