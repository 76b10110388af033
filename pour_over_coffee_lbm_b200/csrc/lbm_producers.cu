// Per-step producers of body_force / phase / rho that run next to the D3Q19 step:
//   MultiphaseFlow3D        src/core/multiphase_3d.py   surface-tension chain (:111-149, :313-332, :354-363) and the
//                                                       phase-field step (:151-197, :334-352, :383-387, :365-381)
//   PrecisePouringSystem    src/physics/precise_pouring.py   nozzle force (:131-163), gradual phase change (:165-196)
// compat = reference arithmetic: IEEE f32, the reference's evaluation order, compiled with -fmad=false (no contraction);
// bit-exact against oracle/producers_ref.py except the Gaussian of the nozzle profile (expf, <= 2 ulp).
//
// The reference runs the chain as 4 + 4 full-grid Taichi kernels (8 fields re-read between them); here it is 2 + 2
// launches, and the nozzle kernels visit the nozzle's bounding box (~10^2 cells) instead of the whole grid.
// All HBM-bound: cell loops on an (x-chunk, y, z) grid, x fastest, 4 cells per thread on 128-bit accesses, no index divisions.
#include <stdint.h>
#include <stdlib.h>
#include "lbm_common.cuh"

namespace lbm {

namespace {

__device__ __forceinline__ float norm3(float x, float y, float z) { return sqrtf(dot3(x, y, z, x, y, z)); }
// ti.max(-1.0, ti.min(1.0, v)) with the reference's operand order (a NaN in v survives, as in the reference)
__device__ __forceinline__ float clamp_pm1(float v) {
    const float t = (1.0f <= v) ? 1.0f : v;
    return (-1.0f >= t) ? -1.0f : t;
}

// ---- cell kernels: VEC x-consecutive cells per thread ------------------------------------------------------------------
// VEC = 4 (nx % 4 == 0, 16-byte aligned fields): every field moves as 128-bit loads / stores; the x-1 / x+1 neighbours of
// the 7-point stencils come from the thread's own vector plus one scalar load on each side (L1 hits: the neighbouring
// thread loads that line anyway).  VEC = 1 is the same code with scalar accesses (ragged nx, LBM_PRODUCERS_VEC=1).
// First measurement of the one-cell-per-thread version at 512^3 (profiles/r01_time_producers_512_scalar.log): 3.0-4.9 TB/s
// on 4-byte accesses -- the reason for the vector path.  The arithmetic of a cell is the same statement sequence for
// every VEC, so results are bit-identical between the two.
template <int VEC> struct Vec;
template <> struct Vec<1> {
    static __device__ __forceinline__ void ld(const float *p, float (&v)[1]) { v[0] = *p; }
    static __device__ __forceinline__ void st(float *p, const float (&v)[1]) { *p = v[0]; }
    static __device__ __forceinline__ unsigned ldflags(const uint8_t *p) { return *p; }
};
template <> struct Vec<4> {
    static __device__ __forceinline__ void ld(const float *p, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ unsigned ldflags(const uint8_t *p) { return *reinterpret_cast<const unsigned *>(p); }   // byte i = cell i
};
// centre values of cells x0 .. x0+VEC-1 of the row at p, and their x-1 / x+1 neighbours (0 where the row ends: only
// non-interior cells would use those, and their results are never stored)
template <int VEC>
__device__ __forceinline__ void load_x(const float *p, bool has_left, bool has_right, float (&ctr)[VEC], float (&xm)[VEC], float (&xp)[VEC]) {
    Vec<VEC>::ld(p, ctr);
    const float l = has_left ? p[-1] : 0.0f, r = has_right ? p[VEC] : 0.0f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) { xm[i] = i == 0 ? l : ctr[i > 0 ? i - 1 : 0]; xp[i] = i == VEC - 1 ? r : ctr[i < VEC - 1 ? i + 1 : 0]; }
}
// store only the lanes in `in` (all of them at once when `full`)
template <int VEC>
__device__ __forceinline__ void store_masked(float *p, const float (&v)[VEC], const bool (&in)[VEC], bool full) {
    if (full) { Vec<VEC>::st(p, v); return; }
#pragma unroll
    for (int i = 0; i < VEC; ++i) if (in[i]) p[i] = v[i];
}
struct CellPos { int x0, y, zp, k; long long c; bool row_interior; };
template <int VEC>
__device__ __forceinline__ bool cell_pos(const Grid &G, CellPos &P) {
    P.x0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC; P.y = blockIdx.y; P.zp = blockIdx.z + G.zg; P.k = G.z0 + P.zp - G.zg;
    if (P.x0 >= G.nx) return false;
    P.c = ((long long)P.zp * G.ny + P.y) * G.nx + P.x0;
    P.row_interior = P.y >= 1 && P.y <= G.ny - 2 && P.k >= 1 && P.k <= G.nz_global - 2;
    return true;
}
template <int VEC>
__device__ __forceinline__ bool lanes_interior(const Grid &G, int x0, bool (&in)[VEC]) {
    bool full = true;
#pragma unroll
    for (int i = 0; i < VEC; ++i) { in[i] = x0 + i >= 1 && x0 + i <= G.nx - 2; full &= in[i]; }
    return full;
}
__device__ __forceinline__ float laplacian7(float xp, float xm, float yp, float ym, float zp, float zm, float c) {
    return (((((xp + xm) + yp) + ym) + zp) + zm) - 6.0f * c;
}

// compute_chemical_potential, multiphase_3d.py:80-109 (run once by standardize_initial_state :542-571; the live step()
// never refreshes mu).  Both loops of the reference touch only the cell itself after the Laplacian, so they fuse.
template <int VEC>
__global__ void mp_chemical_potential_kernel(Grid G, const float *__restrict__ phi, float *__restrict__ laplacian, float *__restrict__ mu,
                                             float kappa) {
    CellPos P;
    if (!cell_pos<VEC>(G, P) || !P.row_interior) return;
    float p0[VEC], xm[VEC], xp[VEC], ym[VEC], yp[VEC], zm[VEC], zq[VEC], lap[VEC], m[VEC];
    bool in[VEC];
    const bool full = lanes_interior<VEC>(G, P.x0, in);
    load_x<VEC>(phi + P.c, P.x0 > 0, P.x0 + VEC < G.nx, p0, xm, xp);
    Vec<VEC>::ld(phi + P.c - G.nx, ym); Vec<VEC>::ld(phi + P.c + G.nx, yp);
    Vec<VEC>::ld(phi + P.c - G.plane, zm); Vec<VEC>::ld(phi + P.c + G.plane, zq);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        lap[i] = laplacian7(xp[i], xm[i], yp[i], ym[i], zq[i], zm[i], p0[i]);
        m[i] = ((p0[i] * p0[i]) * p0[i] - p0[i]) + (-kappa) * lap[i];
    }
    if (laplacian) store_masked<VEC>(laplacian + P.c, lap, in, full);
    store_masked<VEC>(mu + P.c, m, in, full);
}

// compute_gradients, multiphase_3d.py:111-132
template <int VEC>
__global__ void mp_gradients_kernel(Grid G, const float *__restrict__ phi, const float *__restrict__ mu, float *__restrict__ grad_phi,
                                    float *__restrict__ grad_mu, float *__restrict__ normal) {
    CellPos P;
    if (!cell_pos<VEC>(G, P) || !P.row_interior) return;
    const long long n = G.vol, c = P.c;
    bool in[VEC];
    const bool full = lanes_interior<VEC>(G, P.x0, in);
    const bool hl = P.x0 > 0, hr = P.x0 + VEC < G.nx;
    float ctr[VEC], xm[VEC], xp[VEC], ym[VEC], yp[VEC], zm[VEC], zq[VEC], gx[VEC], gy[VEC], gz[VEC];
    load_x<VEC>(phi + c, hl, hr, ctr, xm, xp);
    Vec<VEC>::ld(phi + c - G.nx, ym); Vec<VEC>::ld(phi + c + G.nx, yp);
    Vec<VEC>::ld(phi + c - G.plane, zm); Vec<VEC>::ld(phi + c + G.plane, zq);
#pragma unroll
    for (int i = 0; i < VEC; ++i) { gx[i] = (xp[i] - xm[i]) * 0.5f; gy[i] = (yp[i] - ym[i]) * 0.5f; gz[i] = (zq[i] - zm[i]) * 0.5f; }
    store_masked<VEC>(grad_phi + c, gx, in, full); store_masked<VEC>(grad_phi + n + c, gy, in, full); store_masked<VEC>(grad_phi + 2 * n + c, gz, in, full);
    float nx_[VEC], ny_[VEC], nz_[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float mag = norm3(gx[i], gy[i], gz[i]);
        const bool ok = mag > 1e-10f;
        nx_[i] = ok ? gx[i] / mag : 0.0f; ny_[i] = ok ? gy[i] / mag : 0.0f; nz_[i] = ok ? gz[i] / mag : 0.0f;
    }
    store_masked<VEC>(normal + c, nx_, in, full); store_masked<VEC>(normal + n + c, ny_, in, full); store_masked<VEC>(normal + 2 * n + c, nz_, in, full);
    if (mu && grad_mu) {
        load_x<VEC>(mu + c, hl, hr, ctr, xm, xp);
        Vec<VEC>::ld(mu + c - G.nx, ym); Vec<VEC>::ld(mu + c + G.nx, yp);
        Vec<VEC>::ld(mu + c - G.plane, zm); Vec<VEC>::ld(mu + c + G.plane, zq);
#pragma unroll
        for (int i = 0; i < VEC; ++i) { gx[i] = (xp[i] - xm[i]) * 0.5f; gy[i] = (yp[i] - ym[i]) * 0.5f; gz[i] = (zq[i] - zm[i]) * 0.5f; }
        store_masked<VEC>(grad_mu + c, gx, in, full); store_masked<VEC>(grad_mu + n + c, gy, in, full); store_masked<VEC>(grad_mu + 2 * n + c, gz, in, full);
    }
}

// body_force += s / rho on the fluid lanes with rho > 1e-10 (apply_surface_tension :354-363).  A thread whose cells are
// all solid touches neither rho nor body_force: the solid 65 % of a V60 box costs one flag word per 4 cells.
template <int VEC>
__device__ __forceinline__ void apply_lanes(const Grid &G, long long c, const float *__restrict__ rho, const uint8_t *__restrict__ flags,
                                            float *__restrict__ body_force, const float (&sx)[VEC], const float (&sy)[VEC], const float (&sz)[VEC]) {
    const unsigned fw = Vec<VEC>::ldflags(flags + c);
    bool fluid[VEC], any = false;
#pragma unroll
    for (int i = 0; i < VEC; ++i) { fluid[i] = !((fw >> (8 * i)) & LBM_FLAG_SOLID); any |= fluid[i]; }
    if (!any) return;
    const long long n = G.vol;
    float r[VEC], bx[VEC], by[VEC], bz[VEC];
    Vec<VEC>::ld(rho + c, r);
    Vec<VEC>::ld(body_force + c, bx); Vec<VEC>::ld(body_force + n + c, by); Vec<VEC>::ld(body_force + 2 * n + c, bz);
    bool on[VEC], all = true; any = false;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        on[i] = fluid[i] && r[i] > 1e-10f; all &= on[i]; any |= on[i];
        if (on[i]) { bx[i] = bx[i] + sx[i] / r[i]; by[i] = by[i] + sy[i] / r[i]; bz[i] = bz[i] + sz[i] / r[i]; }
    }
    if (!any) return;
    store_masked<VEC>(body_force + c, bx, on, all); store_masked<VEC>(body_force + n + c, by, on, all); store_masked<VEC>(body_force + 2 * n + c, bz, on, all);
}

// compute_curvature :134-149 + compute_surface_tension_force :313-332 on interior cells, then apply_surface_tension
// :354-363 on every cell (the outer layer applies whatever surface_force holds there: it is never written).
template <int VEC>
__global__ void mp_curvature_force_kernel(Grid G, const float *__restrict__ phi, const float *__restrict__ rho, const uint8_t *__restrict__ flags,
                                          const float *__restrict__ grad_phi, const float *__restrict__ normal, float *__restrict__ curvature,
                                          float *surface_force, float *__restrict__ body_force, float sigma) {
    CellPos P;
    if (!cell_pos<VEC>(G, P)) return;
    const long long n = G.vol, c = P.c;
    float sx[VEC], sy[VEC], sz[VEC];
    if (P.row_interior) {
        bool in[VEC];
        const bool full = lanes_interior<VEC>(G, P.x0, in);
        float nx0[VEC], nxm[VEC], nxp[VEC], ny0[VEC], nym[VEC], nyp[VEC], nz0[VEC], nzm[VEC], nzq[VEC], g0[VEC], g1[VEC], g2[VEC], ph[VEC], curv[VEC];
        load_x<VEC>(normal + c, P.x0 > 0, P.x0 + VEC < G.nx, nx0, nxm, nxp);
        Vec<VEC>::ld(normal + n + c, ny0); Vec<VEC>::ld(normal + n + c - G.nx, nym); Vec<VEC>::ld(normal + n + c + G.nx, nyp);
        Vec<VEC>::ld(normal + 2 * n + c, nz0); Vec<VEC>::ld(normal + 2 * n + c - G.plane, nzm); Vec<VEC>::ld(normal + 2 * n + c + G.plane, nzq);
        Vec<VEC>::ld(grad_phi + c, g0); Vec<VEC>::ld(grad_phi + n + c, g1); Vec<VEC>::ld(grad_phi + 2 * n + c, g2);
        Vec<VEC>::ld(phi + c, ph);
        if (!full && body_force) {                 // x = 0 / nx-1 of an interior row: the stored force is applied as it is
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                if (!in[i]) { sx[i] = surface_force[c + i]; sy[i] = surface_force[n + c + i]; sz[i] = surface_force[2 * n + c + i]; }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            if (!in[i]) { curv[i] = 0.0f; continue; }
            float cv = 0.0f;
            if (norm3(nx0[i], ny0[i], nz0[i]) > 1e-10f) {
                const float dnx = (nxp[i] - nxm[i]) * 0.5f, dny = (nyp[i] - nym[i]) * 0.5f, dnz = (nzq[i] - nzm[i]) * 0.5f;
                cv = (dnx + dny) + dnz;
            }
            curv[i] = cv;
            sx[i] = sy[i] = sz[i] = 0.0f;
            if (fabsf(ph[i]) < 0.9f) {
                const float grad_mag = norm3(g0[i], g1[i], g2[i]);
                if (grad_mag > 1e-10f) {
                    const float fm = (sigma * cv) * grad_mag;
                    sx[i] = fm * nx0[i]; sy[i] = fm * ny0[i]; sz[i] = fm * nz0[i];
                }
            }
        }
        store_masked<VEC>(curvature + c, curv, in, full);
        store_masked<VEC>(surface_force + c, sx, in, full); store_masked<VEC>(surface_force + n + c, sy, in, full);
        store_masked<VEC>(surface_force + 2 * n + c, sz, in, full);
        if (!body_force) return;
    } else {
        if (!body_force) return;
        Vec<VEC>::ld(surface_force + c, sx); Vec<VEC>::ld(surface_force + n + c, sy); Vec<VEC>::ld(surface_force + 2 * n + c, sz);
    }
    apply_lanes<VEC>(G, c, rho, flags, body_force, sx, sy, sz);
}

// ---- the same chain WITHOUT materialising grad_phi / normal / curvature / surface_force ---------------------------------
// What main.py needs from accumulate_surface_tension_pre_collision is body_force; the four fields are diagnostics
// (get_interface_statistics).  surface_force is zero outside the interface band |phi| < 0.9, so a thread first looks at its
// own 4 cells (one 128-bit load of phi, one flag word) and leaves unless a fluid cell lies in the band; a band cell then
// evaluates its own gradient and the gradients of its six neighbours from a 25-point stencil of phi (the neighbours'
// normals are recomputed instead of read back) with exactly the statements of the two-launch version, so body_force is
// bit-identical.  Traffic: 5 B/cell + the stencil around the band instead of 117 B/cell.
// Cells of the outer layer never get a normal / surface_force from the reference's kernels; whatever those arrays hold
// there is used when the caller passes them (normal_outer / force_outer, read on the outer layer only), else zero.
__device__ __forceinline__ float unit_component(float gx, float gy, float gz, int comp) {
    const float mag = norm3(gx, gy, gz);
    const float g = comp == 0 ? gx : (comp == 1 ? gy : gz);
    return mag > 1e-10f ? g / mag : 0.0f;
}
template <int VEC>
__global__ void mp_surface_tension_lean_kernel(Grid G, const float *__restrict__ phi, const float *__restrict__ rho, const uint8_t *__restrict__ flags,
                                               const float *__restrict__ normal_outer, const float *__restrict__ force_outer,
                                               float *__restrict__ body_force, float sigma) {
    CellPos P;
    if (!cell_pos<VEC>(G, P)) return;
    const long long n = G.vol, c = P.c;
    const int nx = G.nx;
    const long long pl = G.plane;
    bool in[VEC];
    lanes_interior<VEC>(G, P.x0, in);
    const unsigned fw = Vec<VEC>::ldflags(flags + c);
    bool fluid[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) fluid[i] = !((fw >> (8 * i)) & LBM_FLAG_SOLID);
    float sx[VEC], sy[VEC], sz[VEC];
    bool work[VEC], any = false;
#pragma unroll
    for (int i = 0; i < VEC; ++i) { sx[i] = sy[i] = sz[i] = 0.0f; work[i] = false; }
    // outer layer: the stored force, if the caller keeps one
    if (force_outer) {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            if (fluid[i] && !(P.row_interior && in[i])) {
                sx[i] = force_outer[c + i]; sy[i] = force_outer[n + c + i]; sz[i] = force_outer[2 * n + c + i];
                work[i] = true; any = true;
            }
    }
    if (P.row_interior) {
        float p0[VEC];
        Vec<VEC>::ld(phi + c, p0);
        bool band[VEC], any_band = false;
#pragma unroll
        for (int i = 0; i < VEC; ++i) { band[i] = in[i] && fluid[i] && fabsf(p0[i]) < 0.9f; any_band |= band[i]; }
        if (any_band) {
            const int y = P.y, k = P.k;
            const bool ym_in = y - 1 >= 1, yp_in = y + 1 <= G.ny - 2, km_in = k - 1 >= 1, kp_in = k + 1 <= G.nz_global - 2;
            // value of phi at (x0 + dx, y + dy, k + dk); 0 beyond the row ends (only cells whose result is unused read those)
            auto at = [&](int dx, int dy, int dk) -> float {
                const int x = P.x0 + dx;
                return (x >= 0 && x < nx) ? phi[c + dx + (long long)dy * nx + (long long)dk * pl] : 0.0f;
            };
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                if (!band[i]) continue;
                const int x = P.x0 + i;
                const float gx = (at(i + 1, 0, 0) - at(i - 1, 0, 0)) * 0.5f, gy = (at(i, 1, 0) - at(i, -1, 0)) * 0.5f,
                            gz = (at(i, 0, 1) - at(i, 0, -1)) * 0.5f;
                const float mag = norm3(gx, gy, gz);
                if (!(mag > 1e-10f)) continue;                                    // normal = 0: no curvature, no force
                const float n0x = gx / mag, n0y = gy / mag, n0z = gz / mag;
                float curv = 0.0f;
                if (norm3(n0x, n0y, n0z) > 1e-10f) {
                    float nxp, nxm, nyp, nym, nzp, nzm;
                    if (x + 1 <= nx - 2) nxp = unit_component((at(i + 2, 0, 0) - at(i, 0, 0)) * 0.5f, (at(i + 1, 1, 0) - at(i + 1, -1, 0)) * 0.5f,
                                                              (at(i + 1, 0, 1) - at(i + 1, 0, -1)) * 0.5f, 0);
                    else nxp = normal_outer ? normal_outer[c + i + 1] : 0.0f;
                    if (x - 1 >= 1) nxm = unit_component((at(i, 0, 0) - at(i - 2, 0, 0)) * 0.5f, (at(i - 1, 1, 0) - at(i - 1, -1, 0)) * 0.5f,
                                                         (at(i - 1, 0, 1) - at(i - 1, 0, -1)) * 0.5f, 0);
                    else nxm = normal_outer ? normal_outer[c + i - 1] : 0.0f;
                    if (yp_in) nyp = unit_component((at(i + 1, 1, 0) - at(i - 1, 1, 0)) * 0.5f, (at(i, 2, 0) - at(i, 0, 0)) * 0.5f,
                                                    (at(i, 1, 1) - at(i, 1, -1)) * 0.5f, 1);
                    else nyp = normal_outer ? normal_outer[n + c + i + nx] : 0.0f;
                    if (ym_in) nym = unit_component((at(i + 1, -1, 0) - at(i - 1, -1, 0)) * 0.5f, (at(i, 0, 0) - at(i, -2, 0)) * 0.5f,
                                                    (at(i, -1, 1) - at(i, -1, -1)) * 0.5f, 1);
                    else nym = normal_outer ? normal_outer[n + c + i - nx] : 0.0f;
                    if (kp_in) nzp = unit_component((at(i + 1, 0, 1) - at(i - 1, 0, 1)) * 0.5f, (at(i, 1, 1) - at(i, -1, 1)) * 0.5f,
                                                    (at(i, 0, 2) - at(i, 0, 0)) * 0.5f, 2);
                    else nzp = normal_outer ? normal_outer[2 * n + c + i + pl] : 0.0f;
                    if (km_in) nzm = unit_component((at(i + 1, 0, -1) - at(i - 1, 0, -1)) * 0.5f, (at(i, 1, -1) - at(i, -1, -1)) * 0.5f,
                                                    (at(i, 0, 0) - at(i, 0, -2)) * 0.5f, 2);
                    else nzm = normal_outer ? normal_outer[2 * n + c + i - pl] : 0.0f;
                    const float dnx = (nxp - nxm) * 0.5f, dny = (nyp - nym) * 0.5f, dnz = (nzp - nzm) * 0.5f;
                    curv = (dnx + dny) + dnz;
                }
                const float fm = (sigma * curv) * mag;
                sx[i] = fm * n0x; sy[i] = fm * n0y; sz[i] = fm * n0z;
                work[i] = true; any = true;
            }
        }
    }
    if (!any) return;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        if (!work[i]) continue;
        const float r = rho[c + i];
        if (r > 1e-10f) {
            body_force[c + i] = body_force[c + i] + sx[i] / r; body_force[n + c + i] = body_force[n + c + i] + sy[i] / r;
            body_force[2 * n + c + i] = body_force[2 * n + c + i] + sz[i] / r;
        }
    }
}

// apply_surface_tension :354-363 alone (MultiphaseFlow3D.step with precollision_applied = False re-applies a stored force)
template <int VEC>
__global__ void mp_apply_surface_tension_kernel(Grid G, const float *__restrict__ surface_force, const float *__restrict__ rho,
                                                const uint8_t *__restrict__ flags, float *__restrict__ body_force) {
    CellPos P;
    if (!cell_pos<VEC>(G, P)) return;
    const long long n = G.vol, c = P.c;
    const unsigned fw = Vec<VEC>::ldflags(flags + c);
    bool any = false;
#pragma unroll
    for (int i = 0; i < VEC; ++i) any |= !((fw >> (8 * i)) & LBM_FLAG_SOLID);
    if (!any) return;
    float sx[VEC], sy[VEC], sz[VEC];
    Vec<VEC>::ld(surface_force + c, sx); Vec<VEC>::ld(surface_force + n + c, sy); Vec<VEC>::ld(surface_force + 2 * n + c, sz);
    apply_lanes<VEC>(G, c, rho, flags, body_force, sx, sy, sz);
}

// update_phase_field_cahn_hilliard :151-197 + apply_phase_separation :334-352: both read phi (old) and write phi_new on
// interior cells, so they fuse; mu == nullptr is the all-zero field the live step() leaves it at.
template <int VEC>
__global__ void mp_phase_update_kernel(Grid G, const float *__restrict__ phi, const float *__restrict__ mu, const float *__restrict__ u,
                                       float *__restrict__ phi_new, float mobility, float dt) {
    CellPos P;
    if (!cell_pos<VEC>(G, P) || !P.row_interior) return;
    const long long n = G.vol, c = P.c;
    bool in[VEC];
    const bool full = lanes_interior<VEC>(G, P.x0, in);
    const bool hl = P.x0 > 0, hr = P.x0 + VEC < G.nx;
    float p0[VEC], pxm[VEC], pxp[VEC], pym[VEC], pyp[VEC], pzm[VEC], pzp[VEC], ux[VEC], uy[VEC], uz[VEC], lap_mu[VEC], pn[VEC];
    load_x<VEC>(phi + c, hl, hr, p0, pxm, pxp);
    Vec<VEC>::ld(phi + c - G.nx, pym); Vec<VEC>::ld(phi + c + G.nx, pyp);
    Vec<VEC>::ld(phi + c - G.plane, pzm); Vec<VEC>::ld(phi + c + G.plane, pzp);
    Vec<VEC>::ld(u + c, ux); Vec<VEC>::ld(u + n + c, uy); Vec<VEC>::ld(u + 2 * n + c, uz);
#pragma unroll
    for (int i = 0; i < VEC; ++i) lap_mu[i] = 0.0f;
    if (mu) {
        float m0[VEC], mxm[VEC], mxp[VEC], mym[VEC], myp[VEC], mzm[VEC], mzp[VEC];
        load_x<VEC>(mu + c, hl, hr, m0, mxm, mxp);
        Vec<VEC>::ld(mu + c - G.nx, mym); Vec<VEC>::ld(mu + c + G.nx, myp);
        Vec<VEC>::ld(mu + c - G.plane, mzm); Vec<VEC>::ld(mu + c + G.plane, mzp);
#pragma unroll
        for (int i = 0; i < VEC; ++i) lap_mu[i] = laplacian7(mxp[i], mxm[i], myp[i], mym[i], mzp[i], mzm[i], m0[i]);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float dx = ux[i] > 0.0f ? p0[i] - pxm[i] : pxp[i] - p0[i];
        const float dy = uy[i] > 0.0f ? p0[i] - pym[i] : pyp[i] - p0[i];
        const float dz = uz[i] > 0.0f ? p0[i] - pzm[i] : pzp[i] - p0[i];
        const float convection = -((ux[i] * dx + uy[i] * dy) + uz[i] * dz);
        const float diffusion = mobility * lap_mu[i];
        float v = clamp_pm1(p0[i] + dt * (convection + diffusion));
        if (fabsf(p0[i]) < 0.99f) {
            const float lap = laplacian7(pxp[i], pxm[i], pyp[i], pym[i], pzp[i], pzm[i], p0[i]);
            const float chem = p0[i] * (p0[i] * p0[i] - 1.0f) - 0.01f * lap;
            v = v + (-0.001f * chem) * dt;
        }
        pn[i] = v;
    }
    store_masked<VEC>(phi_new + c, pn, in, full);
}

// copy_phase_field :383-387 + update_density_from_phase :365-381 over every cell (src == dst: density only).
template <int VEC>
__global__ void mp_copy_density_kernel(Grid G, const float *src, float *dst, float *__restrict__ rho, float *__restrict__ phase, float rho_air,
                                       float drho) {
    CellPos P;
    if (!cell_pos<VEC>(G, P)) return;
    float raw[VEC], r[VEC], ph[VEC];
    Vec<VEC>::ld(src + P.c, raw);
    if (dst != src) Vec<VEC>::st(dst + P.c, raw);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float p1 = clamp_pm1(raw[i]) + 1.0f;
        r[i] = rho_air + (drho * p1) / 2.0f;
        ph[i] = p1 / 2.0f;
    }
    Vec<VEC>::st(rho + P.c, r); Vec<VEC>::st(phase + P.c, ph);
}

// FilterPaperSystem.update_dynamic_resistance, filter_paper.py:703-746 (filter-zone cells):
// blockage <- 0.95 blockage + 0.05 * 0.9 (1 - exp(-0.1 accumulated)); accumulated *= 0.999.  The blockage field is an
// input of the step kernel's filter damping (apply_filter_effects :578-586).  The zone is a one-cell shell: a thread scans
// 16 flag bytes with one 128-bit load (cell count and base % 16 == 0; else 4 or 1) and touches the two f32 fields only
// where the bit is set.  (First version: an (x, y, z) launch grid with 32-thread blocks per row -- 0.14 ms at 512^3, a
// third of the flag bytes' DRAM time lost to block scheduling.)
template <int CELLS>
__global__ void __launch_bounds__(256) dynamic_resistance_kernel(long long begin, long long count, const uint8_t *__restrict__ flags,
                                                                 float *__restrict__ blockage, float *__restrict__ accumulated) {
    // no stencil: the owned cells are one contiguous range [begin, begin + count), CELLS of them per thread
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t * CELLS >= count) return;
    const long long c = begin + t * CELLS;
    constexpr int WORDS = CELLS >= 4 ? CELLS / 4 : 1;
    unsigned w[WORDS];
    if constexpr (CELLS == 16) { const uint4 q = *reinterpret_cast<const uint4 *>(flags + c); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
    else if constexpr (CELLS == 4) w[0] = *reinterpret_cast<const unsigned *>(flags + c);
    else w[0] = flags[c];
    const unsigned filter_bits = LBM_FLAG_FILTER * 0x01010101u;
    bool any = false;
#pragma unroll
    for (int j = 0; j < WORDS; ++j) any |= (w[j] & filter_bits) != 0u;
    if (!any) return;
#pragma unroll
    for (int i = 0; i < CELLS; ++i) {
        if (!((w[i / 4] >> (8 * (i % 4))) & LBM_FLAG_FILTER)) continue;
        const float acc = accumulated[c + i];
        const float nb = 0.9f * (1.0f - expf(-0.1f * acc));
        blockage[c + i] = 0.95f * blockage[c + i] + 0.05f * nb;
        accumulated[c + i] = acc * 0.999f;
    }
}

// Counter-based uniform [0, 1): the reference draws ti.random() from Taichi's unseeded per-thread generator, which no
// implementation can reproduce; here a draw is a pure function of (seed, particle, draw index) -- lowbias32 hash.
__device__ __forceinline__ float uniform01(unsigned seed, unsigned p, unsigned d) {
    unsigned h = seed ^ (p * 0x9E3779B9u) ^ (d * 0x85EBCA6Bu);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}
// FilterPaperSystem.block_particles_at_filter, filter_paper.py:616-700: a particle over a filter-zone cell (5 planes
// around its own) that moves down bounces with restitution 0.3, gets a small horizontal kick, and leaves 0.01 in
// accumulated_particles.  The reference divides the (lattice-unit) position by SCALE_LENGTH here (quirk Q9: mixed
// units) -- reproduced, the caller passes the divisor.
__global__ void particles_block_at_filter_kernel(Grid G, lbm_particles P, const uint8_t *__restrict__ flags, float *__restrict__ accumulated,
                                                 float scale_length, float noise, unsigned seed) {
    const int n = P.n, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || P.active[p] == 0) return;
    const int gx = (int)(P.pos[p] / scale_length), gy = (int)(P.pos[n + p] / scale_length), gz = (int)(P.pos[2 * n + p] / scale_length);
    if (gx < 0 || gx >= G.nx || gy < 0 || gy >= G.ny || gz < 0 || gz >= G.nz_global) return;
    for (int off = -2; off <= 2; ++off) {
        const int k = gz + off;
        if (k < 0 || k >= G.nz_global) continue;
        const int zp = k - G.z0 + G.zg;
        if (zp < 0 || zp >= G.nz + 2 * G.zg) continue;                 // beyond this slab's planes
        const long long c = ((long long)zp * G.ny + gy) * G.nx + gx;
        if (!(flags[c] & LBM_FLAG_FILTER)) continue;
        const float vz = P.vel[2 * n + p];
        if (vz < 0.0f) {
            P.vel[2 * n + p] = (-vz) * 0.3f;
            P.vel[p] = P.vel[p] + (uniform01(seed, (unsigned)p, 0u) - 0.5f) * noise;
            P.vel[n + p] = P.vel[n + p] + (uniform01(seed, (unsigned)p, 1u) - 0.5f) * noise;
            atomicAdd(accumulated + c, 0.01f);
        }
        break;
    }
}

struct PourArgs {
    float pour_x, pour_y, radius; int pour_z;
    float velocity, flow_rate, dt;
    float decay[5];                 // exp(-d / 2.0), d = pour_z - k = 0..4: a constant expression in the reference (folded in f64)
    int x0, y0, k0, wx, wy, wk;     // bounding box of the nozzle
};
// _is_in_pouring_region, precise_pouring.py:99-129
__device__ __forceinline__ float pour_intensity(const PourArgs &P, int x, int y, int k) {
    const float dx = (float)x - P.pour_x, dy = (float)y - P.pour_y;
    const float dist = sqrtf(dx * dx + dy * dy);
    const int d = P.pour_z - k;
    if (!(dist <= P.radius) || d < 0 || d > 4) return 0.0f;
    const float t = dist / P.radius;
    return expf(-0.5f * (t * t)) * P.decay[d];
}
// apply_pouring_force :131-163 (MODE 0) and apply_gradual_phase_change :165-196 (MODE 1) over the nozzle's bounding box
template <int MODE>
__global__ void pour_kernel(Grid G, PourArgs P, const uint8_t *__restrict__ flags, float *__restrict__ field) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.wx * P.wy * P.wk) return;
    const int x = P.x0 + t % P.wx, y = P.y0 + (t / P.wx) % P.wy, k = P.k0 + t / (P.wx * P.wy);
    const int zp = k - G.z0 + G.zg;
    if (zp < G.zg || zp >= G.zg + G.nz) return;                       // another slab's plane
    const long long n = G.vol, c = ((long long)zp * G.ny + y) * G.nx + x;
    if (flags[c] & LBM_FLAG_SOLID) return;
    const float total = pour_intensity(P, x, y, k);
    if (!(total > 0.0f)) return;
    if (MODE == 0) {
        float accel = 0.0f;
        if (P.dt > 1e-8f) accel = ((P.velocity * total) * P.flow_rate) / P.dt;
        accel = accel <= 10.0f ? accel : 10.0f;
        field[c] = field[c] + 0.0f; field[n + c] = field[n + c] + 0.0f; field[2 * n + c] = field[2 * n + c] + (-accel);
    } else {
        const float cur = field[c];
        float rate = (1.0f - cur) / 0.05f;
        rate = (2.0f <= rate) ? 2.0f : rate;
        rate = (-2.0f >= rate) ? -2.0f : rate;
        const float change = ((rate * total) * P.dt) * P.flow_rate;
        field[c] = clamp_pm1(cur + change);
    }
}

// launch geometry: one thread per VEC x-consecutive cells on an (x-chunk, y, owned z) grid -- no index divisions
inline bool scalar_forced() {
    static const bool v = [] { const char *e = getenv("LBM_PRODUCERS_VEC"); return e && atoi(e) == 1; }();
    return v;
}
template <class... Ptr> inline bool aligned16(Ptr... p) { return (((uintptr_t)p | ...) & 15u) == 0; }
inline int pick_vec(const Grid &G, bool aligned) { return (G.nx % 4 == 0 && aligned && !scalar_forced()) ? 4 : 1; }
inline int cell_block(const Grid &G, int vec) { const int t = (G.nx + vec - 1) / vec; return t >= 128 ? 128 : (t >= 64 ? 64 : 32); }
inline dim3 cell_grid(const Grid &G, int vec, int b) { const int t = (G.nx + vec - 1) / vec; return dim3((unsigned)((t + b - 1) / b), (unsigned)G.ny, (unsigned)G.nz); }
inline bool grid_ok(const Grid &G) { return G.ny <= 65535 && G.nz <= 65535; }
inline int dynamic_resistance_cells(long long begin, long long count, const uint8_t *flags, bool scalar) {
    if (scalar) return 1;
    if (begin % 16 == 0 && count % 16 == 0 && ((uintptr_t)flags & 15u) == 0) return 16;
    if (begin % 4 == 0 && count % 4 == 0 && ((uintptr_t)flags & 3u) == 0) return 4;
    return 1;
}
#ifndef LBM_EMULATE_ON_HOST      /* tests/emu compiles the kernels above with g++ and runs them thread by thread */
#define LAUNCH_CELLS(kernel, vec, s, ...)                                                              \
    do {                                                                                               \
        if ((vec) == 4) { const int b_ = cell_block(G, 4); kernel<4><<<cell_grid(G, 4, b_), b_, 0, s>>>(__VA_ARGS__); } \
        else { const int b_ = cell_block(G, 1); kernel<1><<<cell_grid(G, 1, b_), b_, 0, s>>>(__VA_ARGS__); }            \
    } while (0)

#endif
}  // namespace

#ifndef LBM_EMULATE_ON_HOST
cudaError_t launch_chemical_potential(const Grid &G, const float *phi, float *laplacian, float *mu, float kappa, cudaStream_t s) {
    if (!grid_ok(G)) return cudaErrorInvalidValue;
    const int vec = pick_vec(G, aligned16(phi, laplacian, mu));
    LAUNCH_CELLS(mp_chemical_potential_kernel, vec, s, G, phi, laplacian, mu, kappa);
    return cudaGetLastError();
}
// stage bit 0: compute_gradients, bit 1: curvature + force + apply.  On z-slabs the caller refreshes the ghost planes of
// `normal` between the two stages (the curvature stencil reads n_z at k -+ 1).
cudaError_t launch_surface_tension(const Grid &G, int stages, const float *phi, const float *mu, const float *rho, const uint8_t *flags,
                                   float *grad_phi, float *grad_mu, float *normal, float *curvature, float *surface_force, float *body_force,
                                   float sigma, cudaStream_t s) {
    if (!grid_ok(G)) return cudaErrorInvalidValue;
    const int vec = pick_vec(G, aligned16(phi, mu, rho, grad_phi, grad_mu, normal, curvature, surface_force, body_force) && ((uintptr_t)flags & 3u) == 0);
    if (stages & 1) LAUNCH_CELLS(mp_gradients_kernel, vec, s, G, phi, mu, grad_phi, grad_mu, normal);
    if (stages & 2) LAUNCH_CELLS(mp_curvature_force_kernel, vec, s, G, phi, rho, flags, grad_phi, normal, curvature, surface_force, body_force, sigma);
    return cudaGetLastError();
}
cudaError_t launch_surface_tension_lean(const Grid &G, const float *phi, const float *rho, const uint8_t *flags, const float *normal_outer,
                                        const float *force_outer, float *body_force, float sigma, cudaStream_t s) {
    if (!grid_ok(G)) return cudaErrorInvalidValue;
    const int vec = pick_vec(G, aligned16(phi) && ((uintptr_t)flags & 3u) == 0);
    LAUNCH_CELLS(mp_surface_tension_lean_kernel, vec, s, G, phi, rho, flags, normal_outer, force_outer, body_force, sigma);
    return cudaGetLastError();
}
cudaError_t launch_apply_surface_tension(const Grid &G, const float *surface_force, const float *rho, const uint8_t *flags, float *body_force,
                                         cudaStream_t s) {
    if (!grid_ok(G)) return cudaErrorInvalidValue;
    const int vec = pick_vec(G, aligned16(surface_force, rho, body_force) && ((uintptr_t)flags & 3u) == 0);
    LAUNCH_CELLS(mp_apply_surface_tension_kernel, vec, s, G, surface_force, rho, flags, body_force);
    return cudaGetLastError();
}
cudaError_t launch_phase_field_step(const Grid &G, float *phi, float *phi_new, const float *mu, const float *u, float *rho, float *phase,
                                    float mobility, float dt, float rho_air, float drho, cudaStream_t s) {
    if (!grid_ok(G)) return cudaErrorInvalidValue;
    const int vec = pick_vec(G, aligned16(phi, phi_new, mu, u, rho, phase));
    LAUNCH_CELLS(mp_phase_update_kernel, vec, s, G, phi, mu, u, phi_new, mobility, dt);
    LAUNCH_CELLS(mp_copy_density_kernel, vec, s, G, phi_new, phi, rho, phase, rho_air, drho);
    return cudaGetLastError();
}
cudaError_t launch_density_from_phase(const Grid &G, const float *phi, float *rho, float *phase, float rho_air, float drho, cudaStream_t s) {
    if (!grid_ok(G)) return cudaErrorInvalidValue;
    const int vec = pick_vec(G, aligned16(phi, rho, phase));
    LAUNCH_CELLS(mp_copy_density_kernel, vec, s, G, phi, const_cast<float *>(phi), rho, phase, rho_air, drho);
    return cudaGetLastError();
}
cudaError_t launch_dynamic_resistance(const Grid &G, const uint8_t *flags, float *blockage, float *accumulated, cudaStream_t s) {
    const long long begin = (long long)G.zg * G.plane, count = (long long)G.nz * G.plane;
    const int cells = dynamic_resistance_cells(begin, count, flags, scalar_forced());
    const long long threads = (count + cells - 1) / cells, blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (blocks == 0) return cudaSuccess;
    if (cells == 16) dynamic_resistance_kernel<16><<<(unsigned)blocks, 256, 0, s>>>(begin, count, flags, blockage, accumulated);
    else if (cells == 4) dynamic_resistance_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(begin, count, flags, blockage, accumulated);
    else dynamic_resistance_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(begin, count, flags, blockage, accumulated);
    return cudaGetLastError();
}
cudaError_t launch_particles_block_at_filter(const Grid &G, const lbm_particles &ps, const uint8_t *flags, float *accumulated, float scale_length,
                                             float noise, unsigned seed, cudaStream_t s) {
    const int b = 256, gr = (ps.n + b - 1) / b;
    if (ps.n > 0) particles_block_at_filter_kernel<<<gr, b, 0, s>>>(G, ps, flags, accumulated, scale_length, noise, seed);
    return cudaGetLastError();
}
// mode 0: body_force (3 components), mode 1: phi.  *launched = 0 when the nozzle's box misses this slab.
cudaError_t launch_pour(const Grid &G, const lbm_pour &pr, const float decay[5], int mode, const uint8_t *flags, float *field, cudaStream_t s,
                        int *launched) {
    PourArgs P{};
    P.pour_x = pr.pour_x; P.pour_y = pr.pour_y; P.radius = pr.radius; P.pour_z = pr.pour_z;
    P.velocity = pr.velocity; P.flow_rate = pr.flow_rate; P.dt = pr.dt;
    for (int i = 0; i < 5; ++i) P.decay[i] = decay[i];
    const int x0 = max(0, (int)floorf(pr.pour_x - pr.radius)), x1 = min(G.nx - 1, (int)ceilf(pr.pour_x + pr.radius));
    const int y0 = max(0, (int)floorf(pr.pour_y - pr.radius)), y1 = min(G.ny - 1, (int)ceilf(pr.pour_y + pr.radius));
    const int k0 = max(max(0, G.z0), pr.pour_z - 4), k1 = min(min(G.nz_global - 1, G.z0 + G.nz - 1), pr.pour_z);
    *launched = 0;
    if (x1 < x0 || y1 < y0 || k1 < k0) return cudaSuccess;
    P.x0 = x0; P.y0 = y0; P.k0 = k0; P.wx = x1 - x0 + 1; P.wy = y1 - y0 + 1; P.wk = k1 - k0 + 1;
    const long long cells = (long long)P.wx * P.wy * P.wk;
    if (cells > (1LL << 30)) return cudaErrorInvalidValue;
    const int b = 128, gr = (int)((cells + b - 1) / b);
    if (mode == 0) pour_kernel<0><<<gr, b, 0, s>>>(G, P, flags, field);
    else pour_kernel<1><<<gr, b, 0, s>>>(G, P, flags, field);
    *launched = 1;
    return cudaGetLastError();
}

#endif  // LBM_EMULATE_ON_HOST

}  // namespace lbm
