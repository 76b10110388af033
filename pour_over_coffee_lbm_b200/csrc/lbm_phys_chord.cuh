// compat = physical behind walls, four cells per thread on chord-fitted tiles, populations staged in shared memory
// (sm_100a) -- the default walls kernel.
//
// Why (ncu, V60 512^3, B200).  The two-cell register-staged kernel (lbm_phys.cuh) ran at 57 % of DRAM peak with 927
// warp-instructions per 64 cells; a first four-cell version that kept the dense kernel's structure (19 x LDG.128 into
// registers, shuffles) did no better: 1637 instructions per 128-cell tile of which 508 were arithmetic -- ~400 register moves
// (shifting / packing cell pairs for the f32x2 collision, unpacking for 128-bit stores), ~250 64-bit address updates, 58
// spill instructions -- and 19 % of all stall samples sat on two spill STOREs inside the load phase: a loaded value has to
// arrive before it can be spilled, so the warp's loads went out in two or three serialised round trips
// (profiles/r02_chord_v1_*).  This version takes the populations out of the register file while they are in flight:
//   * tiles are CHORD-FITTED (lbm_aux.cu, build_chord_lists): up to 32 consecutive quads (4 cells, 16-byte aligned) of one
//     row starting at the chord's first active quad, with a 32-bit lane mask; 48.2 M cell slots are launched for the 47.7 M
//     fluid cells of the V60 512^3 mask (x-aligned 64-cell tiles: 58.9 M).  Lanes outside the mask load and store nothing;
//   * every live lane issues 19 cp.async of 16 bytes (global -> shared, no register, L1 bypassed) into the warp's private
//     staging rows, one row of 4 + 128 + 4 floats per population: lane l owns words [4 + 4l, 8 + 4l).  The one-cell shift in
//     x of the 10 moving populations is an ADDRESS offset when the row is read back (cells x-1 .. x+2 are words 3 + 4l ..),
//     so there are no shuffles and no register moves; the x-1 / x+4 neighbour that no live lane brings (first / last lane of
//     a chord, row ends, periodic wrap) arrives by a 4-byte cp.async in the word next to the lane's own;
//   * the two cell pairs of a thread are collided one after the other (packed f32x2); each pair reads its 19 inputs from the
//     row (LDS.64, or two LDS.32 for the shifted ones) directly into register pairs and writes its results back into the
//     lane's own words (STS.64), so only one pair is in registers at a time;
//   * write-back: LDS.128 + 128-bit streaming store per population for all-fluid quads.  Everything irregular is a
//     precomputed LINK of the tile (one u32 per (cell, population)): the wall links of halfway bounce-back on the write side
//     (post-collision f_q of a fluid cell -> slot opp(q) of its solid neighbour, lbm_phys.cuh) and the "self" links of the
//     fluid cells of a quad a chord ends in (19 each).  All 32 lanes walk the link list together, one link per lane and
//     round (LDS + one 4-byte store); a tile without links skips it on a warp-uniform branch.  The neighbour masks are not
//     read by this kernel, the flag word only for the solid / filter / LES bits;
//   * LBM_FEAT_DRIVE: the pressure-gradient drive (pressure_gradient_drive.py:124-193) is evaluated from the PREVIOUS
//     step's rho inside this kernel (same statements as the stand-alone producer, lbm_common.cuh) while the populations
//     are still in flight, and added to the body force: the separate producer pass and its 12 B force round trip disappear.
#pragma once
#include "lbm_phys.cuh"

namespace lbm {
#ifndef LBM_EMULATE_ON_HOST

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float lds32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ P2 lds64(unsigned a) { P2 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v.v) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(unsigned a, P2 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v.v) : "memory"); }
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}

constexpr int CHORD_ROW = 4 + 128 + 4;                // floats per staged population row: left edge | 32 lanes x 4 | right edge
constexpr int CHORD_WARP_BYTES = Q * CHORD_ROW * 4;   // 10336 B per warp

template <bool FORCED, bool LES, bool POROUS, bool DRIVE, int BLOCK, bool COLLIDE, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) phys_chord_kernel(const __grid_constant__ StepArgs P) {
    constexpr bool HAS_F = FORCED || DRIVE;
    __shared__ __align__(16) float stage[BLOCK / 32][Q][CHORD_ROW];
    const Grid &G = P.g;
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const int w = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
    if (w >= P.n_items) return;                                          // warp-uniform
    const uint4 e = __ldg(P.ctiles + P.item_begin + w);
    const unsigned lmask = e.z, n_links = e.x >> 12;
    const bool live = ((lmask >> lane) & 1u) != 0;
    const int x0 = ((int)(e.x & 0xfffu) + (live ? (int)lane : 0)) * 4;  // dead lanes shadow lane 0 (always live) for their addresses
    const int y = (int)(e.y & 0xffffu), z = (int)(e.y >> 16);
    const int zp = z + G.zg;
    const unsigned row0 = ((unsigned)zp * (unsigned)G.ny + (unsigned)y) * (unsigned)G.nx;
    const unsigned own = row0 + (unsigned)x0;
    const unsigned vol = (unsigned)G.vol;

    // neighbour rows as 32-bit index deltas: periodic wrap, else clamp (a clamped source lies outside an open face and is
    // replaced by w_q below)
    const int nxi = G.nx, plane = (int)G.plane;
    int dym = -nxi; if (y == 0) dym = G.per_y ? (G.ny - 1) * nxi : 0;
    int dyq = nxi; if (y == G.ny - 1) dyq = G.per_y ? -(G.ny - 1) * nxi : 0;
    int dzm = -plane, dzq = plane;
    if (!G.zg) {
        if (z == 0) dzm = G.per_z ? (G.nz - 1) * plane : 0;
        if (z == G.nz - 1) dzq = G.per_z ? -(G.nz - 1) * plane : 0;
    }
    // shared-window address of this lane's words of row 0; row q is q * CHORD_ROW * 4 bytes further
    // (dead lanes read lane 0's words -- benign values for the arithmetic they run along with the warp -- and write nothing)
    const unsigned s_own = (unsigned)__cvta_generic_to_shared(&stage[wib][0][4 + 4 * (live ? lane : 0u)]);
    constexpr unsigned ROWB = CHORD_ROW * 4;

    // (1) populations: global -> shared, nothing held in registers while in flight
    if (live) {
        const float *rowp[3][3];
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
                rowp[dz + 1][dy + 1] = P.src + (own + (unsigned)(dy < 0 ? dym : (dy > 0 ? dyq : 0)) + (unsigned)(dz < 0 ? dzm : (dz > 0 ? dzq : 0)));
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            cp_async16(s_own + q * ROWB, plane_of(rowp[1 - cz(q)][1 - cy(q)], vol, q));
        });
        // x-1 / x+4 neighbour of the quad when no live lane brings it: first / last lane, a gap in the lane mask, row ends
        if (lane == 0 || !((lmask >> (lane - 1)) & 1u)) {
            int dxm = -1; if (x0 == 0) dxm = G.per_x ? G.nx - 1 : 0;
            static_for<0, Q>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                if constexpr (cx(q) > 0) cp_async4(s_own + q * ROWB - 4, plane_of(rowp[1 - cz(q)][1 - cy(q)], vol, q) + dxm);
            });
        }
        if (lane == 31 || !((lmask >> (lane + 1)) & 1u)) {
            int dxq = 4; if (x0 == G.nx - 4) dxq = G.per_x ? -(G.nx - 4) : 3;
            static_for<0, Q>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                if constexpr (cx(q) < 0) cp_async4(s_own + q * ROWB + 16, plane_of(rowp[1 - cz(q)][1 - cy(q)], vol, q) + dxq);
            });
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // (2) flags, body force, phase, rho stencil of the fused drive: plain loads (registers are free while the populations travel)
    const unsigned flag_word = __ldg(reinterpret_cast<const unsigned *>(P.flags + own));
    bool has_force = false, has_phase = false;
    float F[3][4], ph[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { F[0][c] = F[1][c] = F[2][c] = 0.0f; ph[c] = 0.0f; }
    if constexpr (HAS_F) {
        has_phase = FORCED && P.phase != nullptr;
        has_force = DRIVE || (FORCED && (P.force != nullptr || (has_phase && P.gravity_lu != 0.0f)));
        float4 bf[3];
        bool have_bf = false;
        if constexpr (FORCED) {
            have_bf = P.force != nullptr;
            if (have_bf) {
#pragma unroll
                for (int d = 0; d < 3; ++d) bf[d] = __ldg(reinterpret_cast<const float4 *>(plane_of(P.force + own, vol, d)));
            }
            if (has_phase) { const float4 t = __ldg(reinterpret_cast<const float4 *>(P.phase + own)); ph[0] = t.x; ph[1] = t.y; ph[2] = t.z; ph[3] = t.w; }
        }
        if constexpr (DRIVE) {
            const float *pr = P.rho_src + own;
            const float4 r = __ldg(reinterpret_cast<const float4 *>(pr));
            const float4 rym = __ldg(reinterpret_cast<const float4 *>(pr + dym)), ryp = __ldg(reinterpret_cast<const float4 *>(pr + dyq));
            const float4 rzm = __ldg(reinterpret_cast<const float4 *>(pr + dzm)), rzp = __ldg(reinterpret_cast<const float4 *>(pr + dzq));
            const float rxm = __ldg(pr - (x0 > 0 ? 1 : 0)), rxp = __ldg(pr + (x0 + 4 < nxi ? 4 : 3));
            const float r0[4] = {r.x, r.y, r.z, r.w}, lo[4] = {rxm, r.x, r.y, r.z}, hi[4] = {r.y, r.z, r.w, rxp};
            const float ym[4] = {rym.x, rym.y, rym.z, rym.w}, yq[4] = {ryp.x, ryp.y, ryp.z, ryp.w};
            const float zm[4] = {rzm.x, rzm.y, rzm.z, rzm.w}, zq[4] = {rzp.x, rzp.y, rzp.z, rzp.w};
            const int kg = G.z0 + z;
            const int ypos = y == 0 ? -1 : (y == G.ny - 1 ? 1 : 0), zpos = kg == 0 ? -1 : (kg == G.nz_global - 1 ? 1 : 0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int x = x0 + c;
                const float gx = pressure_gradient_diff(r0[c], lo[c], hi[c], x == 0 ? -1 : (x == nxi - 1 ? 1 : 0));
                const float gy = pressure_gradient_diff(r0[c], ym[c], yq[c], ypos);
                const float gz = pressure_gradient_diff(r0[c], zm[c], zq[c], zpos);
                pressure_gradient_value(r0[c], gx, gy, gz, P.drive_max_force, P.drive_scale, F[0][c], F[1][c], F[2][c]);
            }
        }
        if (have_bf) {      // body_force (+ drive: the sum the producer leaves in body_force in accumulate mode)
            const float b[3][4] = {{bf[0].x, bf[0].y, bf[0].z, bf[0].w}, {bf[1].x, bf[1].y, bf[1].z, bf[1].w}, {bf[2].x, bf[2].y, bf[2].z, bf[2].w}};
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int c = 0; c < 4; ++c) F[d][c] = DRIVE ? b[d][c] + F[d][c] : b[d][c];
        }
    }
    unsigned fl[4];
    bool mine[4], all_mine = true;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        fl[c] = (flag_word >> (8 * c)) & 0xffu;
        mine[c] = live && !(fl[c] & LBM_FLAG_SOLID);
        all_mine &= mine[c];
    }
    // open faces: sources outside the box deliver w_q
    bool ylo = false, yhi = false, zlo = false, zhi = false, xlo = false, xhi = false;
    if (!(G.per_x && G.per_y && G.per_z)) {
        const int zglob = G.z0 + z;
        ylo = !G.per_y && y == 0; yhi = !G.per_y && y == G.ny - 1;
        zlo = !G.per_z && zglob == 0; zhi = !G.per_z && zglob == G.nz_global - 1;
        xlo = !G.per_x && x0 == 0; xhi = !G.per_x && x0 == G.nx - 4;                // cell 0 / cell 3 of this thread
    }
    const bool on_face = ylo || yhi || zlo || zhi || xlo || xhi;

    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();

    // (3) the two cell pairs, one after the other.  Pair h = cells 2h, 2h + 1 of the quad; population q of those cells:
    //     cx = 0: words 2h, 2h + 1 of the lane; cx > 0 (source x - 1): words 2h - 1, 2h; cx < 0 (source x + 1): words 2h + 1, 2h + 2.
    //     Before pair 0 writes its results into words 0, 1, every word of pair 1 that a result could overwrite is taken:
    //     the lane's own word 1 (cx > 0) and the next lane's word 0 (cx < 0).
    float keep[Q];
    static_for<0, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        if constexpr (cx(q) > 0) keep[q] = lds32(s_own + q * ROWB + 4);
        if constexpr (cx(q) < 0) keep[q] = lds32(s_own + q * ROWB + 16);
    });
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        P2 fp[Q];
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            const unsigned a = s_own + q * ROWB + 8 * h;
            if constexpr (cx(q) == 0) fp[q] = lds64(a);
            else if constexpr (cx(q) > 0) fp[q] = h == 0 ? p2_make(lds32(a - 4), lds32(a)) : p2_make(keep[q], lds32(a));
            else fp[q] = h == 0 ? p2_make(lds32(a + 4), lds32(a + 8)) : p2_make(lds32(a + 4), keep[q]);
        });
        if (on_face) {
            static_for<1, Q>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                const bool row_out = (cy(q) > 0 && ylo) || (cy(q) < 0 && yhi) || (cz(q) > 0 && zlo) || (cz(q) < 0 && zhi);
                const bool out0 = row_out || (cx(q) > 0 && h == 0 && xlo), out1 = row_out || (cx(q) < 0 && h == 1 && xhi);
                if (out0 || out1) fp[q] = p2_make(out0 ? wq(q) : p2_lo(fp[q]), out1 ? wq(q) : p2_hi(fp[q]));
            });
        }
        CellIn<P2> in;
        in.Fx = p2_make(F[0][2 * h], F[0][2 * h + 1]); in.Fy = p2_make(F[1][2 * h], F[1][2 * h + 1]); in.Fz = p2_make(F[2][2 * h], F[2][2 * h + 1]);
        in.phase = p2_make(ph[2 * h], ph[2 * h + 1]);
        in.flag[0] = fl[2 * h]; in.flag[1] = fl[2 * h + 1];
        CellMacro<P2> mac;
        collide_phys<P2, HAS_F, LES, POROUS, COLLIDE>(fp, in, mac, P, has_phase, has_force);
        if constexpr (COLLIDE) {
            if (h == 0) __syncwarp();                                    // every lane holds its pair-1 words (keep[]) and has read pair 0
            if (live) {
                static_for<0, Q>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    sts64(s_own + q * ROWB + 8 * h, fp[q]);
                });
            }
        }
        // rho, u of this pair: 64-bit stores now (default cache policy: the two halves of a sector meet in L2) instead of eight
        // more registers held through the second collision
        if (P.write_macro) {
            if (mine[2 * h] && mine[2 * h + 1]) {
                float *pu = P.u_dst + own + 2 * h;
                *reinterpret_cast<unsigned long long *>(P.rho + own + 2 * h) = mac.rho.v;
                *reinterpret_cast<unsigned long long *>(pu) = mac.ux.v;
                *reinterpret_cast<unsigned long long *>(plane_of(pu, vol, 1)) = mac.uy.v;
                *reinterpret_cast<unsigned long long *>(plane_of(pu, vol, 2)) = mac.uz.v;
            } else {
#pragma unroll
                for (int l = 0; l < 2; ++l)
                    if (mine[2 * h + l]) {
                        const unsigned c = own + 2 * h + l;
                        P.rho[c] = Ops<P2>::get(mac.rho, l);
                        P.u_dst[c] = Ops<P2>::get(mac.ux, l);
                        *plane_of(P.u_dst + c, vol, 1) = Ops<P2>::get(mac.uy, l);
                        *plane_of(P.u_dst + c, vol, 2) = Ops<P2>::get(mac.uz, l);
                    }
            }
        }
    }

    // (4) write-back
    if constexpr (COLLIDE) {
        if (all_mine) {
            float *pd = P.dst + own;
            static_for<0, Q>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                __stcs(reinterpret_cast<float4 *>(plane_of(pd, vol, q)), lds128(s_own + q * ROWB));
            });
        }
        // the tile's links: wall links (halfway bounce-back, write side) and the cells of quads a chord ends in
        if (n_links) {                                                   // warp-uniform
            __syncwarp();
            const unsigned s_row = (unsigned)__cvta_generic_to_shared(&stage[wib][0][4]);
            for (unsigned i = lane; i < n_links; i += 32u) {
                const unsigned L = __ldg(P.links + e.w + i);
                const float v = lds32(s_row + ((L >> 7) & 31u) * ROWB + (((L & 31u) << 2) + ((L >> 5) & 3u)) * 4u);
                const unsigned cyl = (L >> 17) & 3u, czl = (L >> 19) & 3u;
                const unsigned t = row0 + (L >> 21) + (unsigned)(cyl == 0 ? dym : (cyl == 2 ? dyq : 0)) + (unsigned)(czl == 0 ? dzm : (czl == 2 ? dzq : 0));
                *plane_of(P.dst + t, vol, (int)((L >> 12) & 31u)) = v;
            }
        }
    }
}

#endif  // LBM_EMULATE_ON_HOST
}  // namespace lbm
