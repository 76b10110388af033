// compat = physical behind walls, four cells per thread on a packed list of active quads, populations staged in shared memory
// (sm_100a) -- the default walls kernel.
//
// Why (ncu, V60 512^3, B200).  The two-cell register-staged kernel (lbm_phys.cuh) ran at 57 % of DRAM peak with 927
// warp-instructions per 64 cells; a first four-cell version that kept the dense kernel's structure (19 x LDG.128 into
// registers, shuffles) did no better: 1637 instructions per 128-cell tile of which 508 were arithmetic -- ~400 register moves
// (shifting / packing cell pairs for the f32x2 collision, unpacking for 128-bit stores), ~250 64-bit address updates, 58
// spill instructions -- and 19 % of all stall samples sat on two spill STOREs inside the load phase: a loaded value has to
// arrive before it can be spilled, so the warp's loads went out in two or three serialised round trips
// (profiles/r02_chord_v1_*).  This version takes the populations out of the register file while they are in flight:
//   * the work list is the sequence of ALL active quads (4 cells, 16-byte aligned, at least one fluid) of a plane in memory
//     order, cut into tiles of 32: one warp per tile, one lane per quad, tiles run across row ends (lbm_aux.cu,
//     build_chord_lists).  A warp costs the same whether 11 or 32 of its lanes work: tiles cut per chord left 18 % of the lanes
//     of a V60 512^3 step idle (0.67 of the HBM peak against 0.86 on a box where every lane is alive); the packed list launches
//     12.05 M lane slots for 12.05 M active quads.  Each lane reads its own (quad, y, z) entry;
//   * every lane issues 19 cp.async of 16 bytes (global -> shared, no register, L1 bypassed) into the warp's private
//     staging rows, one row of 4 + 128 + 4 floats per population: lane l owns words [4 + 4l, 8 + 4l).  The one-cell shift in
//     x of the 10 moving populations is an ADDRESS offset when the row is read back (cells x-1 .. x+2 are words 3 + 4l ..),
//     so there are no shuffles and no register moves; where the neighbouring lane does not hold the neighbouring quad
//     (chord ends, row changes inside a tile, row ends, periodic wrap) the lane fetches the one word itself by a 4-byte
//     cp.async into a small per-warp edge array;
//   * the two cell pairs of a thread are collided one after the other (packed f32x2); each pair reads its 19 inputs from the
//     row (LDS.64, or two LDS.32 for the shifted ones) directly into register pairs and writes its results back into the
//     lane's own words (STS.64), so only one pair is in registers at a time;
//   * write-back: LDS.128 + 128-bit streaming store per population for all-fluid quads; the fluid cells of a quad a chord
//     ends in go out one cell at a time with lane q < 19 storing population q (warp-uniform loop over a ballot).  Halfway
//     bounce-back stays on the write side (post-collision f_q of a fluid cell -> slot opp(q) of its solid neighbour,
//     lbm_phys.cuh) but is a precomputed WALL LINK list of the tile (one u64 per (cell, q)): all 32 lanes walk it together,
//     one link per lane and round (LDS + one 4-byte store), the first round prefetched with the tile.  The neighbour masks
//     are not read by this kernel, the flag word only for the solid / filter / LES bits;
//   * tried on top of this and rejected, with numbers (profiles/r02_tune_chord_pipelined.log, r02_exp_l2_prefetch.log,
//     r02_exp_entry_load.log, r02_exp_bulk_store.log, r02_exp_even_quad_alignment.log; V60 512^3, same box): a persistent launch
//     with two stages per warp and the next tile's loads in flight during the collision (8 warps per SM 1.92 ms, 10 warps
//     2.20 ms, against 1.87 ms for one tile per warp at 16 warps: half the warps do hide the memory latency, but then the ~6
//     cycles between two issues of one warp are the limit); prefetching the inputs of the tile 256 .. 16384 places ahead into
//     L2 (1.78 .. 2.19 ms against 1.79 ms); a persisting L2 window over the tile list and tile coordinates computed instead
//     of loaded (no change); write-back through cp.async.bulk (each UBLKCP drags ~19 uniform-datapath instructions: +1 %);
//     tiles starting on 32-byte boundaries (+2 %);
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
constexpr unsigned CHORD_ROWB = CHORD_ROW * 4;        // bytes per row
constexpr int CHORD_STAGE = Q * CHORD_ROW;            // floats per stage (one tile): 10336 B

// what a lane knows about its quad from its list entry (lbm_aux.cu: bits 0-11 quad, 12-27 y, 28-43 z, 44 live, 45 / 46 the
// neighbouring lane holds the neighbouring quad)
struct ChordGeom {
    bool live, left_adj, right_adj;
    int x0, y, z;
    unsigned own, row0;
    int dym, dyq, dzm, dzq;          // neighbour rows as 32-bit index deltas: periodic wrap, else clamp (a clamped source lies
};                                   // outside an open face and is replaced by w_q)
__device__ __forceinline__ ChordGeom chord_decode(const unsigned long long e, const Grid &G) {
    ChordGeom t;
    t.live = ((e >> 44) & 1ull) != 0; t.left_adj = ((e >> 45) & 1ull) != 0; t.right_adj = ((e >> 46) & 1ull) != 0;
    t.x0 = (int)(e & 0xfffull) * 4;         // dead lanes (padding of a plane's last tile) sit on cell (0, 0, 0): loaded, never stored
    t.y = (int)((e >> 12) & 0xffffull); t.z = (int)((e >> 28) & 0xffffull);
    t.row0 = ((unsigned)(t.z + G.zg) * (unsigned)G.ny + (unsigned)t.y) * (unsigned)G.nx;
    t.own = t.row0 + (unsigned)t.x0;
    const int nxi = G.nx, plane = (int)G.plane;
    t.dym = -nxi; if (t.y == 0) t.dym = G.per_y ? (G.ny - 1) * nxi : 0;
    t.dyq = nxi; if (t.y == G.ny - 1) t.dyq = G.per_y ? -(G.ny - 1) * nxi : 0;
    t.dzm = -plane; t.dzq = plane;
    if (!G.zg) {
        if (t.z == 0) t.dzm = G.per_z ? (G.nz - 1) * plane : 0;
        if (t.z == G.nz - 1) t.dzq = G.per_z ? -(G.nz - 1) * plane : 0;
    }
    return t;
}
// index of a moving population in the per-warp edge array: cx > 0 (1, 7, 9, 11, 13) -> 0..4, cx < 0 (2, 8, 10, 12, 14) -> 5..9
__host__ __device__ constexpr int edge_slot(int q) { return q == 1 ? 0 : q == 2 ? 5 : (cx(q) > 0 ? (q - 7) / 2 + 1 : (q - 8) / 2 + 6); }
constexpr unsigned CHORD_EDGEB = 32 * 4;              // bytes per population in the edge array

// (1) populations: global -> shared, nothing held in registers while in flight.  s_own = shared-window address of this lane's
// words of row 0 of the stage, s_edge = of this lane's word of population slot 0 in the edge array.
__device__ __forceinline__ void chord_issue_loads(const StepArgs &P, const ChordGeom &t, const unsigned s_own, const unsigned s_edge) {
    const Grid &G = P.g;
    const unsigned vol = (unsigned)G.vol;
    const float *rowp[3][3];
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
            rowp[dz + 1][dy + 1] = P.src + (t.own + (unsigned)(dy < 0 ? t.dym : (dy > 0 ? t.dyq : 0)) + (unsigned)(dz < 0 ? t.dzm : (dz > 0 ? t.dzq : 0)));
    static_for<0, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        cp_async16(s_own + q * CHORD_ROWB, plane_of(rowp[1 - cz(q)][1 - cy(q)], vol, q));
    });
    // x-1 / x+4 neighbour of the quad when the neighbouring lane does not bring it
    if (!t.left_adj) {
        int dxm = -1; if (t.x0 == 0) dxm = G.per_x ? G.nx - 1 : 0;
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if constexpr (cx(q) > 0) cp_async4(s_edge + edge_slot(q) * CHORD_EDGEB, plane_of(rowp[1 - cz(q)][1 - cy(q)], vol, q) + dxm);
        });
    }
    if (!t.right_adj) {
        int dxq = 4; if (t.x0 == G.nx - 4) dxq = G.per_x ? -(G.nx - 4) : 3;
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if constexpr (cx(q) < 0) cp_async4(s_edge + edge_slot(q) * CHORD_EDGEB, plane_of(rowp[1 - cz(q)][1 - cy(q)], vol, q) + dxq);
        });
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// (2) flags, body force, phase, rho stencil of the fused drive, first round of wall links: plain loads into registers
struct ChordAux {
    unsigned flag_word;
    unsigned long long link0;
    float4 bf[3], ph;
    float4 r, rym, ryp, rzm, rzp;
    float rxm, rxp;
};
template <bool FORCED, bool DRIVE>
__device__ __forceinline__ void chord_load_aux(const StepArgs &P, const ChordGeom &t, const uint2 tl, const unsigned lane, ChordAux &a) {
    const Grid &G = P.g;
    const unsigned vol = (unsigned)G.vol;
    a.flag_word = __ldg(reinterpret_cast<const unsigned *>(P.flags + t.own));
    a.link0 = 0;
    if (lane < tl.y) a.link0 = __ldg(P.links + tl.x + lane);
    if constexpr (FORCED) {
        if (P.force != nullptr) {
#pragma unroll
            for (int d = 0; d < 3; ++d) a.bf[d] = __ldg(reinterpret_cast<const float4 *>(plane_of(P.force + t.own, vol, d)));
        }
        if (P.phase != nullptr) a.ph = __ldg(reinterpret_cast<const float4 *>(P.phase + t.own));
    }
    if constexpr (DRIVE) {
        const float *pr = P.rho_src + t.own;
        a.r = __ldg(reinterpret_cast<const float4 *>(pr));
        a.rym = __ldg(reinterpret_cast<const float4 *>(pr + t.dym)); a.ryp = __ldg(reinterpret_cast<const float4 *>(pr + t.dyq));
        a.rzm = __ldg(reinterpret_cast<const float4 *>(pr + t.dzm)); a.rzp = __ldg(reinterpret_cast<const float4 *>(pr + t.dzq));
        a.rxm = __ldg(pr - (t.x0 > 0 ? 1 : 0)); a.rxp = __ldg(pr + (t.x0 + 4 < G.nx ? 4 : 3));
    }
}

// body force of the lane's four cells: the fused pressure-gradient drive (from the previous step's rho) + the body_force field.
// Runs BEFORE the populations have landed, i.e. inside the memory wait.
template <bool FORCED, bool DRIVE>
__device__ __forceinline__ void chord_force(const StepArgs &P, const ChordGeom &t, const ChordAux &a, float (&F)[3][4], float (&ph)[4]) {
    const Grid &G = P.g;
#pragma unroll
    for (int c = 0; c < 4; ++c) { F[0][c] = F[1][c] = F[2][c] = 0.0f; ph[c] = 0.0f; }
    if constexpr (DRIVE) {
        const float r0[4] = {a.r.x, a.r.y, a.r.z, a.r.w}, lo[4] = {a.rxm, a.r.x, a.r.y, a.r.z}, hi[4] = {a.r.y, a.r.z, a.r.w, a.rxp};
        const float ym[4] = {a.rym.x, a.rym.y, a.rym.z, a.rym.w}, yq[4] = {a.ryp.x, a.ryp.y, a.ryp.z, a.ryp.w};
        const float zm[4] = {a.rzm.x, a.rzm.y, a.rzm.z, a.rzm.w}, zq[4] = {a.rzp.x, a.rzp.y, a.rzp.z, a.rzp.w};
        const int kg = G.z0 + t.z;
        const int ypos = t.y == 0 ? -1 : (t.y == G.ny - 1 ? 1 : 0), zpos = kg == 0 ? -1 : (kg == G.nz_global - 1 ? 1 : 0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int x = t.x0 + c;
            const float gx = pressure_gradient_diff(r0[c], lo[c], hi[c], x == 0 ? -1 : (x == G.nx - 1 ? 1 : 0));
            const float gy = pressure_gradient_diff(r0[c], ym[c], yq[c], ypos);
            const float gz = pressure_gradient_diff(r0[c], zm[c], zq[c], zpos);
            const float3 f = pressure_gradient_force(r0[c], gx, gy, gz, P.drive_max_force, P.drive_scale);
            F[0][c] = f.x; F[1][c] = f.y; F[2][c] = f.z;
        }
    }
    if constexpr (FORCED) {
        if (P.force != nullptr) {      // body_force (+ drive: the sum the producer leaves in body_force in accumulate mode)
            const float b[3][4] = {{a.bf[0].x, a.bf[0].y, a.bf[0].z, a.bf[0].w}, {a.bf[1].x, a.bf[1].y, a.bf[1].z, a.bf[1].w},
                                   {a.bf[2].x, a.bf[2].y, a.bf[2].z, a.bf[2].w}};
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int c = 0; c < 4; ++c) F[d][c] = DRIVE ? b[d][c] + F[d][c] : b[d][c];
        }
        if (P.phase != nullptr) { ph[0] = a.ph.x; ph[1] = a.ph.y; ph[2] = a.ph.z; ph[3] = a.ph.w; }
    }
}

// (3) + (4): collide the tile staged at s_own (after its cp.async group has landed and the warp has synchronised), write back.
// s_row0 = shared-window address of word 0 of lane 0 in row 0 of the stage, s_park = of this lane's 8 bytes in the warp's
// 4 x 256 B parking area for the first pair's rho, u.
template <bool FORCED, bool LES, bool POROUS, bool DRIVE, bool COLLIDE>
__device__ __forceinline__ void chord_compute_store(const StepArgs &P, const ChordGeom &t, const uint2 tl, const ChordAux &a, const float (&F)[3][4],
                                                    const float (&ph)[4], const unsigned lane, const unsigned s_own, const unsigned s_row0,
                                                    const unsigned s_edge, const unsigned s_park) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool HAS_F = FORCED || DRIVE;
    constexpr unsigned ROWB = CHORD_ROWB;
    const Grid &G = P.g;
    const unsigned vol = (unsigned)G.vol, own = t.own, n_links = tl.y;
    const int x0 = t.x0, y = t.y, z = t.z;
    const bool live = t.live;
    const bool has_phase = FORCED && P.phase != nullptr;
    const bool has_force = DRIVE || (FORCED && (P.force != nullptr || (has_phase && P.gravity_lu != 0.0f)));
    unsigned fl[4], mine_bits = 0;
    bool mine[4], all_mine = true;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        fl[c] = (a.flag_word >> (8 * c)) & 0xffu;
        mine[c] = live && !(fl[c] & LBM_FLAG_SOLID);
        all_mine &= mine[c];
        mine_bits |= mine[c] ? (1u << c) : 0u;
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

    // The two cell pairs, one after the other.  Pair h = cells 2h, 2h + 1 of the quad; population q of those cells:
    //   cx = 0: words 2h, 2h + 1 of the lane; cx > 0 (source x - 1): words 2h - 1, 2h; cx < 0 (source x + 1): words 2h + 1, 2h + 2.
    // Before pair 0 writes its results into words 0, 1, every word of pair 1 that a result could overwrite is taken:
    // the lane's own word 1 (cx > 0) and the next lane's word 0 (cx < 0; from the edge array where that lane holds another quad).
    float keep[Q];
    static_for<0, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        if constexpr (cx(q) > 0) keep[q] = lds32(s_own + q * ROWB + 4);
        if constexpr (cx(q) < 0) keep[q] = lds32(t.right_adj ? s_own + q * ROWB + 16 : s_edge + edge_slot(q) * CHORD_EDGEB);
    });
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        P2 fp[Q];
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            const unsigned sa = s_own + q * ROWB + 8 * h;
            if constexpr (cx(q) == 0) fp[q] = lds64(sa);
            else if constexpr (cx(q) > 0)
                fp[q] = h == 0 ? p2_make(lds32(t.left_adj ? sa - 4 : s_edge + edge_slot(q) * CHORD_EDGEB), lds32(sa)) : p2_make(keep[q], lds32(sa));
            else fp[q] = h == 0 ? p2_make(lds32(sa + 4), lds32(sa + 8)) : p2_make(lds32(sa + 4), keep[q]);
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
        // rho, u: all-fluid quads go out as 128-bit vectors after the second pair (a sector written in two halves costs a fill
        // from HBM: measured +0.23 ms per step at V60 512^3); the first pair's four values wait in shared memory, not in registers
        if (P.write_macro) {
            if (all_mine) {
                if (h == 0) {
                    sts64(s_park, mac.rho); sts64(s_park + 256u, mac.ux); sts64(s_park + 512u, mac.uy); sts64(s_park + 768u, mac.uz);
                } else {
                    const P2 r0 = lds64(s_park), x0p = lds64(s_park + 256u), y0p = lds64(s_park + 512u), z0p = lds64(s_park + 768u);
                    float *pu = P.u_dst + own;
                    __stcs(reinterpret_cast<float4 *>(P.rho + own), make_float4(p2_lo(r0), p2_hi(r0), p2_lo(mac.rho), p2_hi(mac.rho)));
                    __stcs(reinterpret_cast<float4 *>(pu), make_float4(p2_lo(x0p), p2_hi(x0p), p2_lo(mac.ux), p2_hi(mac.ux)));
                    __stcs(reinterpret_cast<float4 *>(plane_of(pu, vol, 1)), make_float4(p2_lo(y0p), p2_hi(y0p), p2_lo(mac.uy), p2_hi(mac.uy)));
                    __stcs(reinterpret_cast<float4 *>(plane_of(pu, vol, 2)), make_float4(p2_lo(z0p), p2_hi(z0p), p2_lo(mac.uz), p2_hi(mac.uz)));
                }
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

    // write-back
    if constexpr (COLLIDE) {
        if (all_mine) {
            float *pd = P.dst + own;
            static_for<0, Q>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                __stcs(reinterpret_cast<float4 *>(plane_of(pd, vol, q)), lds128(s_own + q * ROWB));
            });
        }
        // quads a chord ends in (fluid and solid cells): their fluid cells go out one cell at a time, lane q < 19 storing population q
        const unsigned mixed = __ballot_sync(FULL, live && !all_mine);
        if (mixed | n_links) __syncwarp();                               // results of every lane are in the stage
        for (unsigned m = mixed; m; m &= m - 1) {                        // warp-uniform
            const int ls = __ffs(m) - 1;
            const unsigned bits = __shfl_sync(FULL, mine_bits, ls), cell0 = __shfl_sync(FULL, own, ls);
            if (lane < Q) {
                float *pq = plane_of(P.dst + cell0, vol, (int)lane);
                const unsigned sa = s_row0 + lane * ROWB + (unsigned)ls * 16u;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((bits >> c) & 1u) pq[c] = lds32(sa + 4u * c);
            }
        }
        // wall links (halfway bounce-back, write side): one link per lane and round, the first round was loaded with the tile
        for (unsigned i = lane; i < n_links; i += 32u) {
            const unsigned long long L = i < 32u ? a.link0 : __ldg(P.links + tl.x + i);
            const unsigned hi = (unsigned)(L >> 32);
            const float v = lds32(s_row0 + ((hi >> 7) & 31u) * ROWB + (((hi & 31u) << 2) + ((hi >> 5) & 3u)) * 4u);
            *plane_of(P.dst + (unsigned)L, vol, (int)((hi >> 12) & 31u)) = v;
        }
    }
}

// One tile (32 consecutive entries of the packed quad list) per warp.
template <bool FORCED, bool LES, bool POROUS, bool DRIVE, int BLOCK, bool COLLIDE, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) phys_chord_kernel(const __grid_constant__ StepArgs P) {
    __shared__ __align__(16) float stage[BLOCK / 32][CHORD_STAGE];
    __shared__ float edge[BLOCK / 32][10][32];
    __shared__ __align__(8) float park[BLOCK / 32][4][64];
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const int w = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
    if (w >= P.n_items) return;                                          // warp-uniform
    const size_t tile = (size_t)P.item_begin + (size_t)w;
    const unsigned long long e = __ldg(P.quads + tile * 32 + lane);
    const uint2 tl = __ldg(P.tile_links + tile);
    const ChordGeom t = chord_decode(e, P.g);
    const unsigned s_row0 = (unsigned)__cvta_generic_to_shared(&stage[wib][4]);
    const unsigned s_own = s_row0 + 16u * lane;
    const unsigned s_edge = (unsigned)__cvta_generic_to_shared(&edge[wib][0][lane]);
    chord_issue_loads(P, t, s_own, s_edge);
    ChordAux a{};
    chord_load_aux<FORCED, DRIVE>(P, t, tl, lane, a);
    float F[3][4], ph[4];
    chord_force<FORCED, DRIVE>(P, t, a, F, ph);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    chord_compute_store<FORCED, LES, POROUS, DRIVE, COLLIDE>(P, t, tl, a, F, ph, lane, s_own, s_row0, s_edge,
                                                             (unsigned)__cvta_generic_to_shared(&park[wib][0][2 * lane]));
}

#endif  // LBM_EMULATE_ON_HOST
}  // namespace lbm
