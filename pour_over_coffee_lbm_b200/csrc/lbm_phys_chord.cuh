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
// 16 bytes, or 16 bytes of zeros without touching global memory when bytes == 0
__device__ __forceinline__ void cp_async16_or_zero(unsigned smem_dst, const void *gsrc, unsigned bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(bytes) : "memory");
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

// Loads whose ISSUE ORDER matters are volatile asm: ptxas otherwise sinks a load below the first use of an unrelated value that
// happens to share its scoreboard (the list entry of the tile's links waited for the rho stencil: one serialised DRAM round trip
// per tile, 13 % of all stall samples of the headline kernel -- profiles/r02_headline_*; found with scripts/sass_ctrl.py).
#ifndef LBM_CHORD_ORDERED
#define LBM_CHORD_ORDERED 1
#endif
#ifndef LBM_CHORD_PREFETCH
#define LBM_CHORD_PREFETCH 1024
#endif
#if LBM_CHORD_ORDERED
__device__ __forceinline__ unsigned long long ldo_u64(const void *p) { unsigned long long v; asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint2 ldo_u32x2(const void *p) { uint2 v; asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ unsigned ldo_u32(const void *p) { unsigned v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ float ldo_f32(const void *p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ float4 ldo_f32x4(const void *p) {
    float4 v; asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v;
}
#else
__device__ __forceinline__ unsigned long long ldo_u64(const void *p) { return __ldg(reinterpret_cast<const unsigned long long *>(p)); }
__device__ __forceinline__ uint2 ldo_u32x2(const void *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ unsigned ldo_u32(const void *p) { return __ldg(reinterpret_cast<const unsigned *>(p)); }
__device__ __forceinline__ float ldo_f32(const void *p) { return __ldg(reinterpret_cast<const float *>(p)); }
__device__ __forceinline__ float4 ldo_f32x4(const void *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
#endif

// One staged population row: the 32 lanes' 4 words, then one EDGE LINE of 32 words (one per lane).  The edge line of a population
// that moves in x holds the word the neighbouring lane does not bring (fetched by the lane itself), later the word the lane's second
// pair needs and its first pair overwrites; the edge lines of eight populations that do not move in x park the first pair's rho, u.
constexpr unsigned CHORD_ROWB = 512 + 128;            // bytes per row
constexpr unsigned CHORD_EDGE = 512;                  // byte offset of the edge line inside its row
constexpr int CHORD_STAGE = Q * (int)(CHORD_ROWB / 4);   // floats per stage (one tile): 12160 B

// what a lane knows about its quad from its list entry (lbm_aux.cu: bits 0-11 quad, 12-27 y, 28-43 z, 44 live, 45 / 46 the
// neighbouring lane holds the neighbouring quad)
struct ChordGeom {
    bool live, left_adj, right_adj;
    unsigned dead_rows;              // bit per neighbouring row: its quad at this x brings nothing the collision uses (lbm_aux.cu)
    int x0, y, z;
    unsigned own, row0;
    int dym, dyq, dzm, dzq;          // neighbour rows as 32-bit index deltas: periodic wrap, else clamp (a clamped source lies
};                                   // outside an open face and is replaced by w_q)
__device__ __forceinline__ ChordGeom chord_decode(const unsigned long long e, const Grid &G) {
    ChordGeom t;
    t.live = ((e >> 44) & 1ull) != 0; t.left_adj = ((e >> 45) & 1ull) != 0; t.right_adj = ((e >> 46) & 1ull) != 0;
    t.dead_rows = (unsigned)(e >> 47) & 0xffu;
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
// (1) populations: global -> shared, nothing held in registers while in flight.  s_own = shared-window address of this lane's
// four words of row 0 of the stage, s_edge = of this lane's word in the edge line of row 0.
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
        if constexpr (cy(q) == 0 && cz(q) == 0) cp_async16(s_own + q * CHORD_ROWB, plane_of(rowp[1][1], vol, q));
        else {      // source row (y - cy, z - cz)
            constexpr int b9 = (1 - cz(q)) * 3 + (1 - cy(q)), bit = b9 > 4 ? b9 - 1 : b9;
            cp_async16_or_zero(s_own + q * CHORD_ROWB, plane_of(rowp[1 - cz(q)][1 - cy(q)], vol, q), ((t.dead_rows >> bit) & 1u) ? 0u : 16u);
        }
    });
    // x-1 / x+4 neighbour of the quad when the neighbouring lane does not bring it
    if (!t.left_adj) {
        int dxm = -1; if (t.x0 == 0) dxm = G.per_x ? G.nx - 1 : 0;
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if constexpr (cx(q) > 0) cp_async4(s_edge + q * CHORD_ROWB, plane_of(rowp[1 - cz(q)][1 - cy(q)], vol, q) + dxm);
        });
    }
    if (!t.right_adj) {
        int dxq = 4; if (t.x0 == G.nx - 4) dxq = G.per_x ? -(G.nx - 4) : 3;
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if constexpr (cx(q) < 0) cp_async4(s_edge + q * CHORD_ROWB, plane_of(rowp[1 - cz(q)][1 - cy(q)], vol, q) + dxq);
        });
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// (2) flags, body force, phase, rho stencil of the fused drive, first round of wall links: plain loads into registers
struct ChordAux {
    unsigned flag_word;
    unsigned link0;                  // first round of the tile's wall links: one link and the value waiting on it per lane
    float wall0;
    float4 bf[3], ph;
    float4 r, rym, ryp, rzm, rzp;
    float rxm, rxp;
};
template <bool FORCED, bool DRIVE>
__device__ __forceinline__ void chord_load_aux(const StepArgs &P, const ChordGeom &t, const uint2 tl, const unsigned lane, ChordAux &a) {
    const Grid &G = P.g;
    const unsigned vol = (unsigned)G.vol;
    a.link0 = 0; a.wall0 = 0.0f;
    if (lane < tl.y) { a.link0 = ldo_u32(P.links + tl.x + lane); a.wall0 = ldo_f32(P.wall + tl.x + lane); }
    a.flag_word = ldo_u32(P.flags + t.own);
    if constexpr (DRIVE) {
        const float *pr = P.rho_src + t.own;
        a.r = ldo_f32x4(pr);
        a.rym = ldo_f32x4(pr + t.dym); a.ryp = ldo_f32x4(pr + t.dyq);
        a.rzm = ldo_f32x4(pr + t.dzm); a.rzp = ldo_f32x4(pr + t.dzq);
        a.rxm = ldo_f32(pr - (t.x0 > 0 ? 1 : 0)); a.rxp = ldo_f32(pr + (t.x0 + 4 < G.nx ? 4 : 3));
    }
    if constexpr (FORCED) {
        if (P.force != nullptr) {
#pragma unroll
            for (int d = 0; d < 3; ++d) a.bf[d] = ldo_f32x4(plane_of(P.force + t.own, vol, d));
        }
        if (P.phase != nullptr) a.ph = ldo_f32x4(P.phase + t.own);
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

// Open faces: a source outside the box delivers w_q.  Written into the STAGE before the collision reads it --
// the lane's four words and its edge word of every population whose source row lies outside (all readers of those words sit in the
// same row), the edge word alone for the first / last quad of a row.  Out of line: few tiles touch a face, and the hot loop stays
// inside the 32 KB instruction cache (B300_MICROARCH.md: L1.5 I-cache 32 KB; the kernel was 41-44 KB, ncu: 61 % of the GPC
// instruction-fetch peak, 1.5 of 14 stall cycles per issue on `no_instruction`).
static __device__ __noinline__ void chord_open_faces(const unsigned s_own, const unsigned s_edge, const unsigned faces) {
    const bool ylo = faces & 1u, yhi = faces & 2u, zlo = faces & 4u, zhi = faces & 8u, xlo = faces & 16u, xhi = faces & 32u;
    static_for<1, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        constexpr float w = wq(q);
        const bool row_out = (cy(q) > 0 && ylo) || (cy(q) < 0 && yhi) || (cz(q) > 0 && zlo) || (cz(q) < 0 && zhi);
        if (row_out) asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(s_own + q * CHORD_ROWB), "f"(w) : "memory");
        if constexpr (cx(q) != 0) {
            if (row_out || (cx(q) > 0 ? xlo : xhi)) asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_edge + q * CHORD_ROWB), "f"(w) : "memory");
        }
    });
}

// (3) + (4): collide the tile staged at s_own (after its cp.async group has landed and the warp has synchronised), write back.
// s_row0 = shared-window address of lane 0's words in row 0 of the stage.
template <bool FORCED, bool LES, bool POROUS, bool DRIVE, bool COLLIDE, bool MRT>
__device__ __forceinline__ void chord_compute_store(const StepArgs &P, const ChordGeom &t, const uint2 tl, const ChordAux &a, const float (&F)[3][4],
                                                    const float (&ph)[4], const unsigned lane, const unsigned s_own, const unsigned s_row0,
                                                    const unsigned s_edge) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool HAS_F = FORCED || DRIVE;
    constexpr unsigned ROWB = CHORD_ROWB;
    const Grid &G = P.g;
    const unsigned vol = (unsigned)G.vol, own = t.own, n_links = tl.y;
    const bool live = t.live;
    const bool has_phase = FORCED && P.phase != nullptr;
    const bool has_force = DRIVE || (FORCED && (P.force != nullptr || (has_phase && P.gravity_lu != 0.0f)));
    unsigned mine_bits = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (live && !((a.flag_word >> (8 * c)) & LBM_FLAG_SOLID)) mine_bits |= 1u << c;
    const bool all_mine = mine_bits == 15u;
    // Halfway bounce-back: the value a cell sent towards a solid neighbour one step ago comes back as the opposite population.  It
    // waited in the per-link buffer and goes into the stage word the pull would have read from the solid cell (lbm_aux.cu).
#pragma unroll 1
    for (unsigned i = lane; i < n_links; i += 32u) {
        const unsigned L = i < 32u ? a.link0 : __ldg(P.links + tl.x + i);
        const float v = i < 32u ? a.wall0 : __ldg(P.wall + tl.x + i);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_row0 + 4u * (L & 0xfffu)), "f"(v) : "memory");
    }
    if (!(G.per_x && G.per_y && G.per_z)) {
        const int zglob = G.z0 + t.z;
        unsigned faces = 0;
        if (!G.per_y) faces |= (t.y == 0 ? 1u : 0u) | (t.y == G.ny - 1 ? 2u : 0u);
        if (!G.per_z) faces |= (zglob == 0 ? 4u : 0u) | (zglob == G.nz_global - 1 ? 8u : 0u);
        if (!G.per_x) faces |= (t.x0 == 0 ? 16u : 0u) | (t.x0 == G.nx - 4 ? 32u : 0u);
        if (faces) chord_open_faces(s_own, s_edge, faces);
    }
    __syncwarp();
    // Pair h = cells 2h, 2h + 1 of the quad; population q of those cells:
    //   cx = 0: words 2h, 2h + 1 of the lane; cx > 0 (source x - 1): words 2h - 1, 2h; cx < 0 (source x + 1): words 2h + 1, 2h + 2.
    // Word -1 is the previous lane's word 3 or the lane's edge word, word 4 the next lane's word 0 or the edge word.  Pair 0 writes its
    // results into words 0, 1 of every lane, so the two words of pair 1 that this would overwrite move into the edge line first: the next
    // lane's word 0 here, the lane's own word 1 once pair 0 has read the edge word.  The loop over the pairs is NOT unrolled (one copy of
    // the collision in the instruction cache); what differs between the pairs is three base addresses.
    if (t.right_adj) {                                     // five loads, then five stores: interleaved, every store waits for its load
        float w0[Q];
        static_for<0, Q>([&](auto qq) { constexpr int q = decltype(qq)::value; if constexpr (cx(q) < 0) w0[q] = lds32(s_own + q * ROWB + 16); });
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if constexpr (cx(q) < 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_edge + q * ROWB), "f"(w0[q]) : "memory");
        });
    }
    unsigned sa = s_own;                                   // the pair's own two words
    unsigned sl = t.left_adj ? s_own - 4 : s_edge;         // cx > 0: its first source word
    unsigned sr = s_own + 8;                               // cx < 0: its second source word
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        P2 fp[Q];
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if constexpr (cx(q) == 0) fp[q] = lds64(sa + q * ROWB);
            else if constexpr (cx(q) > 0) fp[q] = p2_make(lds32(sl + q * ROWB), lds32(sa + q * ROWB));
            else fp[q] = p2_make(lds32(sa + q * ROWB + 4), lds32(sr + q * ROWB));
        });
        if (h == 0) {
            float w1[Q];
            static_for<0, Q>([&](auto qq) { constexpr int q = decltype(qq)::value; if constexpr (cx(q) > 0) w1[q] = lds32(s_own + q * ROWB + 4); });
            static_for<0, Q>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                if constexpr (cx(q) > 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_edge + q * ROWB), "f"(w1[q]) : "memory");
            });
        }
        CellIn<P2> in;
        in.Fx = h ? p2_make(F[0][2], F[0][3]) : p2_make(F[0][0], F[0][1]);
        in.Fy = h ? p2_make(F[1][2], F[1][3]) : p2_make(F[1][0], F[1][1]);
        in.Fz = h ? p2_make(F[2][2], F[2][3]) : p2_make(F[2][0], F[2][1]);
        in.phase = h ? p2_make(ph[2], ph[3]) : p2_make(ph[0], ph[1]);
        // dead lanes (the padding of a plane's last tile) count as solid
        const unsigned fw = live ? a.flag_word >> (16 * h) : 0x0101u;
        in.flag[0] = fw & 0xffu; in.flag[1] = (fw >> 8) & 0xffu;
        CellMacro<P2> mac;
#ifdef LBM_EXP_COPY      /* timing experiment: no collision (wrong results) */
        mac.rho = mac.ux = mac.uy = mac.uz = fp[0];
#else
        collide_phys<P2, HAS_F, LES, POROUS, COLLIDE, true, MRT>(fp, in, mac, P, has_phase, has_force);
#endif
        if constexpr (COLLIDE) {
            __syncwarp();                                                // every lane has read this pair's inputs
            if (live) {
                static_for<0, Q>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    sts64(sa + q * ROWB, fp[q]);
                });
            }
        }
        // rho, u: all-fluid quads go out as 128-bit vectors after the second pair (a sector written in two halves costs a fill
        // from HBM: measured +0.23 ms per step at V60 512^3); the first pair's four values wait in the edge lines of rows 3..6, 15..18
        if (P.write_macro) {
            if (all_mine) {
                const unsigned s_park = s_row0 + CHORD_EDGE + 8u * lane + (lane < 16u ? 0u : ROWB - 128u);
                if (h == 0) {
                    sts64(s_park + 3u * ROWB, mac.rho); sts64(s_park + 5u * ROWB, mac.ux); sts64(s_park + 15u * ROWB, mac.uy); sts64(s_park + 17u * ROWB, mac.uz);
                } else {
                    const P2 r0 = lds64(s_park + 3u * ROWB), x0p = lds64(s_park + 5u * ROWB), y0p = lds64(s_park + 15u * ROWB), z0p = lds64(s_park + 17u * ROWB);
                    float *pu = P.u_dst + own;
                    __stcs(reinterpret_cast<float4 *>(P.rho + own), make_float4(p2_lo(r0), p2_hi(r0), p2_lo(mac.rho), p2_hi(mac.rho)));
                    __stcs(reinterpret_cast<float4 *>(pu), make_float4(p2_lo(x0p), p2_hi(x0p), p2_lo(mac.ux), p2_hi(mac.ux)));
                    __stcs(reinterpret_cast<float4 *>(plane_of(pu, vol, 1)), make_float4(p2_lo(y0p), p2_hi(y0p), p2_lo(mac.uy), p2_hi(mac.uy)));
                    __stcs(reinterpret_cast<float4 *>(plane_of(pu, vol, 2)), make_float4(p2_lo(z0p), p2_hi(z0p), p2_lo(mac.uz), p2_hi(mac.uz)));
                }
            } else {
#pragma unroll
                for (int l = 0; l < 2; ++l)
                    if ((mine_bits >> (2 * h + l)) & 1u) {
                        const unsigned c = own + 2 * h + l;
                        P.rho[c] = Ops<P2>::get(mac.rho, l);
                        P.u_dst[c] = Ops<P2>::get(mac.ux, l);
                        *plane_of(P.u_dst + c, vol, 1) = Ops<P2>::get(mac.uy, l);
                        *plane_of(P.u_dst + c, vol, 2) = Ops<P2>::get(mac.uz, l);
                    }
            }
        }
        sa += 8; sl = s_edge; sr = s_edge;
    }

    // write-back: every listed quad as one 128-bit streaming store per population (both quads of a 32-byte sector are listed, so no
    // sector leaves L2 half written), then the values the wall links wait for, taken from the stage
    if constexpr (COLLIDE) {
        if (live) {
            float *pd = P.dst + own;
            static_for<0, Q>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                __stcs(reinterpret_cast<float4 *>(plane_of(pd, vol, q)), lds128(s_own + q * ROWB));
            });
        }
        if (n_links) {                                                   // warp-uniform
            __syncwarp();                                                // results of every lane are in the stage
#pragma unroll 1
            for (unsigned i = lane; i < n_links; i += 32u) {
                const unsigned L = i < 32u ? a.link0 : __ldg(P.links + tl.x + i);
                P.wall[tl.x + i] = lds32(s_row0 + 4u * ((L >> 12) & 0xfffu));
            }
        }
    }
}

// One tile (32 consecutive entries of the packed quad list) per warp.  MRT = true: the instantiation for lbm_params.mrt_magic > 0 (two-rate
// collision, three more live register pairs); the BGK instantiations, which carry the roofline numbers, are built without it.
template <bool FORCED, bool LES, bool POROUS, bool DRIVE, int BLOCK, bool COLLIDE, int MINB, bool MRT = false>
__global__ void __launch_bounds__(BLOCK, MINB) phys_chord_kernel(const __grid_constant__ StepArgs P) {
    __shared__ __align__(16) float stage[BLOCK / 32][CHORD_STAGE];
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const int w = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
    if (w >= P.n_items) return;                                          // warp-uniform
    const size_t tile = (size_t)P.item_begin + (size_t)w;
    const unsigned long long e = ldo_u64(P.quads + tile * 32 + lane);
    const uint2 tl = ldo_u32x2(P.tile_links + tile);
#if LBM_CHORD_PREFETCH > 0
    // the list entries of the tile a warp of the next wave will start with: its first DRAM round trip becomes an L2 hit
    if (w + LBM_CHORD_PREFETCH < P.n_items) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.quads + (tile + LBM_CHORD_PREFETCH) * 32 + lane));
        if (lane == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tile_links + tile + LBM_CHORD_PREFETCH));
    }
#endif
    const ChordGeom t = chord_decode(e, P.g);
    const unsigned s_row0 = (unsigned)__cvta_generic_to_shared(&stage[wib][0]);
    const unsigned s_own = s_row0 + 16u * lane, s_edge = s_row0 + CHORD_EDGE + 4u * lane;
#if LBM_CHORD_ORDERED
    ChordAux a{};
    chord_load_aux<FORCED, DRIVE>(P, t, tl, lane, a);       // first: what the force computation inside the memory wait needs
    chord_issue_loads(P, t, s_own, s_edge);
#else
    chord_issue_loads(P, t, s_own, s_edge);
    ChordAux a{};
    chord_load_aux<FORCED, DRIVE>(P, t, tl, lane, a);
#endif
    float F[3][4], ph[4];
    chord_force<FORCED, DRIVE>(P, t, a, F, ph);
#if LBM_CHORD_ORDERED
    {   // Every register load above must be ISSUED before the wait below.  ptxas sinks a load towards its first use, and that lies
        // behind the wait for the inputs the collision reads first: the warp then sat through a second DRAM round trip after the
        // populations had landed (ncu: 14 % of all stall samples on the flag word).  One word that depends on every loaded register,
        // stored into an unused edge word of the stage before the wait, pins them (11 instructions per tile).
        unsigned k = a.flag_word ^ a.link0 ^ __float_as_uint(a.wall0);
        if constexpr (FORCED) {
#pragma unroll
            for (int d = 0; d < 3; ++d) k ^= __float_as_uint(a.bf[d].x) ^ __float_as_uint(a.bf[d].y) ^ __float_as_uint(a.bf[d].z) ^ __float_as_uint(a.bf[d].w);
            k ^= __float_as_uint(a.ph.x) ^ __float_as_uint(a.ph.y) ^ __float_as_uint(a.ph.z) ^ __float_as_uint(a.ph.w);
        }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_edge), "r"(k) : "memory");      // row 0 does not move in x: its edge line is free
    }
#endif
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    chord_compute_store<FORCED, LES, POROUS, DRIVE, COLLIDE, MRT>(P, t, tl, a, F, ph, lane, s_own, s_row0, s_edge);
}

#endif  // LBM_EMULATE_ON_HOST
}  // namespace lbm
