// compat = physical, walls path, TMA-staged: the kernel behind the V60 configurations on B200 (sm_100a).
//
// Why: ncu on the register-staged kernel (lbm_phys.cuh:phys_walls_kernel, 114 registers, 16 warps/SM) shows it latency
// bound, not bandwidth bound -- 59 % of DRAM peak, 32 % issue utilisation, 7 cycles of long-scoreboard stall per issue:
// every warp serialises "19 loads -> 1 us of DRAM latency -> ~800 dependent instructions -> 19 stores", and with the
// register file full of in-flight loads there are too few warps to overlap the phases.  Here the loads are taken out of
// the warps' instruction streams altogether:
//
//   * persistent CTAs (a few per SM); each walks the active-tile list with stride gridDim.x.  A tile is 64 x TY cells
//     (64 x-consecutive cells of TY consecutive rows of one plane) that contains at least one fluid cell;
//   * NP PRODUCER warps per CTA (one elected lane each) issue, per tile, 19 + 3 + 1 + 1 tensor-map TMA loads
//     (cp.async.bulk.tensor): population q as the box of rows (y0 - cy, z - cz) -- the y/z part of the pull shift is
//     done by the copy engine, OUT-OF-BOX sources arrive as zeros and are replaced by the open-face rule -- then the
//     body force, the phase field and the flag bytes.  The x part cannot be: in tiled mode the innermost start
//     coordinate must be a multiple of 16 bytes (measured on B200 with scripts/probes/tma_probe.cu: x = 1, 3, -1, 255
//     raise "illegal instruction", x = -4, 4, 224 and any y/z, negative included, are fine).  So populations with
//     cx != 0 come as 68-wide boxes ([x0-4, x0+64) for cx = +1, [x0, x0+68) for cx = -1) and the consumers read them
//     one cell off.  Everything lands in a ring of STAGES shared-memory stages, each guarded by a full / empty
//     mbarrier pair; bytes in flight are set by the ring (up to 195 KB per SM), not by registers or occupancy;
//   * TY CONSUMER warps, one row of the tile each: wait on the stage, read their 19 populations from shared memory
//     (two cells per thread, packed f32x2 arithmetic), release the stage, and run the very same collision /
//     write-back code as the register-staged kernel (phys_finish).  Results go straight to global memory with
//     streaming 64-bit stores; no CTA-wide barrier anywhere.
//
// Used when the box is not periodic in x or y (V60, bounce-back boxes; z may wrap) and nx % 16 == 0; otherwise the
// register-staged kernel runs.  Bit-exact against the oracle like every compat = physical kernel (same operator).
#pragma once
#include "lbm_phys.cuh"

namespace lbm {

constexpr int TMA_TX = 64;                           // tile width in cells (32 lanes x 2 cells)

constexpr int TMA_TXW = 68;                          // box width of the populations that move in x (one 16-byte halo)

template <int TY> struct TmaStage {
    static constexpr int ROW = TMA_TX * 4;           // bytes of one row of a 64-wide f32 box
    static constexpr int ROWW = TMA_TXW * 4;         // ... of a 68-wide one
    static constexpr int BOX = ROW * TY;             // 64-wide f32 box (populations with cx = 0, force, phase)
    static constexpr int BOXW = ROWW * TY;           // 68-wide f32 box (populations with cx != 0): bytes transferred
    static constexpr int BOXW_PAD = ((BOXW + 127) / 128) * 128;      // ... and its 128-byte aligned slot
    // byte offset of population q inside a stage
    static constexpr int off_q(int q) { int o = 0; for (int i = 0; i < q; ++i) o += cx(i) != 0 ? BOXW_PAD : BOX; return o; }
    static constexpr int OFF_FORCE = off_q(Q);
    static constexpr int OFF_PHASE = OFF_FORCE + 3 * BOX;
    static constexpr int OFF_FLAGS = OFF_PHASE + BOX;
    static constexpr int FLAG_BOX = TMA_TX * TY;     // u8
    static constexpr int BYTES = ((OFF_FLAGS + FLAG_BOX + 127) / 128) * 128;
    static constexpr int POP_TX_BYTES = 9 * BOX + 10 * BOXW;         // bytes the 19 population loads deliver
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// no L2 eviction hint: the 16-byte halo sectors of the 68-wide boxes are shared with the x-neighbour tile, which another
// CTA loads at about the same time
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap *map, unsigned bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, unsigned bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// One producer warp's share of a tile's loads: request r (0..18 populations, 19..21 force, 22 phase, 23 flags) belongs to
// producer r % NP.  A single warp needs ~1 us to issue all 24 (each UTMALDG drags ~12 dependent uniform-datapath
// instructions behind it), which starved the consumers in the first version (ncu: 17 % of all stall samples on the
// full-barrier wait); NP warps issue their shares concurrently and each arrives once on the full barrier.
template <int TY, int NP, int PIDX>
__device__ __forceinline__ void tma_produce(unsigned base, unsigned full, const TmaMaps &M, int x0, int y0, int zc, int zlo, int zhi,
                                            bool use_force, bool use_phase) {
    using S = TmaStage<TY>;
    unsigned tx = 0;
    static_for<0, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        if constexpr (q % NP == PIDX) tx += cx(q) != 0 ? S::BOXW : S::BOX;
    });
#pragma unroll
    for (int c = 0; c < 3; ++c) if ((Q + c) % NP == PIDX && use_force) tx += S::BOX;
    if ((Q + 3) % NP == PIDX && use_phase) tx += S::BOX;
    if ((Q + 4) % NP == PIDX) tx += S::FLAG_BOX;
    mbar_expect_tx(full, tx);
    static_for<0, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        if constexpr (q % NP == PIDX) {
            const int zq = cz(q) > 0 ? zlo : (cz(q) < 0 ? zhi : zc);
            if constexpr (cx(q) == 0) tma_load_4d(base + S::off_q(q), &M.pops, full, x0, y0 - cy(q), zq, q);
            else tma_load_4d(base + S::off_q(q), &M.pops_wide, full, cx(q) > 0 ? x0 - 4 : x0, y0 - cy(q), zq, q);
        }
    });
#pragma unroll
    for (int c = 0; c < 3; ++c)
        if ((Q + c) % NP == PIDX && use_force) tma_load_4d(base + S::OFF_FORCE + c * S::BOX, &M.force, full, x0, y0, zc, c);
    if ((Q + 3) % NP == PIDX && use_phase) tma_load_3d(base + S::OFF_PHASE, &M.phase, full, x0, y0, zc);
    if ((Q + 4) % NP == PIDX) tma_load_3d(base + S::OFF_FLAGS, &M.flags, full, x0, y0, zc);
}

template <bool FORCED, bool LES, bool POROUS, int TY, int STAGES, bool COLLIDE, int MINB, int NP>
__global__ void __launch_bounds__((TY + NP) * 32, MINB) phys_tma_kernel(const __grid_constant__ StepArgs P, const __grid_constant__ TmaMaps M) {
    using V = P2;
    using O = Ops<V>;
    using S = TmaStage<TY>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // dynamic shared memory is only guaranteed 16-byte aligned: round up to 128 (the host adds the slack); done on the
    // shared-window offset so that the accesses below stay LDS / STS instead of generic loads
    unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + STAGES * S::BYTES);      // full[STAGES], empty[STAGES]
    const Grid &G = P.g;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31u;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(bars + s), NP);                // each producer's arrive.expect_tx (+ the bytes)
            mbar_init(smem_u32(bars + STAGES + s), TY);       // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_my = (P.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // tiles blockIdx.x + i * gridDim.x
    const unsigned *items = P.items + P.item_begin + blockIdx.x;
    const bool use_force = FORCED && P.force != nullptr;
    const bool use_phase = FORCED && P.phase != nullptr;

    if (warp >= TY) {
        // ---------------- producers ----------------
        const int pidx = warp - TY;
        unsigned e_next = n_my > 0 ? __ldg(items) : 0u;
        for (int i = 0; i < n_my; ++i) {
            const int st = i % STAGES;
            const unsigned e = e_next;
            if (i + 1 < n_my) e_next = __ldg(items + (size_t)(i + 1) * gridDim.x);
            if (i >= STAGES) mbar_wait(smem_u32(bars + STAGES + st), (unsigned)((i / STAGES) & 1) ^ 1u);
            const int x0 = (int)(e & 0xffu) * TMA_TX, y0 = (int)((e >> 8) & 0xfffu), z = (int)(e >> 20);
            const unsigned full = smem_u32(bars + st);
            const unsigned base = smem_u32(smem + st * S::BYTES);
            // UTMALDG is a warp-uniform instruction: issued by ONE thread per producer warp
            if (lane == 0) {
                const int zc = z + G.zg;
                int zlo = z - 1, zhi = z + 1;      // single slab: wrap in z, or leave the box (zero fill -> open-face rule)
                if (!G.zg) {
                    if (zlo < 0) zlo = G.per_z ? G.nz - 1 : -1;
                    if (zhi >= G.nz) zhi = G.per_z ? 0 : G.nz;
                }
                zlo += G.zg; zhi += G.zg;
                static_for<0, NP>([&](auto pp) {
                    constexpr int PIDX = decltype(pp)::value;
                    if (pidx == PIDX) tma_produce<TY, NP, PIDX>(base, full, M, x0, y0, zc, zlo, zhi, use_force, use_phase);
                });
            }
            __syncwarp();
        }
        return;
    }

    // ---------------- consumers: warp w owns row y0 + w of every tile ----------------
    bool has_force = false, has_phase = false;
    if constexpr (FORCED) {
        has_phase = P.phase != nullptr;
        has_force = P.force != nullptr || (has_phase && P.gravity_lu != 0.0f);
    }
    unsigned e_next = n_my > 0 ? __ldg(items) : 0u;
    for (int i = 0; i < n_my; ++i) {
        const int st = i % STAGES;
        const unsigned e = e_next;
        if (i + 1 < n_my) e_next = __ldg(items + (size_t)(i + 1) * gridDim.x);
        const int x0 = (int)(e & 0xffu) * TMA_TX + (int)lane * 2;
        const int y = (int)((e >> 8) & 0xfffu) + warp, z = (int)(e >> 20);
        const unsigned char *stage = smem + st * S::BYTES;
        mbar_wait(smem_u32(bars + st), (unsigned)((i / STAGES) & 1));

        const bool active = x0 < G.nx && y < G.ny;
        unsigned flag_word = *reinterpret_cast<const unsigned short *>(stage + S::OFF_FLAGS + warp * TMA_TX + lane * 2);
        if (!active) flag_word = LBM_FLAG_SOLID | (LBM_FLAG_SOLID << 8);
        const bool any_fluid = __any_sync(0xffffffffu, ((flag_word & LBM_FLAG_SOLID) == 0) || ((flag_word & (LBM_FLAG_SOLID << 8)) == 0));
        V f[Q];
        CellIn<V> in;
        in.Fx = in.Fy = in.Fz = in.phase = O::bc(0.0f);
        if (any_fluid) {
            const unsigned char *mine = stage + warp * S::ROW + lane * 8;
            const unsigned char *mine_w = stage + warp * S::ROWW + lane * 8;
            static_for<0, Q>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                if constexpr (cx(q) == 0) {
                    f[q].v = *reinterpret_cast<const unsigned long long *>(mine + S::off_q(q));
                } else {
                    // cx = +1: box starts at x0 - 4, cells x - 1 = words 3 + 2*lane, 4 + 2*lane; cx = -1: box starts at
                    // x0, cells x + 1 = words 1 + 2*lane, 2 + 2*lane
                    const float *w = reinterpret_cast<const float *>(mine_w + S::off_q(q)) + (cx(q) > 0 ? 3 : 1);
                    f[q] = p2_make(w[0], w[1]);
                }
            });
            if constexpr (FORCED) {
                if (use_force) {
                    in.Fx.v = *reinterpret_cast<const unsigned long long *>(mine + S::OFF_FORCE);
                    in.Fy.v = *reinterpret_cast<const unsigned long long *>(mine + S::OFF_FORCE + S::BOX);
                    in.Fz.v = *reinterpret_cast<const unsigned long long *>(mine + S::OFF_FORCE + 2 * S::BOX);
                }
                if (use_phase) in.phase.v = *reinterpret_cast<const unsigned long long *>(mine + S::OFF_PHASE);
            }
        }
        // release the stage: the shared loads above are ordered before the arrival (release), the producer's wait
        // acquires it before the next TMA write into this stage
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bars + STAGES + st));
        if (!any_fluid) continue;

        const unsigned own = ((unsigned)(z + G.zg) * (unsigned)G.ny + (unsigned)y) * (unsigned)G.nx + (unsigned)(active ? x0 : 0);
        phys_finish<FORCED, LES, POROUS, 2, COLLIDE>(f, in, flag_word, has_phase, has_force, x0, y, z, active, own, P);
    }
}

}  // namespace lbm
