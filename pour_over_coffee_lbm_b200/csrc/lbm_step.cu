// Instantiates one group of step-kernel variants and exports a lookup function for it.
// LBM_GROUP in {0..3} = COMPAT*2 + WALLS.
//   compat = physical (groups 0, 1): every operation is explicitly rounded (lbm_phys.cuh), so there is one build.
//   compat = reference (groups 2, 3): compiled twice, LBM_STRICT_BUILD in {0,1}; strict adds -fmad=false on the nvcc
//   command line (bit-exact against the CPU oracle, which never contracts), see csrc/Makefile.
#include "lbm_step_kernel.cuh"

#ifndef LBM_GROUP
#error "LBM_GROUP must be defined (0..3)"
#endif
#ifndef LBM_STRICT_BUILD
#error "LBM_STRICT_BUILD must be defined (0/1)"
#endif

namespace lbm {

using StepKernel = void (*)(const StepArgs);

constexpr int G_COMPAT = LBM_GROUP / 2;
constexpr bool G_WALLS = (LBM_GROUP % 2) != 0;

// Resident-thread target per SM, enforced through __launch_bounds__(BLOCK, MINB): it caps registers at 64 / 128 / 128
// per thread for VEC = 1 / 2 / 4.  Without the cap ptxas takes 160+ registers for the VEC=4 kernels, only 12 warps stay
// resident and the headline kernel drops from 0.40 to 0.63 ms (measured).
template <int VEC, int BLOCK>
constexpr int min_blocks() { return (VEC == 1 ? 1024 : 512) / BLOCK; }
// the two-cell walls kernel of compat = physical fits 96 registers (16 B of spill in the full-feature variant): 20 warps
// per SM instead of 16 -- V60 512^3 2.22 -> 2.07 ms, all-fluid 512^3 box 4.84 -> 4.50 ms; 24 warps (80 registers, 88 B of
// spill) gives some of it back (2.11 / 4.59 ms)
template <int VEC, int BLOCK>
constexpr int min_blocks_phys_walls() { return VEC == 2 ? (BLOCK == 256 ? 2 : 640 / BLOCK) : min_blocks<VEC, BLOCK>(); }

// chord kernel (lbm_phys_chord.cuh), 64-thread CTAs, one tile per warp: 10.1 KB of shared memory per warp, 8 CTAs per SM = 16
// warps at 128 registers, no spills (measured on B200, V60 512^3: 16 warps 1.82 ms, 18 / 20 warps with ~90 B of spills 1.99 ms).
#ifndef LBM_CHORD_WARPS
#define LBM_CHORD_WARPS 16
#endif
template <int BLOCK>
constexpr int chord_blocks() { return LBM_CHORD_WARPS * 32 / BLOCK; }

// CTA size: the walls path runs best with small CTAs (near-wall warps take longer; a CTA slot is held until its
// slowest warp retires -- V60 512^3 sweep: 64 threads 2.17 ms, 128: 2.20, 256: 2.34)
template <int MODE, int VEC>
constexpr int default_block() { return MODE == MODE_BULK ? 64 : (VEC == 1 ? 256 : 128); }

template <int MODE, bool FORCED, bool LES, bool POROUS, int VEC, bool COLLIDE, bool DRIVE = false, bool MRT = false>
static StepKernel pick() {
    constexpr int BLOCK = default_block<MODE, VEC>();
    if constexpr (POROUS && !G_WALLS) return nullptr;          // the filter zone lives in the flag byte
    else if constexpr (!COLLIDE && LES) return nullptr;
    else if constexpr (G_WALLS && G_COMPAT == LBM_COMPAT_PHYSICAL) {
        // VEC = 4: packed quad list + wall links (lbm_phys_chord.cuh), the only kernel that fuses the pressure drive; the MRT
        // instantiations of it serve lbm_params.mrt_magic > 0 (the one- / two-cell kernels decide MRT at run time)
        if constexpr (VEC == 4) return phys_chord_kernel<FORCED, LES, POROUS, DRIVE, BLOCK, COLLIDE, chord_blocks<BLOCK>(), MRT>;
        else if constexpr (DRIVE || MRT) return nullptr;
        else return phys_walls_kernel<FORCED, LES, POROUS, VEC, BLOCK, COLLIDE, min_blocks_phys_walls<VEC, BLOCK>()>;
    }
    else if constexpr (DRIVE) return nullptr;
    else if constexpr (MRT) {      // dense periodic boxes of compat = physical: the packed VEC = 4 collision with the two-rate relaxation
        if constexpr (G_COMPAT == LBM_COMPAT_PHYSICAL && !G_WALLS && VEC == 4 && COLLIDE)
            return step_kernel<LBM_STRICT_BUILD, G_COMPAT, MODE, FORCED, LES, POROUS, VEC, BLOCK, COLLIDE, min_blocks<VEC, BLOCK>(), true>;
        else return nullptr;
    }
    else if constexpr (!COLLIDE && VEC != 1) return nullptr;
    // VEC = 2: compat = reference only -- two cells per thread, the legacy arithmetic on packed f32x2 (collide_reference_t<P2>; opt-in)
    else if constexpr (VEC == 2 && G_COMPAT != LBM_COMPAT_REFERENCE) return nullptr;
    else return step_kernel<LBM_STRICT_BUILD, G_COMPAT, MODE, FORCED, LES, POROUS, VEC, BLOCK, COLLIDE, min_blocks<VEC, BLOCK>()>;
}

template <int MODE, int VEC, bool COLLIDE, bool DRIVE, bool MRT>
static StepKernel pick_key(int key) {
    switch (key) {
        case 0: return pick<MODE, false, false, false, VEC, COLLIDE, DRIVE, MRT>();
        case 1: return pick<MODE, false, false, true, VEC, COLLIDE, DRIVE, MRT>();
        case 2: return pick<MODE, false, true, false, VEC, COLLIDE, DRIVE, MRT>();
        case 3: return pick<MODE, false, true, true, VEC, COLLIDE, DRIVE, MRT>();
        case 4: return pick<MODE, true, false, false, VEC, COLLIDE, DRIVE, MRT>();
        case 5: return pick<MODE, true, false, true, VEC, COLLIDE, DRIVE, MRT>();
        case 6: return pick<MODE, true, true, false, VEC, COLLIDE, DRIVE, MRT>();
        default: return pick<MODE, true, true, true, VEC, COLLIDE, DRIVE, MRT>();
    }
}
// `forced`: bit 0 = body_force / phase inputs, bit 1 = fused pressure-gradient drive (LBM_FEAT_DRIVE), bit 2 = the MRT instantiation
// of the four-cell kernels (the caller sets it only for vec = 4 in compat = physical with mrt_magic > 0: quad-list kernel behind walls,
// dense kernel on periodic boxes)
template <int MODE, int VEC, bool COLLIDE>
static StepKernel pick_feat(int forced, int les, int porous) {
    const int key = ((forced & 1) ? 4 : 0) | (les ? 2 : 0) | (porous ? 1 : 0);
    if (forced & 6) {
        if constexpr (COLLIDE && VEC == 4) {
            switch (forced & 6) {
                case 2: return pick_key<MODE, VEC, COLLIDE, true, false>(key);
                case 4: return pick_key<MODE, VEC, COLLIDE, false, true>(key);
                default: return pick_key<MODE, VEC, COLLIDE, true, true>(key);
            }
        } else return nullptr;
    }
    return pick_key<MODE, VEC, COLLIDE, false, false>(key);
}

#define LBM_CAT2(a, b, c) a##b##_##c
#define LBM_CAT(a, b, c) LBM_CAT2(a, b, c)
#if LBM_STRICT_BUILD
#define LBM_LOOKUP LBM_CAT(lookup_strict_g, LBM_GROUP, fn)
#else
#define LBM_LOOKUP LBM_CAT(lookup_fast_g, LBM_GROUP, fn)
#endif

// Tuning set (physical walls group, every feature on = the V60 config): VEC x BLOCK.
template <int VEC, int BLOCK>
static StepKernel tuned() {
    if constexpr (G_WALLS && G_COMPAT == LBM_COMPAT_PHYSICAL) {
        // VEC = 4: BLOCK = 128 runs 3 CTAs per SM (168 registers, no spills); BLOCK = 256 is a tuning CODE for the same
        // 128-thread CTAs with plain (not lane-mask predicated) loads
        // VEC = 4 (chord kernel): BLOCK = 128 -> 4 CTAs of 128 threads (16 warps, 128 registers); the code 256 selects
        // 64-thread CTAs at 6 per SM (12 warps, 168 registers, no spills)
        // VEC = 4 (chord kernel, default 64-thread CTAs, 8 per SM): block code 128 -> 9 CTAs per SM (112 registers)
        if constexpr (VEC == 4 && BLOCK == 256) return nullptr;
        else if constexpr (VEC == 4) return phys_chord_kernel<true, true, true, false, 64, true, 9>;
        else return phys_walls_kernel<true, true, true, VEC, BLOCK, true, min_blocks_phys_walls<VEC, BLOCK>()>;
    } else return nullptr;
}
// full-feature two-cell kernel at a higher occupancy cap: MINB CTAs of 64 threads per SM (10 -> 96 registers, 12 -> 80)
template <int MINB>
static StepKernel tuned_occ() {
    if constexpr (G_WALLS && G_COMPAT == LBM_COMPAT_PHYSICAL) return phys_walls_kernel<true, true, true, 2, 64, true, MINB>;
    else return nullptr;
}
static StepKernel pick_tuned(int vec, int block) {
    switch (vec * 1000 + block) {
        case 2065: return tuned_occ<8>();       // block codes 65 / 66: 64-thread CTAs, 16 / 24 resident warps per SM (default 20)
        case 2066: return tuned_occ<12>();
        case 1128: return tuned<1, 128>();
        case 2128: return tuned<2, 128>();
        case 2256: return tuned<2, 256>();
        case 4128: return tuned<4, 128>();
        case 4256: return tuned<4, 256>();
        default: return nullptr;
    }
}

// The step kernel of this group (dense when the group has no walls, one warp per active tile otherwise).
// *block: in = requested CTA size (0 = default), out = CTA size of the returned kernel.
// Returns nullptr when the combination is not built.
StepKernel LBM_LOOKUP(int forced, int les, int porous, int vec, int collide, int *block) {
    StepKernel k = nullptr;
    constexpr int MAIN = G_WALLS ? MODE_BULK : MODE_DENSE;
    const int def_block = vec == 1 ? default_block<MAIN, 1>() : (vec == 2 ? default_block<MAIN, 2>() : default_block<MAIN, 4>());
    if (collide && forced == 1 && les && porous && *block && *block != def_block) {      // BGK, no fused drive
        k = pick_tuned(vec, *block);
        if (k) { if (*block == 65 || *block == 66) *block = 64; if (vec == 4) *block = 64; return k; }
    }
    if (collide) {
        if (vec == 4) k = pick_feat<MAIN, 4, true>(forced, les, porous);
        else if (vec == 2) k = pick_feat<MAIN, 2, true>(forced, les, porous);
        else if (vec == 1) k = pick_feat<MAIN, 1, true>(forced, les, porous);
    } else {
        if (vec == 4 && G_WALLS && G_COMPAT == LBM_COMPAT_PHYSICAL) k = pick_feat<MAIN, 4, false>(forced, false, porous);
        else if (vec == 2) k = pick_feat<MAIN, 2, false>(forced, false, porous);
        else if (vec == 1) k = pick_feat<MAIN, 1, false>(forced, false, porous);
    }
    *block = def_block;
    return k;
}

}  // namespace lbm
