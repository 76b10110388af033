// Instantiates one group of step-kernel variants and exports a lookup function for it.
// Compiled 8 times: LBM_GROUP in {0..3} = COMPAT*2 + WALLS, LBM_STRICT_BUILD in {0,1}
// (strict adds -fmad=false on the nvcc command line; see csrc/Makefile).
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

template <int MODE, bool FORCED, bool LES, bool POROUS, int VEC, bool COLLIDE>
static StepKernel pick() {
    if constexpr (POROUS && !G_WALLS) return nullptr;          // the filter zone lives in the flag byte
    else if constexpr (!COLLIDE && (LES || VEC != 1)) return nullptr;
    else if constexpr (MODE == MODE_BOUNDARY && VEC != 1) return nullptr;
    else return step_kernel<LBM_STRICT_BUILD, G_COMPAT, MODE, FORCED, LES, POROUS, VEC, (VEC == 1 ? 256 : 128), COLLIDE>;
}

template <int MODE, int VEC, bool COLLIDE>
static StepKernel pick_feat(int forced, int les, int porous) {
    const int key = (forced ? 4 : 0) | (les ? 2 : 0) | (porous ? 1 : 0);
    switch (key) {
        case 0: return pick<MODE, false, false, false, VEC, COLLIDE>();
        case 1: return pick<MODE, false, false, true, VEC, COLLIDE>();
        case 2: return pick<MODE, false, true, false, VEC, COLLIDE>();
        case 3: return pick<MODE, false, true, true, VEC, COLLIDE>();
        case 4: return pick<MODE, true, false, false, VEC, COLLIDE>();
        case 5: return pick<MODE, true, false, true, VEC, COLLIDE>();
        case 6: return pick<MODE, true, true, false, VEC, COLLIDE>();
        default: return pick<MODE, true, true, true, VEC, COLLIDE>();
    }
}

#define LBM_CAT2(a, b, c) a##b##_##c
#define LBM_CAT(a, b, c) LBM_CAT2(a, b, c)
#if LBM_STRICT_BUILD
#define LBM_LOOKUP LBM_CAT(lookup_strict_g, LBM_GROUP, fn)
#else
#define LBM_LOOKUP LBM_CAT(lookup_fast_g, LBM_GROUP, fn)
#endif

// Tuning set (physical walls group, every feature on = the V60 config): VEC x BLOCK x occupancy target.
// hi = 1 caps registers through __launch_bounds__ so that 1024 / 640 / 512 threads per SM stay resident (VEC 1/2/4).
template <int VEC, int BLOCK, int MINB>
static StepKernel tuned() {
    if constexpr (G_WALLS && G_COMPAT == LBM_COMPAT_PHYSICAL)
        return step_kernel<LBM_STRICT_BUILD, G_COMPAT, MODE_BULK, true, true, true, VEC, BLOCK, true, MINB>;
    else return nullptr;
}
static StepKernel pick_tuned(int vec, int block, int hi) {
    switch (vec * 1000 + block) {
        case 1064: return hi ? tuned<1, 64, 16>() : tuned<1, 64, 1>();
        case 1128: return hi ? tuned<1, 128, 8>() : tuned<1, 128, 1>();
        case 1256: return hi ? tuned<1, 256, 4>() : tuned<1, 256, 1>();
        case 2064: return hi ? tuned<2, 64, 10>() : tuned<2, 64, 1>();
        case 2128: return hi ? tuned<2, 128, 5>() : tuned<2, 128, 1>();
        case 4064: return hi ? tuned<4, 64, 8>() : tuned<4, 64, 1>();
        case 4128: return hi ? tuned<4, 128, 4>() : tuned<4, 128, 1>();
        default: return nullptr;
    }
}

// `boundary` = 0: the main kernel (dense when the group has no walls, bulk-over-tiles otherwise);
// `boundary` = 1: the near-wall list kernel (walls groups only).  *block: in = requested CTA size (0 = default),
// out = CTA size of the returned kernel.  `hi` = high-occupancy register cap (tuning set only).
// Returns nullptr when the combination is not built.
StepKernel LBM_LOOKUP(int forced, int les, int porous, int vec, int collide, int boundary, int hi, int *block) {
    StepKernel k = nullptr;
    constexpr int MAIN = G_WALLS ? MODE_BULK : MODE_DENSE;
    if (boundary) {
        if constexpr (G_WALLS) {
            k = collide ? pick_feat<MODE_BOUNDARY, 1, true>(forced, les, porous) : pick_feat<MODE_BOUNDARY, 1, false>(forced, les, porous);
        }
        *block = 256;
        return k;
    }
    const int def_block = (vec == 1) ? 256 : 128;
    if (collide && forced && les && porous && (*block != def_block || hi || vec == 2)) {
        const int b = *block ? *block : def_block;
        k = pick_tuned(vec, b, hi);
        if (k) { *block = b; return k; }
    }
    if (collide) {
        if (vec == 4) k = pick_feat<MAIN, 4, true>(forced, les, porous);
        else if (vec == 1) k = pick_feat<MAIN, 1, true>(forced, les, porous);
    } else {
        if (vec == 1) k = pick_feat<MAIN, 1, false>(forced, les, porous);
    }
    *block = def_block;
    return k;
}

}  // namespace lbm
