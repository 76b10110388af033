// Instantiates the TMA-staged walls kernels of compat = physical (lbm_phys_tma.cuh) and exports their lookup.
#include "lbm_phys_tma.cuh"

namespace lbm {

struct TmaKernelInfo {
    void (*kernel)(const StepArgs, const TmaMaps);
    int ty;             // rows per tile = consumer warps per CTA
    int stages;
    int threads;        // (ty + producer warps) * 32
    int smem_bytes;     // dynamic shared memory incl. alignment slack and barriers
    int ctas_per_sm;
};

template <bool FORCED, bool LES, bool POROUS, int TY, int STAGES, bool COLLIDE, int MINB, int NP = 2>
static TmaKernelInfo info() {
    TmaKernelInfo k;
    k.kernel = phys_tma_kernel<FORCED, LES, POROUS, TY, STAGES, COLLIDE, MINB, NP>;
    k.ty = TY; k.stages = STAGES; k.threads = (TY + NP) * 32;
    k.smem_bytes = STAGES * TmaStage<TY>::BYTES + 2 * STAGES * 8 + 128;
    k.ctas_per_sm = MINB;
    return k;
}

template <int TY, int STAGES, int MINB, bool COLLIDE>
static TmaKernelInfo pick_feat(int forced, int les, int porous) {
    const int key = (forced ? 4 : 0) | (les ? 2 : 0) | (porous ? 1 : 0);
    switch (key) {
        case 0: return info<false, false, false, TY, STAGES, COLLIDE, MINB>();
        case 1: return info<false, false, true, TY, STAGES, COLLIDE, MINB>();
        case 2: return info<false, COLLIDE, false, TY, STAGES, COLLIDE, MINB>();
        case 3: return info<false, COLLIDE, true, TY, STAGES, COLLIDE, MINB>();
        case 4: return info<true, false, false, TY, STAGES, COLLIDE, MINB>();
        case 5: return info<true, false, true, TY, STAGES, COLLIDE, MINB>();
        case 6: return info<true, COLLIDE, false, TY, STAGES, COLLIDE, MINB>();
        default: return info<true, COLLIDE, true, TY, STAGES, COLLIDE, MINB>();
    }
}

// Tile shapes / ring depths.  Variant 0 is the default; the others exist for the full-feature step kernel only
// (tuning set, selected with LBM_TMA_VARIANT, see scripts/tune_v60.py).
int tma_variant_ty(int variant) {
    switch (variant) {
        case 2: case 3: case 9: return 8;
        case 4: case 5: return 2;
        default: return 4;
    }
}

bool lookup_tma(int forced, int les, int porous, int collide, int variant, TmaKernelInfo *out) {
    if (!collide) {      // moments only (lbm_macroscopic): one kernel per tile height, on the step kernel's tile list
        switch (tma_variant_ty(variant)) {
            case 8: *out = pick_feat<8, 3, 1, false>(forced, 0, porous); break;
            case 2: *out = pick_feat<2, 4, 4, false>(forced, 0, porous); break;
            default: *out = pick_feat<4, 4, 2, false>(forced, 0, porous); break;
        }
        return variant >= 0 && variant <= 9;
    }
    const bool full = forced && les && porous;
    switch (variant) {       //                                           TY  S  coll CTAs NP
        case 0: *out = pick_feat<4, 4, 2, true>(forced, les, porous); return true;      // NP = 2
        case 1: if (!full) return false; *out = info<true, true, true, 4, 2, true, 3, 2>(); return true;
        case 2: if (!full) return false; *out = info<true, true, true, 8, 4, true, 1, 4>(); return true;
        case 3: if (!full) return false; *out = info<true, true, true, 8, 3, true, 1, 2>(); return true;
        case 4: if (!full) return false; *out = info<true, true, true, 2, 4, true, 4, 1>(); return true;
        case 5: if (!full) return false; *out = info<true, true, true, 2, 5, true, 3, 2>(); return true;
        case 6: if (!full) return false; *out = info<true, true, true, 4, 4, true, 2, 4>(); return true;
        case 7: if (!full) return false; *out = info<true, true, true, 4, 4, true, 2, 1>(); return true;
        case 8: if (!full) return false; *out = info<true, true, true, 4, 3, true, 2, 4>(); return true;
        case 9: if (!full) return false; *out = info<true, true, true, 8, 4, true, 1, 8>(); return true;
        default: return false;
    }
}

}  // namespace lbm
