// C ABI of liblbm_b200.so (see include/lbm_b200.h).  Host-side only: parameter checking,
// kernel selection, launch geometry, slab halo exchange over NCCL.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "lbm_common.cuh"

namespace lbm {
using StepKernel = void (*)(const StepArgs);
#define DECL_LOOKUP(name) StepKernel name(int forced, int les, int porous, int vec, int collide, int *block);
DECL_LOOKUP(lookup_fast_g2_fn) DECL_LOOKUP(lookup_fast_g3_fn)
DECL_LOOKUP(lookup_strict_g0_fn) DECL_LOOKUP(lookup_strict_g1_fn) DECL_LOOKUP(lookup_strict_g2_fn) DECL_LOOKUP(lookup_strict_g3_fn)

cudaError_t launch_init_equilibrium(const Grid &, int, float *, const float *, const float *, float, const float[3], cudaStream_t);
cudaError_t launch_v60_geometry(const Grid &, uint8_t *, int32_t *, const float[5], cudaStream_t);
cudaError_t launch_pack_flags(const Grid &, uint8_t *, const uint8_t *, const int32_t *, const int32_t *, cudaStream_t);
cudaError_t build_work_lists(const Grid &, const uint8_t *, int, int, unsigned **, unsigned **, std::vector<int> &, unsigned long long **, cudaStream_t);
cudaError_t launch_wall_values(const Grid &, const float *, const unsigned long long *, const uint2 *, const unsigned *, float *, int, cudaStream_t);
cudaError_t build_chord_lists(const Grid &, const uint8_t *, unsigned long long **, uint2 **, unsigned **, float **, std::vector<int> &, unsigned long long **, long long *,
                              cudaStream_t);
cudaError_t launch_convert_f(const Grid &, bool, const float *, const uint8_t *, float *, cudaStream_t);
cudaError_t launch_face_bc(const Grid &, float *, const uint8_t *, cudaStream_t, int *);
cudaError_t launch_pressure_gradient(const Grid &, const float *, const uint8_t *, float *, float, float, int, const unsigned *, const unsigned long long *, int, int,
                                     cudaStream_t);
cudaError_t launch_forchheimer_force(const Grid &, const float *, const uint8_t *, float *, float, float, float, float, float, cudaStream_t);
cudaError_t launch_density_drive(const Grid &, float *, const uint8_t *, const float *, float, float, float, float, cudaStream_t);
cudaError_t launch_add_reaction(const Grid &, const float *, const uint8_t *, float *, cudaStream_t);
cudaError_t run_selftest_math(unsigned long long[7], const StepArgs &, cudaStream_t);
cudaError_t launch_field_statistics(const Grid &, const float *, const float *, const uint8_t *, const unsigned long long *, long long, void *, int, double *, cudaStream_t);
cudaError_t launch_bounce_slots(const Grid &, float *, const uint8_t *, const unsigned long long *, int, int, cudaStream_t);
cudaError_t launch_particles_couple(const Grid &, const float *, float *, const lbm_particles &, float, float, float, cudaStream_t);
cudaError_t launch_particles_under_relax(const lbm_particles &, float, cudaStream_t);
cudaError_t launch_particles_clear_deposits(const Grid &, float *, const lbm_particles &, cudaStream_t);
cudaError_t launch_particles_advance(const lbm_particles &, float *, const lbm_particle_bounds &, float, int *, cudaStream_t);
cudaError_t launch_surface_tension(const Grid &, int, const float *, const float *, const float *, const uint8_t *, float *, float *, float *, float *,
                                   float *, float *, float, cudaStream_t);
cudaError_t launch_apply_surface_tension(const Grid &, const float *, const float *, const uint8_t *, float *, cudaStream_t);
cudaError_t launch_surface_tension_lean(const Grid &, const float *, const float *, const uint8_t *, const float *, const float *, float *, float,
                                        cudaStream_t);
cudaError_t launch_particles_fluid_forces(const Grid &, const float *, const lbm_particles &, float *, float, float, float, float, float, int *,
                                          cudaStream_t);
cudaError_t launch_dynamic_resistance(const Grid &, const uint8_t *, float *, float *, cudaStream_t);
cudaError_t launch_particles_block_at_filter(const Grid &, const lbm_particles &, const uint8_t *, float *, float, float, unsigned, cudaStream_t);
cudaError_t launch_chemical_potential(const Grid &, const float *, float *, float *, float, cudaStream_t);
cudaError_t launch_phase_field_step(const Grid &, float *, float *, const float *, const float *, float *, float *, float, float, float, float,
                                    cudaStream_t);
cudaError_t launch_density_from_phase(const Grid &, const float *, float *, float *, float, float, cudaStream_t);
cudaError_t launch_pour(const Grid &, const lbm_pour &, const float[5], int, const uint8_t *, float *, cudaStream_t, int *);
// TMA-staged walls kernels of compat = physical (lbm_step_tma.cu)
struct TmaKernelInfo {
    void (*kernel)(const StepArgs, const TmaMaps);
    int ty, stages, threads, smem_bytes, ctas_per_sm;
};
int tma_variant_ty(int variant);
bool lookup_tma(int forced, int les, int porous, int collide, int variant, TmaKernelInfo *out);
}  // namespace lbm

using namespace lbm;

// ---- minimal NCCL binding, resolved at run time (no link-time dependency) ------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7, ncclUint8 = 1 };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::string g_error;

static int load_nccl() {
    if (g_nccl.handle) return 0;
    // torch's bundled libnccl.so.2 is already in the process when the host is PyTorch: reuse it
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_error = std::string("cannot load libnccl: ") + dlerror(); return 1; }
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { g_error = "libnccl lacks " name; return 1; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return 0;
}

struct lbm_ctx {
    int device = 0;
    lbm_params p{};
    Grid g{};
    std::string error;
    long long launches = 0;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int sm_count = 148;
    size_t max_window = 0;
    cudaEvent_t ev_boundary = nullptr, ev_comm = nullptr;
    // work lists of the walls path (built by lbm_pack_flags for the flag field it packed)
    unsigned *d_tiles = nullptr;               // active warp-tiles, packed x_segment | y << 8 | z << 20
    unsigned *d_tile_mask = nullptr;           // per warp-tile: lanes that must load (unused by the current kernels)
    unsigned long long *d_ctiles = nullptr;    // vec = 4: packed quad list (lbm_phys_chord.cuh), one u64 per lane slot ...
    uint2 *d_tile_links = nullptr;             // ... per tile: first wall link, number of links ...
    unsigned *d_links = nullptr;               // ... the wall links ...
    float *d_wall = nullptr;                   // ... and the value waiting on each link (halfway bounce-back, lbm_aux.cu)
    const float *wall_valid = nullptr;         // population buffer d_wall belongs to (nullptr: rebuild from the populations)
    long long n_links = 0;
    cudaStream_t window_stream = nullptr; bool window_set = false;
    unsigned long long *d_nbr = nullptr;       // neighbour masks [vol]
    int list_block = 0;
    std::vector<int> tile_off;                 // per owned plane offsets into the tile list (nz+1 entries)
    const uint8_t *list_flags = nullptr;
    int list_vec = 0;
    // compat = physical, walls: the population buffer whose bounce-back slots (solid-cell slots next to fluid cells,
    // see lbm_phys.cuh) are known to be current; anything else gets them rebuilt before it is stepped
    const float *slots_valid = nullptr;
    // TMA-staged walls path (lbm_phys_tma.cuh): tile height of the current list (1 = warp-tile list of the
    // register-staged kernels), tuning variant, and the tensor maps of the field sets seen so far
    void *d_stat_scratch = nullptr;            // per-block partials of lbm_field_statistics
    int list_ty = 1;
    int tma_variant = 0;
    bool tma_enabled = false;                 // opt-in (LBM_TMA=1): measured slower than the VEC = 4 register-staged kernel
    struct MapEntry { const float *pops, *force, *phase; const uint8_t *flags; int ty; TmaMaps maps; };
    std::vector<MapEntry> maps;
};

static int fail(lbm_ctx *ctx, const std::string &msg) {
    g_error = msg;
    if (ctx) ctx->error = msg;
    return 1;
}
#define CUDA_OK(ctx, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(ctx, std::string(#expr ": ") + cudaGetErrorString(e_)); } while (0)
#define NCCL_OK(ctx, expr) do { int e_ = (expr); if (e_ != ncclSuccess) return fail(ctx, std::string(#expr ": ") + g_nccl.GetErrorString(e_)); } while (0)

static int make_grid(lbm_ctx *ctx, const lbm_params *p, Grid *g) {
    if (p->nx <= 0 || p->ny <= 0 || p->nz <= 0) return fail(ctx, "grid extents must be positive");
    if (p->zghost != 0 && p->zghost != 1) return fail(ctx, "zghost must be 0 or 1");
    if (p->zghost == 0 && (p->z0 != 0 || p->nz_global != p->nz)) return fail(ctx, "zghost=0 requires the whole domain on one slab");
    if (p->z0 < 0 || p->z0 + p->nz > p->nz_global) return fail(ctx, "slab outside the global z range");
    if (!(p->features & LBM_FEAT_WALLS) && (p->periodic & 7) != 7) return fail(ctx, "without LBM_FEAT_WALLS the box must be fully periodic");
    if ((p->features & LBM_FEAT_POROUS) && !(p->features & LBM_FEAT_WALLS)) return fail(ctx, "LBM_FEAT_POROUS needs LBM_FEAT_WALLS (filter zone lives in the flag byte)");
    if (p->compat != LBM_COMPAT_PHYSICAL && p->compat != LBM_COMPAT_REFERENCE) return fail(ctx, "unknown compat mode");
    g->nx = p->nx; g->ny = p->ny; g->nz = p->nz; g->zg = p->zghost; g->nz_global = p->nz_global; g->z0 = p->z0;
    g->per_x = p->periodic & 1; g->per_y = (p->periodic >> 1) & 1; g->per_z = (p->periodic >> 2) & 1;
    g->plane = (long long)p->nx * p->ny;
    g->vol = g->plane * (p->nz + 2 * p->zghost);
    return 0;
}

static bool phys_walls(const lbm_params &p) { return p.compat == LBM_COMPAT_PHYSICAL && (p.features & LBM_FEAT_WALLS); }

static bool tma_eligible(const lbm_ctx *ctx);
static bool tma_eligible_params(const lbm_ctx *ctx) {
    const lbm_params &p = ctx->p;
    return phys_walls(p) && ctx->tma_enabled && p.vec == 0 && !(p.periodic & 3) && p.nx % 16 == 0 && p.nx >= 16 && !(p.mrt_magic > 0.0f);
}

static int pick_vec(const lbm_ctx *ctx) {
    int vec = ctx->p.vec;
    // the two-rate MRT collision (lbm_phys.cuh:collide_phys): the four-cell kernels (quad list behind walls, dense on periodic boxes) have
    // separate instantiations for it (lookup), the one- / two-cell kernels decide at run time
    if (phys_walls(ctx->p)) {
        // four cells per thread on chord-fitted tiles (lbm_phys_chord.cuh) when rows are 16-byte multiples, else two cells per
        // thread on packed f32x2 registers (lbm_phys.cuh) on 8-byte rows, else one.  The TMA-staged kernel (LBM_TMA=1, works on
        // the two-cell lists) is opt-in.
        if (vec == 0) vec = tma_eligible_params(ctx) ? 2 : 4;
        if (vec == 4 && (ctx->g.nx % 4 != 0 || ctx->g.nx < 8 || ctx->g.nx > 16384 || ctx->g.ny > 65535 || ctx->g.nz + 2 * ctx->g.zg > 65535 ||
                         ctx->g.vol >= (1ll << 32))) vec = 2;
        if (vec == 2 && (ctx->g.nx % 2 != 0 || ctx->g.nx < 4)) vec = 1;
        return vec;
    }
    // 128-bit path for dense periodic boxes (99 % of the copy bandwidth); compat = reference behind a flag field
    // runs one cell per thread (partially filled warps at every chord end make wider threads slower there)
    if (vec == 0) vec = (ctx->p.features & LBM_FEAT_WALLS) ? 1 : 4;
    // vec = 2 (opt-in): compat = reference with the legacy arithmetic on packed cell pairs (lbm_step_kernel.cuh:collide_reference_t)
    if (vec == 2 && (ctx->p.compat != LBM_COMPAT_REFERENCE || ctx->g.nx % 2 != 0 || ctx->g.nx < 4)) vec = 1;
    if (vec == 4 && (ctx->g.nx % 4 != 0 || ctx->g.nx < 8)) vec = 1;
    return vec;
}

// CTA size of the main kernel: lbm_params.block when it is one of the built sizes, else the default for `vec`
static int pick_block(const lbm_ctx *ctx, int vec) {
    const int b = ctx->p.block;
    if (b == 64 || b == 128 || b == 256 || b == 65 || b == 66) return b;      // 65 / 66: occupancy tuning codes (lbm_step.cu)
    if (ctx->p.features & LBM_FEAT_WALLS) return 64;
    return vec == 1 ? 256 : 128;
}

static StepKernel lookup(const lbm_params &p, int vec, int collide, int *block);

// The TMA-staged kernel serves compat = physical behind walls when the box does not wrap in x or y (the copy engine
// zero-fills sources outside the tensor; a periodic wrap inside a box is not expressible), rows are 16-byte multiples
// for every field incl. the u8 flags, and the caller left `vec` on auto (vec = 1 / 2 select the register-staged kernels).
static bool tma_eligible(const lbm_ctx *ctx) {
    const lbm_params &p = ctx->p;
    return phys_walls(p) && ctx->tma_enabled && p.vec == 0 && !(p.periodic & 3) && p.nx % 16 == 0 && p.nx >= 16 && !(p.mrt_magic > 0.0f);
}
static void feature_bits(const lbm_params &p, int *forced, int *les, int *porous) {
    *forced = ((p.features & (LBM_FEAT_FORCE | LBM_FEAT_PHASE)) != 0 ? 1 : 0) | ((p.features & LBM_FEAT_DRIVE) ? 2 : 0);
    *les = (p.features & LBM_FEAT_LES) != 0;
    *porous = (p.features & LBM_FEAT_POROUS) != 0;
}
// the tuning variants exist for the full-feature step kernel only: anything else runs variant 0
static int tma_variant_for(const lbm_ctx *ctx) {
    int forced, les, porous;
    feature_bits(ctx->p, &forced, &les, &porous);
    TmaKernelInfo k;
    return lookup_tma(forced, les, porous, 1, ctx->tma_variant, &k) ? ctx->tma_variant : 0;
}
// tile height of the step kernel's work list: TY of the TMA variant, or 1 for the warp-tile list
static int pick_ty(const lbm_ctx *ctx) { return tma_eligible(ctx) ? tma_variant_ty(tma_variant_for(ctx)) : 1; }

static bool chord_lists(const lbm_ctx *ctx, int vec) { return phys_walls(ctx->p) && vec == 4; }

static int rebuild_lists(lbm_ctx *ctx, const uint8_t *flags, int vec, int ty, int block, cudaStream_t s) {
    if (chord_lists(ctx, vec) && ty == 1) {
        CUDA_OK(ctx, build_chord_lists(ctx->g, flags, &ctx->d_ctiles, &ctx->d_tile_links, &ctx->d_links, &ctx->d_wall, ctx->tile_off, &ctx->d_nbr, &ctx->n_links, s));
        ctx->wall_valid = nullptr;
        ctx->launches += 6;
    } else {
        CUDA_OK(ctx, build_work_lists(ctx->g, flags, vec, ty, &ctx->d_tiles, &ctx->d_tile_mask, ctx->tile_off, &ctx->d_nbr, s));
        ctx->launches += 6;
    }
    ctx->list_flags = flags; ctx->list_vec = vec; ctx->list_ty = ty; ctx->list_block = block; ctx->window_set = false;
    ctx->slots_valid = nullptr; ctx->wall_valid = nullptr;
    return 0;
}

// ---- tensor maps ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// [comps][nzp][ny][nx] field of `elem`-byte elements, box width x ty x 1 (x 1); comps = 0: no component dimension
static int encode_map(lbm_ctx *ctx, CUtensorMap *m, const void *base, int elem, int comps, int ty, int width = 64) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(ctx, "cuTensorMapEncodeTiled is not available from this driver");
    if (((uintptr_t)base & 15u) != 0) return fail(ctx, "field pointers must be 16-byte aligned for the TMA-staged kernel (vec = 2 selects the register-staged one)");
    const Grid &G = ctx->g;
    const cuuint64_t nzp = (cuuint64_t)(G.nz + 2 * G.zg);
    cuuint64_t dims[4] = {(cuuint64_t)G.nx, (cuuint64_t)G.ny, nzp, (cuuint64_t)(comps > 0 ? comps : 1)};
    cuuint64_t strides[3] = {(cuuint64_t)G.nx * elem, (cuuint64_t)G.plane * elem, (cuuint64_t)G.vol * elem};
    cuuint32_t box[4] = {(cuuint32_t)width, (cuuint32_t)ty, 1u, 1u};
    cuuint32_t es[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(m, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, comps > 0 ? 4u : 3u,
                           const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           width == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return 0;
}
static int tensor_maps(lbm_ctx *ctx, const StepArgs &a, int ty, const TmaMaps **out) {
    for (const auto &e : ctx->maps)
        if (e.pops == a.src && e.force == a.force && e.phase == a.phase && e.flags == a.flags && e.ty == ty) { *out = &e.maps; return 0; }
    if (ctx->maps.size() >= 16) ctx->maps.clear();
    lbm_ctx::MapEntry e;
    memset(&e.maps, 0, sizeof e.maps);
    e.pops = a.src; e.force = a.force; e.phase = a.phase; e.flags = a.flags; e.ty = ty;
    if (encode_map(ctx, &e.maps.pops, a.src, 4, Q, ty)) return 1;
    if (encode_map(ctx, &e.maps.pops_wide, a.src, 4, Q, ty, 68)) return 1;
    if (a.force && encode_map(ctx, &e.maps.force, a.force, 4, 3, ty)) return 1;
    if (a.phase && encode_map(ctx, &e.maps.phase, a.phase, 4, 0, ty)) return 1;
    if (encode_map(ctx, &e.maps.flags, a.flags, 1, 0, ty)) return 1;
    ctx->maps.push_back(e);
    *out = &ctx->maps.back().maps;
    return 0;
}

extern "C" {

int lbm_version(void) { return 100; }

const char *lbm_last_error(lbm_ctx *ctx) { return ctx ? ctx->error.c_str() : g_error.c_str(); }

int lbm_create(lbm_ctx **out, int device, const lbm_params *p) {
    if (!out || !p) return fail(nullptr, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(nullptr, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (liblbm_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, "device index out of range");
    cudaDeviceProp prop;
    CUDA_OK(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        char buf[160];
        snprintf(buf, sizeof buf, "device %d is sm_%d%d; liblbm_b200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return fail(nullptr, buf);
    }
    lbm_ctx *ctx = new lbm_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_window = (size_t)prop.accessPolicyMaxWindowSize;
    if (make_grid(nullptr, p, &ctx->g)) { delete ctx; return 1; }
    ctx->p = *p;
    // tuning / diagnosis knobs of the TMA-staged walls path (scripts/tune_v60.py)
    if (const char *v = getenv("LBM_TMA_VARIANT")) ctx->tma_variant = atoi(v);
    if (const char *v = getenv("LBM_TMA")) ctx->tma_enabled = atoi(v) != 0;

    CUDA_OK(nullptr, cudaSetDevice(device));
    cudaEventCreateWithFlags(&ctx->ev_boundary, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming);
    *out = ctx;
    return 0;
}

int lbm_set_params(lbm_ctx *ctx, const lbm_params *p) {
    if (!ctx || !p) return fail(ctx, "null argument");
    Grid g;
    if (make_grid(ctx, p, &g)) return 1;
    if (g.nx != ctx->g.nx || g.ny != ctx->g.ny || g.nz != ctx->g.nz || g.zg != ctx->g.zg) {
        if (ctx->d_nbr) { cudaFree(ctx->d_nbr); ctx->d_nbr = nullptr; }
        ctx->list_flags = nullptr;
    }
    if (p->vec != ctx->p.vec || p->periodic != ctx->p.periodic || p->compat != ctx->p.compat || p->features != ctx->p.features) ctx->list_flags = nullptr;
    ctx->maps.clear();
    if (p->compat != ctx->p.compat || p->periodic != ctx->p.periodic || p->features != ctx->p.features) { ctx->slots_valid = nullptr; ctx->wall_valid = nullptr; }
    ctx->g = g; ctx->p = *p;
    return 0;
}

void lbm_destroy(lbm_ctx *ctx) {
    if (!ctx) return;
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    if (ctx->ev_boundary) cudaEventDestroy(ctx->ev_boundary);
    if (ctx->ev_comm) cudaEventDestroy(ctx->ev_comm);
    if (ctx->d_tiles) cudaFree(ctx->d_tiles);
    if (ctx->d_tile_mask) cudaFree(ctx->d_tile_mask);
    if (ctx->d_ctiles) cudaFree(ctx->d_ctiles);
    if (ctx->d_links) cudaFree(ctx->d_links);
    if (ctx->d_wall) cudaFree(ctx->d_wall);
    if (ctx->d_tile_links) cudaFree(ctx->d_tile_links);
    if (ctx->d_stat_scratch) cudaFree(ctx->d_stat_scratch);
    if (ctx->d_nbr) cudaFree(ctx->d_nbr);
    delete ctx;
}

long long lbm_launch_count(lbm_ctx *ctx) { return ctx ? ctx->launches : 0; }

int lbm_selftest_math(lbm_ctx *ctx, unsigned long long mismatches[7], void *stream) {
    if (!ctx || !mismatches) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.g = ctx->g;
    a.tau_water = 0.53f; a.tau_air = 0.8f; a.gravity_lu = 1e-5f; a.tau_min = 0.55f; a.tau_max = 1.9f;
    a.les_k = (float)(18.0 * sqrt(2.0) * 0.18 * 0.18); a.porous_darcy = 0.37f; a.porous_forch = 0.9f;
    CUDA_OK(ctx, run_selftest_math(mismatches, a, (cudaStream_t)stream));
    ctx->launches += 2;
    return 0;
}

int lbm_populations_changed(lbm_ctx *ctx) {
    if (!ctx) return fail(ctx, "null argument");
    ctx->slots_valid = nullptr; ctx->wall_valid = nullptr;
    return 0;
}

int lbm_init_equilibrium(lbm_ctx *ctx, float *g, const float *rho, const float *u, float rho0, const float u0[3], void *stream) {
    if (!ctx || !g) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    const float z[3] = {0, 0, 0};
    CUDA_OK(ctx, launch_init_equilibrium(ctx->g, ctx->p.compat, g, rho, u, rho0, u0 ? u0 : z, (cudaStream_t)stream));
    ctx->launches++;
    ctx->slots_valid = nullptr; ctx->wall_valid = nullptr;
    return 0;
}

int lbm_build_v60_geometry(lbm_ctx *ctx, uint8_t *solid, int32_t *filter_zone, const float geom[5], void *stream) {
    if (!ctx || !geom) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_v60_geometry(ctx->g, solid, filter_zone, geom, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_pack_flags(lbm_ctx *ctx, uint8_t *flags, const uint8_t *solid, const int32_t *filter_zone, const int32_t *les_mask, void *stream) {
    if (!ctx || !flags || !solid) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_pack_flags(ctx->g, flags, solid, filter_zone, les_mask, (cudaStream_t)stream));
    ctx->launches++;
    // active-tile list for the bulk kernel and the compact list of near-wall cells (synchronises the stream)
    const int vec = pick_vec(ctx);
    int block = pick_block(ctx, vec);
    lookup(ctx->p, vec, 1, &block);                 // the CTA size the step kernel will really use
    return rebuild_lists(ctx, flags, vec, pick_ty(ctx), block, (cudaStream_t)stream);
}

}  // extern "C"

// ---- step ---------------------------------------------------------------------------------------
static StepKernel lookup(const lbm_params &p, int vec, int collide, int *block) {
    const int walls = (p.features & LBM_FEAT_WALLS) != 0;
    int forced, les, porous;
    feature_bits(p, &forced, &les, &porous);
    if (!collide) forced &= 1;                       // moments only: the drive does not enter
    else if (p.compat == LBM_COMPAT_PHYSICAL && vec == 4 && p.mrt_magic > 0.0f) forced |= 4;      // MRT instantiation of the four-cell kernels
    const int group = p.compat * 2 + walls;
    const bool strict = (p.features & LBM_FEAT_STRICT) != 0;
#define LBM_PICK(g) (strict ? lookup_strict_g##g##_fn(forced, les, porous, vec, collide, block) \
                            : lookup_fast_g##g##_fn(forced, les, porous, vec, collide, block))
    switch (group) {
        // compat = physical: explicitly rounded operations, one build (LBM_FEAT_STRICT has nothing to select)
        case 0: return lookup_strict_g0_fn(forced, les, porous, vec, collide, block);
        case 1: return lookup_strict_g1_fn(forced, les, porous, vec, collide, block);
        case 2: return LBM_PICK(2);
        default: return LBM_PICK(3);
    }
#undef LBM_PICK
}

static int fill_args(lbm_ctx *ctx, const lbm_fields *f, StepArgs *a) {
    const lbm_params &p = ctx->p;
    memset(a, 0, sizeof *a);
    a->g = ctx->g;
    a->src = f->f_src; a->dst = f->f_dst; a->rho = f->rho; a->u_src = f->u_src; a->u_dst = f->u_dst;
    a->force = (p.features & LBM_FEAT_FORCE) ? f->body_force : nullptr;
    a->phase = (p.features & LBM_FEAT_PHASE) ? f->phase : nullptr;
    a->blockage = f->blockage; a->flags = f->flags;
    a->rho_src = f->rho_src; a->drive_max_force = p.drive_max_force; a->drive_scale = p.drive_scale;
    a->tau_water = p.tau_water; a->tau_air = p.tau_air; a->gravity_lu = p.gravity_lu;
    a->tau_min = p.tau_min; a->tau_max = p.tau_max;
    a->mrt_magic = p.compat == LBM_COMPAT_PHYSICAL ? p.mrt_magic : 0.0f;
    if (p.compat == LBM_COMPAT_REFERENCE) a->les_k = (p.cs_smag * 1.0f) * (p.cs_smag * 1.0f);     // les_turbulence.py:369
    else a->les_k = (float)(18.0 * sqrt(2.0) * (double)p.cs_smag * (double)p.cs_smag);
    a->porous_darcy = p.porous_darcy; a->porous_forch = p.porous_forch;
    a->K_lu = p.K_lu; a->beta_lu = p.beta_lu; a->c_darcy = p.c_darcy; a->c_forch = p.c_forch;
    if (!f->f_src) return fail(ctx, "f_src is NULL");
    if ((p.features & LBM_FEAT_WALLS) && !f->flags) return fail(ctx, "LBM_FEAT_WALLS requires a flags field");
    if ((p.features & LBM_FEAT_FORCE) && !f->body_force) return fail(ctx, "LBM_FEAT_FORCE requires body_force");
    if ((p.features & LBM_FEAT_PHASE) && !f->phase) return fail(ctx, "LBM_FEAT_PHASE requires phase");
    return 0;
}

// The kernel of one step: dense grid (periodic, no flags) or bulk over the active-tile list (walls).
struct Launcher {
    StepKernel main = nullptr;
    int block = 0, vec = 1;
    bool walls = false;
    bool tma = false;          // TMA-staged persistent kernel (compat = physical behind walls)
    bool quads = false;        // four-cell kernel on the packed quad list: bounce-back through the per-link buffer, not through slots
    TmaKernelInfo tk{};
};

static int make_launcher(lbm_ctx *ctx, const lbm_params &p, const lbm_fields *f, int vec, int collide, Launcher *L) {
    L->vec = vec;
    L->walls = (p.features & LBM_FEAT_WALLS) != 0;
    L->block = pick_block(ctx, vec);
    const int ty = pick_ty(ctx);
    if (ty > 1) {
        int forced, les, porous;
        feature_bits(p, &forced, &les, &porous);
        if (!lookup_tma(forced, les, porous, collide, tma_variant_for(ctx), &L->tk) || L->tk.ty != ty)
            return fail(ctx, "no TMA-staged kernel built for this feature combination");
        CUDA_OK(ctx, cudaFuncSetAttribute((const void *)L->tk.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L->tk.smem_bytes));
        L->tma = true;
    } else {
        L->main = lookup(p, vec, collide, &L->block);      // may fall back to the default CTA size for this variant
        if (!L->main) return fail(ctx, "no step kernel built for this feature combination");
        L->quads = chord_lists(ctx, vec);
        if (chord_lists(ctx, vec))      // the chord kernels live on shared memory: take the largest carve-out
            cudaFuncSetAttribute((const void *)L->main, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (L->walls && (ctx->list_flags != f->flags || ctx->list_vec != vec || ctx->list_ty != ty || (int)ctx->tile_off.size() != ctx->g.nz + 1 || !ctx->d_nbr))
        return fail(ctx, "work lists are stale: call lbm_pack_flags on this flags field (after any geometry or vec change)");
    return 0;
}

// launch one step on owned planes [z_begin, z_end); z_begin < 0: the two boundary planes 0 and nz - 1 of a slab in ONE launch
// (dense: blockIdx.y strides by nz - 1; walls: the copy of the two planes' list entries behind the list, see build_work_lists)
static int launch_planes(lbm_ctx *ctx, StepArgs &a, const Launcher &L, int z_begin, int z_end, cudaStream_t s) {
    const bool both = z_begin < 0;
    if (!both && z_end <= z_begin) return 0;
    if (!L.walls) {
        a.z_begin = both ? 0 : z_begin; a.z_end = both ? ctx->g.nz : z_end; a.z_stride = both ? ctx->g.nz - 1 : 1;
        const long long per_plane = (long long)(ctx->g.nx / L.vec) * ctx->g.ny;
        dim3 grid((unsigned)((per_plane + L.block - 1) / L.block), (unsigned)(both ? 2 : z_end - z_begin));
        L.main<<<grid, L.block, 0, s>>>(a);
    } else {
        const int nz = ctx->g.nz, n_t = ctx->tile_off[nz];
        const int t0 = both ? n_t : ctx->tile_off[z_begin];
        const int t1 = both ? n_t + (ctx->tile_off[1] - ctx->tile_off[0]) + (ctx->tile_off[nz] - ctx->tile_off[nz - 1]) : ctx->tile_off[z_end];
        if (t1 <= t0) return 0;
        a.items = ctx->d_tiles; a.item_mask = ctx->d_tile_mask; a.item_begin = t0; a.n_items = t1 - t0; a.nbr = ctx->d_nbr;
        a.quads = ctx->d_ctiles; a.tile_links = ctx->d_tile_links; a.links = ctx->d_links; a.wall = ctx->d_wall;
        if (L.tma) {
            const TmaMaps *maps = nullptr;
            if (tensor_maps(ctx, a, L.tk.ty, &maps)) return 1;
            const int grid = std::min(t1 - t0, ctx->sm_count * L.tk.ctas_per_sm);      // persistent CTAs, tiles strided by gridDim
            L.tk.kernel<<<(unsigned)grid, L.tk.threads, L.tk.smem_bytes, s>>>(a, *maps);
        } else {
            const int warps_per_cta = L.block / 32;
            const long long grid = (t1 - t0 + warps_per_cta - 1) / warps_per_cta;
            L.main<<<(unsigned)grid, L.block, 0, s>>>(a);
        }
    }
    CUDA_OK(ctx, cudaGetLastError());
    ctx->launches++;
    return 0;
}

// q-planes that cross a z interface: cz=+1 travel up, cz=-1 travel down
static const int UP_Q[5] = {5, 11, 12, 15, 16};
static const int DOWN_Q[5] = {6, 13, 14, 17, 18};

static int exchange(lbm_ctx *ctx, float *g, float *vec3, float *scalar, cudaStream_t s) {
    const Grid &G = ctx->g;
    if (!G.zg) return 0;
    const size_t plane = (size_t)G.plane;
    const int nq = g ? 5 : 0;                               // g == NULL: only the whole-plane fields travel
    float *top_owned = g + (size_t)(G.nz) * plane;          // physical plane nz   (last owned)
    float *bot_owned = g + (size_t)1 * plane;               // physical plane 1    (first owned)
    float *ghost_lo = g;                                    // physical plane 0
    float *ghost_hi = g + (size_t)(G.nz + 1) * plane;       // physical plane nz+1
    const int up = ctx->rank + 1, down = ctx->rank - 1;
    const bool has_up = up < ctx->nranks || G.per_z, has_down = down >= 0 || G.per_z;
    const int up_r = (up % ctx->nranks + ctx->nranks) % ctx->nranks, down_r = (down % ctx->nranks + ctx->nranks) % ctx->nranks;
    if (ctx->nranks == 1) {     // single slab with ghosts ("virtual slab" test mode): periodic wrap onto itself
        if (!G.per_z) return 0;
        for (int i = 0; i < nq; ++i) {
            CUDA_OK(ctx, cudaMemcpyAsync(ghost_lo + (size_t)UP_Q[i] * G.vol, top_owned + (size_t)UP_Q[i] * G.vol, plane * 4, cudaMemcpyDeviceToDevice, s));
            CUDA_OK(ctx, cudaMemcpyAsync(ghost_hi + (size_t)DOWN_Q[i] * G.vol, bot_owned + (size_t)DOWN_Q[i] * G.vol, plane * 4, cudaMemcpyDeviceToDevice, s));
        }
        if (vec3) for (int d = 0; d < 3; ++d) {
            float *v = vec3 + (size_t)d * G.vol;
            CUDA_OK(ctx, cudaMemcpyAsync(v, v + (size_t)G.nz * plane, plane * 4, cudaMemcpyDeviceToDevice, s));
            CUDA_OK(ctx, cudaMemcpyAsync(v + (size_t)(G.nz + 1) * plane, v + plane, plane * 4, cudaMemcpyDeviceToDevice, s));
        }
        if (scalar) {
            CUDA_OK(ctx, cudaMemcpyAsync(scalar, scalar + (size_t)G.nz * plane, plane * 4, cudaMemcpyDeviceToDevice, s));
            CUDA_OK(ctx, cudaMemcpyAsync(scalar + (size_t)(G.nz + 1) * plane, scalar + plane, plane * 4, cudaMemcpyDeviceToDevice, s));
        }
        return 0;
    }
    if (!ctx->comm) return fail(ctx, "slab exchange requested but no NCCL communicator attached");
    // NCCL matches the sends/recvs between one pair of ranks in posting order.  On a 2-rank periodic ring both
    // neighbours are the same peer, so the odd rank posts its (down, up) blocks in the opposite order.
    const bool swap_order = has_up && has_down && up_r == down_r && (ctx->rank & 1);
    auto post = [&](bool upward, const float *send, float *recv) -> int {
        const int peer = upward ? up_r : down_r;
        if (g_nccl.Send(send, plane, ncclFloat32, peer, ctx->comm, s) != ncclSuccess) return 1;
        if (g_nccl.Recv(recv, plane, ncclFloat32, peer, ctx->comm, s) != ncclSuccess) return 1;
        return 0;
    };
    int bad = 0;
    NCCL_OK(ctx, g_nccl.GroupStart());
    for (int i = 0; i < nq; ++i) {
        for (int pass = 0; pass < 2; ++pass) {
            const bool upward = (pass == 0) != swap_order;
            if (upward && has_up) bad |= post(true, top_owned + (size_t)UP_Q[i] * G.vol, ghost_hi + (size_t)DOWN_Q[i] * G.vol);
            if (!upward && has_down) bad |= post(false, bot_owned + (size_t)DOWN_Q[i] * G.vol, ghost_lo + (size_t)UP_Q[i] * G.vol);
        }
    }
    if (vec3) for (int d = 0; d < 3; ++d) {
        float *v = vec3 + (size_t)d * G.vol;
        for (int pass = 0; pass < 2; ++pass) {
            const bool upward = (pass == 0) != swap_order;
            if (upward && has_up) bad |= post(true, v + (size_t)G.nz * plane, v + (size_t)(G.nz + 1) * plane);
            if (!upward && has_down) bad |= post(false, v + plane, v);
        }
    }
    if (scalar) {
        for (int pass = 0; pass < 2; ++pass) {
            const bool upward = (pass == 0) != swap_order;
            if (upward && has_up) bad |= post(true, scalar + (size_t)G.nz * plane, scalar + (size_t)(G.nz + 1) * plane);
            if (!upward && has_down) bad |= post(false, scalar + plane, scalar);
        }
    }
    NCCL_OK(ctx, g_nccl.GroupEnd());
    if (bad) return fail(ctx, "ncclSend/ncclRecv failed while posting the halo exchange");
    return 0;
}

// compat = physical, walls: make sure the bounce-back slots of `g` are current (no-op when the step kernel wrote them)
static int ensure_slots(lbm_ctx *ctx, float *g, const uint8_t *flags, cudaStream_t s) {
    if (!phys_walls(ctx->p) || ctx->slots_valid == g) return 0;
    CUDA_OK(ctx, launch_bounce_slots(ctx->g, g, flags, ctx->d_nbr, 0, ctx->g.nz, s));
    ctx->launches++;
    ctx->slots_valid = g;
    return 0;
}
// four-cell kernel: the value waiting on every wall link (lbm_aux.cu) belongs to `g` (no-op when the step kernel left it there)
static int ensure_wall(lbm_ctx *ctx, const float *g, cudaStream_t s) {
    if (ctx->wall_valid == g) return 0;
    CUDA_OK(ctx, launch_wall_values(ctx->g, g, ctx->d_ctiles, ctx->d_tile_links, ctx->d_links, ctx->d_wall, ctx->tile_off[ctx->g.nz], s));
    ctx->launches++;
    ctx->wall_valid = g;
    return 0;
}
// after a halo exchange the incoming ghost planes have overwritten the slots that live in them
static int refresh_boundary_slots(lbm_ctx *ctx, float *g, const uint8_t *flags, cudaStream_t s) {
    if (!phys_walls(ctx->p) || !ctx->g.zg) return 0;
    CUDA_OK(ctx, launch_bounce_slots(ctx->g, g, flags, ctx->d_nbr, 0, 1, s));
    ctx->launches++;
    if (ctx->g.nz > 1) {
        CUDA_OK(ctx, launch_bounce_slots(ctx->g, g, flags, ctx->d_nbr, ctx->g.nz - 1, ctx->g.nz, s));
        ctx->launches++;
    }
    return 0;
}

extern "C" {

int lbm_step(lbm_ctx *ctx, lbm_fields *f, int nsteps, int write_macro_every, void *compute_stream, void *comm_stream) {
    if (!ctx || !f) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    if (!f->f_dst) return fail(ctx, "f_dst is NULL");
    const lbm_params &p = ctx->p;
    const int vec = pick_vec(ctx);
    Launcher L;
    if (make_launcher(ctx, p, f, vec, 1, &L)) return 1;
    const bool ref_les = p.compat == LBM_COMPAT_REFERENCE && (p.features & LBM_FEAT_LES);
    if (ref_les && write_macro_every != 1) return fail(ctx, "compat=reference with LES needs u every step (write_macro_every must be 1)");
    if (ref_les && (!f->u_src || !f->u_dst || f->u_src == f->u_dst)) return fail(ctx, "compat=reference with LES needs distinct u_src/u_dst");
    const bool drive = (p.features & LBM_FEAT_DRIVE) != 0;
    if (drive && !(phys_walls(p) && vec == 4)) return fail(ctx, "LBM_FEAT_DRIVE is fused into the four-cell walls kernel of compat=physical (LBM_FEAT_WALLS, nx % 4 == 0, vec = 0 or 4)");
    if (drive && write_macro_every != 1) return fail(ctx, "LBM_FEAT_DRIVE reads the previous step's rho (write_macro_every must be 1)");
    if (drive && (!f->rho || !f->rho_src || f->rho == f->rho_src)) return fail(ctx, "LBM_FEAT_DRIVE needs distinct rho (written) and rho_src (previous step) fields");
    cudaStream_t cs = (cudaStream_t)compute_stream, ms = (cudaStream_t)comm_stream;
    const bool slabs = ctx->g.zg == 1;
    const bool overlap = slabs && ctx->nranks > 1 && ms != nullptr && ms != cs && ctx->g.nz >= 3;
    if (L.walls && nsteps > 0 && (L.quads ? ensure_wall(ctx, f->f_src, cs) : ensure_slots(ctx, f->f_src, f->flags, cs))) return 1;
    for (int s = 0; s < nsteps; ++s) {
        StepArgs a;
        if (fill_args(ctx, f, &a)) return 1;
        const bool last = s == nsteps - 1;
        a.write_macro = write_macro_every > 0 && (write_macro_every == 1 || last || ((s + 1) % write_macro_every) == 0);
        if (a.write_macro && (!f->rho || !f->u_dst)) return fail(ctx, "write_macro requested but rho/u_dst is NULL");
        if (overlap) {
            // boundary planes first, then the halo travels on the comm stream while the interior runs
            if (L.tma) {      // the TMA-staged kernel addresses tile rows, not list copies: one launch per boundary plane
                if (launch_planes(ctx, a, L, 0, 1, cs)) return 1;
                if (launch_planes(ctx, a, L, ctx->g.nz - 1, ctx->g.nz, cs)) return 1;
            } else if (launch_planes(ctx, a, L, -1, -1, cs)) return 1;
            CUDA_OK(ctx, cudaEventRecord(ctx->ev_boundary, cs));
            CUDA_OK(ctx, cudaStreamWaitEvent(ms, ctx->ev_boundary, 0));
            if (launch_planes(ctx, a, L, 1, ctx->g.nz - 1, cs)) return 1;
            if (exchange(ctx, f->f_dst, (ref_les && a.write_macro) ? f->u_dst : nullptr, drive ? f->rho : nullptr, ms)) return 1;
            CUDA_OK(ctx, cudaEventRecord(ctx->ev_comm, ms));
            CUDA_OK(ctx, cudaStreamWaitEvent(cs, ctx->ev_comm, 0));
        } else {
            if (launch_planes(ctx, a, L, 0, ctx->g.nz, cs)) return 1;
            if (slabs && exchange(ctx, f->f_dst, (ref_les && a.write_macro) ? f->u_dst : nullptr, drive ? f->rho : nullptr, cs)) return 1;
        }
        if (L.quads) { ctx->wall_valid = f->f_dst; ctx->slots_valid = nullptr; }      // the four-cell kernel does not keep slots
        else {
            if (slabs && L.walls && refresh_boundary_slots(ctx, f->f_dst, f->flags, cs)) return 1;
            if (phys_walls(p)) { ctx->slots_valid = f->f_dst; ctx->wall_valid = nullptr; }
        }
        float *t = f->f_src; f->f_src = f->f_dst; f->f_dst = t;
        if (a.write_macro && f->u_src && f->u_src != f->u_dst) { t = f->u_src; f->u_src = f->u_dst; f->u_dst = t; }
        if (drive) { t = f->rho; f->rho = f->rho_src; f->rho_src = t; }      // rho_src = the density this step wrote
    }
    return 0;
}

int lbm_macroscopic(lbm_ctx *ctx, const lbm_fields *f, void *stream) {
    if (!ctx || !f) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    lbm_params p = ctx->p;
    p.features &= ~LBM_FEAT_LES;
    if (p.compat == LBM_COMPAT_REFERENCE) p.features &= ~LBM_FEAT_POROUS;
    const bool walls = (p.features & LBM_FEAT_WALLS) != 0;
    const int step_vec = pick_vec(ctx);
    // moments-only variants: compat = physical behind walls shares the step kernel's tile list (VEC = 1 or 2); every
    // other group has them for VEC = 1 only, so a list built for a wider step kernel is rebuilt around the call
    const int vec = phys_walls(p) ? step_vec : 1;
    const bool relist = walls && vec != step_vec;
    Launcher L;
    if (relist && rebuild_lists(ctx, f->flags, 1, 1, 256, (cudaStream_t)stream)) return 1;
    if (make_launcher(ctx, p, f, vec, 0, &L)) return 1;
    StepArgs a;
    if (fill_args(ctx, f, &a)) return 1;
    if (!f->rho || !f->u_dst) return fail(ctx, "rho/u_dst is NULL");
    if (walls && (L.quads ? ensure_wall(ctx, f->f_src, (cudaStream_t)stream) : ensure_slots(ctx, f->f_src, f->flags, (cudaStream_t)stream))) return 1;
    a.write_macro = 1;
    int rc = launch_planes(ctx, a, L, 0, ctx->g.nz, (cudaStream_t)stream);
    if (relist) {      // restore the lists of the step kernel
        int block = pick_block(ctx, step_vec);
        lookup(ctx->p, step_vec, 1, &block);
        if (rebuild_lists(ctx, f->flags, step_vec, pick_ty(ctx), block, (cudaStream_t)stream)) return 1;
    }
    return rc;
}

int lbm_face_bc(lbm_ctx *ctx, const lbm_fields *f, void *stream) {
    if (!ctx || !f || !f->rho) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    int n = 0;
    CUDA_OK(ctx, launch_face_bc(ctx->g, f->rho, f->flags, (cudaStream_t)stream, &n));
    ctx->launches += n;
    return 0;
}

int lbm_export_f(lbm_ctx *ctx, const float *g, const uint8_t *flags, float *f_out, void *stream) {
    if (!ctx || !g || !f_out || g == f_out) return fail(ctx, "bad argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_convert_f(ctx->g, true, g, flags, f_out, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_import_f(lbm_ctx *ctx, const float *f_in, const uint8_t *flags, float *g, void *stream) {
    if (!ctx || !g || !f_in || g == f_in) return fail(ctx, "bad argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_convert_f(ctx->g, false, f_in, flags, g, (cudaStream_t)stream));
    ctx->launches++;
    ctx->slots_valid = nullptr; ctx->wall_valid = nullptr;
    return 0;
}

int lbm_pressure_gradient_force(lbm_ctx *ctx, const float *rho, const uint8_t *flags, float *body_force, float max_force, float scale, void *stream) {
    if (!ctx || !rho || !body_force) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    const bool chord = chord_lists(ctx, ctx->list_vec);
    const bool listed = flags && ctx->list_flags == flags && ctx->list_ty == 1 && (chord ? ctx->d_ctiles != nullptr : ctx->d_tiles != nullptr) &&
                        (int)ctx->tile_off.size() == ctx->g.nz + 1;
    CUDA_OK(ctx, launch_pressure_gradient(ctx->g, rho, flags, body_force, max_force, scale, 1, listed && !chord ? ctx->d_tiles : nullptr,
                                          listed && chord ? ctx->d_ctiles : nullptr, listed ? ctx->tile_off[ctx->g.nz] : 0, ctx->list_vec,
                                          (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_pressure_gradient_force_set(lbm_ctx *ctx, const float *rho, const uint8_t *flags, float *body_force, float max_force, float scale, void *stream) {
    if (!ctx || !rho || !body_force) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    const bool chord = chord_lists(ctx, ctx->list_vec);
    const bool listed = flags && ctx->list_flags == flags && ctx->list_ty == 1 && (chord ? ctx->d_ctiles != nullptr : ctx->d_tiles != nullptr) &&
                        (int)ctx->tile_off.size() == ctx->g.nz + 1;
    CUDA_OK(ctx, launch_pressure_gradient(ctx->g, rho, flags, body_force, max_force, scale, 0, listed && !chord ? ctx->d_tiles : nullptr,
                                          listed && chord ? ctx->d_ctiles : nullptr, listed ? ctx->tile_off[ctx->g.nz] : 0, ctx->list_vec,
                                          (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_field_statistics(lbm_ctx *ctx, const float *rho, const float *u, const uint8_t *flags, double *out8, void *stream) {
    if (!ctx || !rho || !u || !out8) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    const int blocks = ctx->sm_count * 8;
    if (!ctx->d_stat_scratch) CUDA_OK(ctx, cudaMalloc(&ctx->d_stat_scratch, (size_t)blocks * 8 * sizeof(double)));
    // the packed quad list of the four-cell walls kernel, when it was built for exactly this flag field: the solid part of the box is
    // never visited and a quad's flags and data are fetched together (lbm_aux.cu)
    const bool listed = flags && ctx->list_flags == flags && ctx->list_ty == 1 && chord_lists(ctx, ctx->list_vec) && ctx->d_ctiles != nullptr &&
                        (int)ctx->tile_off.size() == ctx->g.nz + 1;
    CUDA_OK(ctx, launch_field_statistics(ctx->g, rho, u, flags, listed ? ctx->d_ctiles : nullptr, listed ? (long long)ctx->tile_off[ctx->g.nz] * 32 : 0,
                                         ctx->d_stat_scratch, blocks, out8, (cudaStream_t)stream));
    ctx->launches += 2;
    return 0;
}

int lbm_forchheimer_force(lbm_ctx *ctx, const float *u, const uint8_t *flags, float *body_force, float fmax, void *stream) {
    if (!ctx || !u || !flags || !body_force) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    const lbm_params &p = ctx->p;
    CUDA_OK(ctx, launch_forchheimer_force(ctx->g, u, flags, body_force, p.K_lu, p.beta_lu, p.c_darcy, p.c_forch, fmax, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_density_drive(lbm_ctx *ctx, float *rho, const uint8_t *flags, const float *target_z, float rate, float max_adjust, float rho_min,
                      float rho_max, void *stream) {
    if (!ctx || !rho || !flags || !target_z) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_density_drive(ctx->g, rho, flags, target_z, rate, max_adjust, rho_min, rho_max, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_add_reaction_force(lbm_ctx *ctx, const float *reaction, const uint8_t *flags, float *body_force, void *stream) {
    if (!ctx || !reaction || !body_force) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_add_reaction(ctx->g, reaction, flags, body_force, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

// ---- producers next to the step: surface tension, phase-field step, pouring nozzle (lbm_producers.cu) ----
static int single_slab_only(lbm_ctx *ctx, const char *what) {
    if (ctx->g.zg != 0) { ctx->error = std::string(what) + ": 7-point stencils over phi / normal are implemented for a single slab (zghost = 0)"; return 1; }
    return 0;
}

int lbm_surface_tension(lbm_ctx *ctx, const float *phi, const float *mu, const float *rho, const uint8_t *flags, float *grad_phi, float *grad_mu,
                        float *normal, float *curvature, float *surface_force, float *body_force, float sigma, void *stream) {
    if (!ctx || !phi || !grad_phi || !normal || !curvature || !surface_force) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    if (body_force && (!rho || !flags)) return fail(ctx, "lbm_surface_tension: body_force needs rho and flags");
    if (single_slab_only(ctx, "lbm_surface_tension (use the _gradients / _curvature_force pair with a ghost-plane refresh of `normal` between them)")) return 1;
    CUDA_OK(ctx, launch_surface_tension(ctx->g, 3, phi, mu, rho, flags, grad_phi, grad_mu, normal, curvature, surface_force, body_force, sigma,
                                        (cudaStream_t)stream));
    ctx->launches += 2;
    return 0;
}

int lbm_surface_tension_gradients(lbm_ctx *ctx, const float *phi, const float *mu, float *grad_phi, float *grad_mu, float *normal, void *stream) {
    if (!ctx || !phi || !grad_phi || !normal) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_surface_tension(ctx->g, 1, phi, mu, nullptr, nullptr, grad_phi, grad_mu, normal, nullptr, nullptr, nullptr, 0.0f,
                                        (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_surface_tension_curvature_force(lbm_ctx *ctx, const float *phi, const float *rho, const uint8_t *flags, const float *grad_phi,
                                        const float *normal, float *curvature, float *surface_force, float *body_force, float sigma, void *stream) {
    if (!ctx || !phi || !grad_phi || !normal || !curvature || !surface_force) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    if (body_force && (!rho || !flags)) return fail(ctx, "lbm_surface_tension_curvature_force: body_force needs rho and flags");
    CUDA_OK(ctx, launch_surface_tension(ctx->g, 2, phi, nullptr, rho, flags, const_cast<float *>(grad_phi), nullptr, const_cast<float *>(normal),
                                        curvature, surface_force, body_force, sigma, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_surface_tension_body_force(lbm_ctx *ctx, const float *phi, const float *rho, const uint8_t *flags, const float *normal_outer,
                                   const float *surface_force_outer, float *body_force, float sigma, void *stream) {
    if (!ctx || !phi || !rho || !flags || !body_force) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    if (single_slab_only(ctx, "lbm_surface_tension_body_force")) return 1;
    CUDA_OK(ctx, launch_surface_tension_lean(ctx->g, phi, rho, flags, normal_outer, surface_force_outer, body_force, sigma, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_chemical_potential(lbm_ctx *ctx, const float *phi, float *laplacian_phi, float *mu, float kappa, void *stream) {
    if (!ctx || !phi || !mu) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_chemical_potential(ctx->g, phi, laplacian_phi, mu, kappa, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_apply_surface_tension(lbm_ctx *ctx, const float *surface_force, const float *rho, const uint8_t *flags, float *body_force, void *stream) {
    if (!ctx || !surface_force || !rho || !flags || !body_force) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_apply_surface_tension(ctx->g, surface_force, rho, flags, body_force, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_phase_field_step(lbm_ctx *ctx, float *phi, float *phi_new, const float *mu, const float *u, float *rho, float *phase, float mobility,
                         float dt, double rho_water, double rho_air, void *stream) {
    if (!ctx || !phi || !phi_new || !u || !rho || !phase) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    if (phi == phi_new) return fail(ctx, "lbm_phase_field_step: phi and phi_new must be distinct buffers");
    CUDA_OK(ctx, launch_phase_field_step(ctx->g, phi, phi_new, mu, u, rho, phase, mobility, dt, (float)rho_air, (float)(rho_water - rho_air),
                                         (cudaStream_t)stream));
    ctx->launches += 2;
    return 0;
}

int lbm_density_from_phase(lbm_ctx *ctx, const float *phi, float *rho, float *phase, double rho_water, double rho_air, void *stream) {
    if (!ctx || !phi || !rho || !phase) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_density_from_phase(ctx->g, phi, rho, phase, (float)rho_air, (float)(rho_water - rho_air), (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_particles_fluid_forces(lbm_ctx *ctx, const float *u, lbm_particles *ps, float *force, double water_density, double water_viscosity,
                               double gravity, int32_t *counters, void *stream) {
    if (!ctx || !u || !ps || !force) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    const Grid &g = ctx->g;
    const float max_coord = (float)std::max(g.nx, std::max(g.ny, g.nz_global));
    const float mu_safe = (float)std::max(1e-8, water_viscosity);             // ti.max(1e-8, self.water_viscosity): folded in f64
    const float vol_k = (float)((4.0 / 3.0) * 3.14159);                        // (4.0/3.0) * 3.14159: folded in f64
    CUDA_OK(ctx, launch_particles_fluid_forces(g, u, *ps, force, (float)water_density, mu_safe, (float)gravity, vol_k, max_coord, counters,
                                               (cudaStream_t)stream));
    ctx->launches += ps->n > 0 ? 1 : 0;
    return 0;
}

int lbm_filter_dynamic_resistance(lbm_ctx *ctx, const uint8_t *flags, float *blockage, float *accumulated, void *stream) {
    if (!ctx || !flags || !blockage || !accumulated) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_dynamic_resistance(ctx->g, flags, blockage, accumulated, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_particles_block_at_filter(lbm_ctx *ctx, lbm_particles *ps, const uint8_t *flags, float *accumulated, float scale_length, float noise,
                                  unsigned seed, void *stream) {
    if (!ctx || !ps || !flags || !accumulated) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    if (!(scale_length > 0.0f)) return fail(ctx, "lbm_particles_block_at_filter: scale_length must be positive");
    CUDA_OK(ctx, launch_particles_block_at_filter(ctx->g, *ps, flags, accumulated, scale_length, noise, seed, (cudaStream_t)stream));
    ctx->launches += ps->n > 0 ? 1 : 0;
    return 0;
}

static int pour_common(lbm_ctx *ctx, const lbm_pour *pour, int mode, const uint8_t *flags, float *field, void *stream) {
    if (!ctx || !pour || !flags || !field) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);
    if (!(pour->radius > 0.0f)) return fail(ctx, "lbm_pour: radius must be positive");
    float decay[5];
    for (int d = 0; d < 5; ++d) decay[d] = (float)exp(-(double)d / 2.0);     // the reference folds this constant expression in f64
    int launched = 0;
    CUDA_OK(ctx, launch_pour(ctx->g, *pour, decay, mode, flags, field, (cudaStream_t)stream, &launched));
    ctx->launches += launched;
    return 0;
}
int lbm_pouring_force(lbm_ctx *ctx, const lbm_pour *pour, const uint8_t *flags, float *body_force, void *stream) {
    return pour_common(ctx, pour, 0, flags, body_force, stream);
}
int lbm_pouring_phase_change(lbm_ctx *ctx, const lbm_pour *pour, const uint8_t *flags, float *phi, void *stream) {
    return pour_common(ctx, pour, 1, flags, phi, stream);
}

int lbm_particles_couple(lbm_ctx *ctx, const float *u, float *reaction, lbm_particles *ps, float water_density,
                         float water_viscosity, float relax, void *stream) {
    if (!ctx || !u || !reaction || !ps) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, cudaMemsetAsync(reaction, 0, (size_t)ctx->g.vol * 3 * sizeof(float), (cudaStream_t)stream));
    CUDA_OK(ctx, launch_particles_couple(ctx->g, u, reaction, *ps, water_density, water_viscosity, relax, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

int lbm_particles_couple_sparse(lbm_ctx *ctx, const float *u, float *reaction, lbm_particles *ps, float water_density,
                                float water_viscosity, float relax, void *stream) {
    return lbm_particles_couple_slab(ctx, u, reaction, ps, water_density, water_viscosity, relax, 1, stream);
}

int lbm_particles_couple_slab(lbm_ctx *ctx, const float *u, float *reaction, lbm_particles *ps, float water_density,
                              float water_viscosity, float relax, int clear_interface_planes, void *stream) {
    if (!ctx || !u || !reaction || !ps || !ps->cell) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);
    const Grid &G = ctx->g;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_OK(ctx, launch_particles_clear_deposits(G, reaction, *ps, s));
    if (G.zg && clear_interface_planes) {      // slabs: what the neighbours' particles left in the interface planes (slab.reduce_ghost_up) is not in this rank's cell list
        for (int d = 0; d < 3; ++d) {
            float *r = reaction + (size_t)d * G.vol;
            CUDA_OK(ctx, cudaMemsetAsync(r, 0, (size_t)G.plane * 2 * sizeof(float), s));
            CUDA_OK(ctx, cudaMemsetAsync(r + (size_t)(G.nz + 1) * G.plane, 0, (size_t)G.plane * sizeof(float), s));
        }
    }
    CUDA_OK(ctx, launch_particles_couple(G, u, reaction, *ps, water_density, water_viscosity, relax, s));
    ctx->launches += ps->n > 0 ? 2 : 0;
    return 0;
}

int lbm_particles_advance(lbm_ctx *ctx, lbm_particles *ps, float *force, const lbm_particle_bounds *bounds, float dt, int32_t *counters, void *stream) {
    if (!ctx || !ps || !bounds || !counters) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    if (!ps->pos || !ps->vel || !ps->mass || !ps->active) return fail(ctx, "particle arrays pos/vel/mass/active are required");
    CUDA_OK(ctx, launch_particles_advance(*ps, force, *bounds, dt, counters, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

int lbm_particles_under_relax(lbm_ctx *ctx, lbm_particles *ps, float relax, void *stream) {
    if (!ctx || !ps) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    CUDA_OK(ctx, launch_particles_under_relax(*ps, relax, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

int lbm_nccl_unique_id(void *out128) {
    if (!out128) return fail(nullptr, "null argument");
    if (load_nccl()) return 1;
    ncclUniqueId id;
    NCCL_OK(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return 0;
}

int lbm_attach_nccl(lbm_ctx *ctx, const void *unique_id128, int rank, int nranks) {
    if (!ctx || !unique_id128) return fail(ctx, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, "bad rank/nranks");
    ctx->rank = rank; ctx->nranks = nranks;
    if (nranks == 1) return 0;
    if (load_nccl()) return fail(ctx, g_error);
    ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof id);
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    NCCL_OK(ctx, g_nccl.CommInitRank(&ctx->comm, nranks, id, rank));
    return 0;
}

int lbm_halo_exchange(lbm_ctx *ctx, float *g, float *vec3_or_null, void *stream) {
    if (!ctx || !g) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    ctx->slots_valid = nullptr; ctx->wall_valid = nullptr;      // the incoming ghost planes overwrite the bounce-back slots that live in them
    return exchange(ctx, g, vec3_or_null, nullptr, (cudaStream_t)stream);
}

int lbm_halo_exchange_field(lbm_ctx *ctx, float *scalar_or_null, float *vec3_or_null, void *stream) {
    if (!ctx || (!scalar_or_null && !vec3_or_null)) return fail(ctx, "null argument");
    cudaSetDevice(ctx->device);      // launches follow the context's device, whatever the caller's current device is
    return exchange(ctx, nullptr, vec3_or_null, scalar_or_null, (cudaStream_t)stream);
}

}  // extern "C"
