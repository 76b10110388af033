/*
 * lbm_b200.h -- C ABI of liblbm_b200.so: the B200-native (sm_100a) D3Q19 hot path
 * of pour-over-coffee-lbm.
 *
 * This is the drop-in boundary.  Each entry point names the reference interface
 * (file:line under the reference root) whose Taichi kernels it replaces.  The
 * reference is Python, so the binding a maintainer adds is a ctypes stub; it is
 * shown in INTEGRATION.md and implemented in pour_over_coffee_lbm_b200/_lib.py.
 *
 * Conventions
 *   - plain C: device pointers + sizes, no torch / Taichi types;
 *   - every function returns 0 on success, non-zero on failure (lbm_last_error());
 *   - all device work is enqueued on the caller's stream (cudaStream_t as void*),
 *     there is no hidden synchronisation except where stated;
 *   - device memory is owned by the CALLER (torch tensors in the Python host).
 *
 * Memory layout in HBM (x fastest, z slowest; SoA):
 *   populations  g[q][zp][y][x]   f32, q = 0..18 (config/core.py:36-38 ordering),
 *                zp = z + zghost, zghost in {0,1} ghost planes on each side of a slab
 *   scalars      s[zp][y][x]      (rho, phase, flags u8, nu_sgs)
 *   vectors      v[c][zp][y][x]   c = 0..2   (u, body_force)
 * The populations stored are POST-COLLISION values; streaming happens on read
 * (pull scheme).  lbm_export_f / lbm_import_f convert to/from the reference's
 * pre-collision `f[q,i,j,k]` view exactly (pure data movement).
 *
 * Environment (read by lbm_create; diagnosis / tuning only, results are bit-identical):
 *   LBM_TMA=1            compat = physical behind walls: use the TMA-staged persistent kernel (csrc/lbm_phys_tma.cuh)
 *                        when the box does not wrap in x or y and nx % 16 == 0
 *   LBM_TMA_VARIANT=0..9 its tile height / ring depth / producer-warp count (csrc/lbm_step_tma.cu)
 *   LBM_PRODUCERS_VEC=1  multiphase / filter producers (csrc/lbm_producers.cu): one cell per thread instead of four
 *                        (read at the first launch; the four-cell path needs nx % 4 == 0 and 16-byte aligned fields anyway)
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_Q 19

/* compat modes */
#define LBM_COMPAT_PHYSICAL  0   /* consistent lattice, standard Guo, local-stress LES, Guo-Zhao drag */
#define LBM_COMPAT_REFERENCE 1   /* legacy LBMSolver arithmetic, quirks Q1..Q7 kept verbatim */

/* feature bits (lbm_params.features) */
#define LBM_FEAT_WALLS   1    /* flags consulted: solid cells skipped, halfway bounce-back, open faces */
#define LBM_FEAT_FORCE   2    /* body_force field enters through the forcing term */
#define LBM_FEAT_PHASE   4    /* phase field read: tau by phase, gravity*phase */
#define LBM_FEAT_LES     8    /* Smagorinsky eddy viscosity (physical: local Pi^neq; reference: FD on lagged u) */
#define LBM_FEAT_POROUS  16   /* filter-zone drag (physical: force; reference: post-step u damping) */
#define LBM_FEAT_DRIVE   32   /* compat=physical behind walls, four-cell kernel: the pressure-gradient drive
                                 (PressureGradientDrive, pressure_gradient_drive.py:124-193) is evaluated inside the step kernel from
                                 the previous step's rho (lbm_fields.rho_src) and added to body_force: same bits as
                                 lbm_pressure_gradient_force + lbm_step, one pass less */
#define LBM_FEAT_STRICT  64   /* compat=reference: use the -fmad=false build, bit-exact against the CPU oracle
                                 (compat=physical rounds every operation explicitly: one build, always bit-exact) */

/* flag byte per cell (lbm_fields.flags) */
#define LBM_FLAG_SOLID   1    /* LBMSolver.solid != 0           legacy/lbm_solver.py:293 */
#define LBM_FLAG_FILTER  2    /* FilterPaperSystem.filter_zone   filter_paper.py:288-364 */
#define LBM_FLAG_LES     4    /* LBMSolver.les_mask != 0         legacy/lbm_solver.py:203-205 */
#define LBM_FLAG_NEAR    8    /* some D3Q19 neighbour is solid or outside an open face (derived) */

typedef struct lbm_ctx lbm_ctx;

typedef struct {
    int nx, ny, nz;           /* local extent; nz = owned planes of this slab */
    int nz_global, z0;        /* global z extent, global index of owned plane 0 */
    int zghost;               /* ghost planes on each z side of every field (0 single GPU, 1 slabs) */
    int periodic;             /* bit0 x, bit1 y, bit2 z (global) */
    int compat;               /* LBM_COMPAT_* */
    int features;             /* LBM_FEAT_* */
    float tau_water, tau_air; /* config.TAU_WATER / TAU_AIR (0.53 / 0.8) */
    float gravity_lu;         /* config.GRAVITY_LU */
    float cs_smag;            /* Smagorinsky constant (reference hard-codes 0.18, les_turbulence.py:95) */
    float tau_min, tau_max;   /* clamp on tau_eff (0.55 / 1.90, legacy/lbm_solver.py:571) */
    float porous_darcy;       /* physical: nu/K   [1/ts]  */
    float porous_forch;       /* physical: F_eps/sqrt(K) [1/lu] */
    float K_lu, beta_lu;      /* reference: filter_paper.py:423-469 */
    float c_darcy, c_forch;   /* reference: constant folds of filter_paper.py:578-586 */
    int vec;                  /* tuning: cells per thread along x (1, 2 or 4; 0 = auto: 4 in periodic boxes and behind walls in
                                 compat = physical -- chord-fitted tiles, csrc/lbm_phys_chord.cuh -- when nx % 4 == 0 and
                                 nx <= 2048, else 2 / 1; 1 behind walls in compat = reference, where 2 selects the legacy arithmetic
                                 on packed cell pairs, csrc/lbm_step_kernel.cuh:collide_reference_t) */
    int block;                /* tuning: threads per CTA (0 = auto; 64 / 128 / 256; behind walls in compat = physical these are
                                 occupancy codes of the kernel `vec` selects, see csrc/lbm_step.cu) */
    float drive_max_force;    /* LBM_FEAT_DRIVE: clamp on |F| (PressureGradientDrive.MAX_PRESSURE_FORCE) ... */
    float drive_scale;        /* ... and the factor of the accumulation (1 in force mode, 0.5 in mixed mode) */
    float mrt_magic;          /* compat = physical: 0 = BGK (one rate).  > 0 = multiple-relaxation-time collision in its two-rate
                                 form: the even moments (density, stress: pair sums f_q + f_opp(q)) relax at 1/tau -- tau still sets
                                 the viscosity, LES acts on it -- the odd moments (momentum flux: pair differences) at 1/tau_odd with
                                 (tau - 1/2)(tau_odd - 1/2) = mrt_magic (3/16: exact wall location of halfway bounce-back, 1/4: most
                                 stable).  Every kernel of compat = physical has it (the four-cell kernels as separate instantiations,
                                 the fused drive included).  The reference names MRT only in a docstring (legacy/lbm_solver.py:833). */
} lbm_params;

typedef struct {
    float *f_src, *f_dst;     /* populations (post-collision), ping-pong; lbm_step swaps them */
    float *rho;               /* written when write_macro */
    float *u_src, *u_dst;     /* velocity ping-pong: u_src = previous step's u (FD-LES input), u_dst written */
    float *body_force;        /* may be NULL unless LBM_FEAT_FORCE */
    float *phase;             /* may be NULL unless LBM_FEAT_PHASE */
    float *blockage;          /* FilterPaperSystem.filter_blockage, may be NULL */
    uint8_t *flags;           /* may be NULL unless LBM_FEAT_WALLS */
    float *rho_src;           /* LBM_FEAT_DRIVE: rho of the previous step (read); lbm_step swaps rho / rho_src every step */
} lbm_fields;

/* ---- life cycle -------------------------------------------------------------------------- */
int  lbm_version(void);
const char *lbm_last_error(lbm_ctx *ctx);      /* ctx may be NULL: last global error */
/* Replaces backend construction + validate_platform(): src/core/lbm_unified.py:58-126,
 * src/core/backends/cuda_backend.py:40-75.  Fails (non-zero) unless `device` is sm_100. */
int  lbm_create(lbm_ctx **out, int device, const lbm_params *p);
int  lbm_set_params(lbm_ctx *ctx, const lbm_params *p);
void lbm_destroy(lbm_ctx *ctx);
/* number of kernels this library has launched since creation (bench.py gpu_launches) */
long long lbm_launch_count(lbm_ctx *ctx);

/* ---- initialisation ---------------------------------------------------------------------- */
/* LBMSolver.init_fields legacy/lbm_solver.py:1067-1112 and UnifiedLBMSolver._init_equilibrium
 * lbm_unified.py:236-248: g[q] = f_eq(rho,u) with the mode's equilibrium.  rho/u may be NULL
 * (=> rho0, u0 uniform). */
int  lbm_init_equilibrium(lbm_ctx *ctx, float *g, const float *rho, const float *u,
                          float rho0, const float u0[3], void *stream);
/* FilterPaperSystem._setup_v60_geometry / _setup_filter_zones, filter_paper.py:206-364.
 * geom = {top_radius_lu, bottom_radius_lu, cup_height_lu, air_gap_lu, paper_thickness_lu}
 * (f32-rounded Python-scope constants).  Writes solid (u8) and filter_zone (i32), either may be NULL. */
int  lbm_build_v60_geometry(lbm_ctx *ctx, uint8_t *solid, int32_t *filter_zone, const float geom[5], void *stream);
/* Packs solid/filter_zone/les_mask into the flag byte and derives LBM_FLAG_NEAR.
 * filter_zone / les_mask may be NULL (=> 0 / 1). */
int  lbm_pack_flags(lbm_ctx *ctx, uint8_t *flags, const uint8_t *solid, const int32_t *filter_zone,
                    const int32_t *les_mask, void *stream);

/* ---- the hot path ------------------------------------------------------------------------ */
/* Replaces LBMSolver.step() legacy/lbm_solver.py:817-867 (LES pre-pass, macroscopic,
 * collide+stream, swap, filter damping) and ComputeBackend.execute_collision_streaming
 * backends/cuda_backend.py:197-241 by ONE fused pull kernel per step.
 * Runs `nsteps` steps; rho/u are written on every step when write_macro_every == 1, on every
 * k-th and the last step when k > 1, never when 0.  f_src/f_dst (and u_src/u_dst) in *fields are
 * swapped in place so that f_src always holds the newest state on return.
 * comm_stream is used only when a communicator is attached (slab halo exchange). */
int  lbm_step(lbm_ctx *ctx, lbm_fields *fields, int nsteps, int write_macro_every,
              void *compute_stream, void *comm_stream);
/* compat = physical behind walls implements halfway bounce-back (legacy/lbm_solver.py:609-628) on the WRITE side:
 * a fluid cell next to a solid cell also stores its post-collision f_q into the solid cell's slot g[opp q][x+e_q],
 * where the next step's pull finds it, so solid-cell slots of g are scratch in that mode.  The library rebuilds the
 * slots itself after lbm_init_equilibrium / lbm_import_f / lbm_pack_flags / lbm_halo_exchange; a caller that writes
 * the population buffers directly (cudaMemcpy, torch ops) announces it here before the next lbm_step. */
int  lbm_populations_changed(lbm_ctx *ctx);
/* Self-test of the step kernel's packed (two cells per instruction, f32x2) arithmetic against the scalar IEEE
 * operations.  mismatches[0], [1]: correctly rounded reciprocal / square root over all 2^32 f32 bit patterns;
 * [2..5]: packed add, sub, mul, fma on pseudo-random operands (register, broadcast and literal operand forms);
 * [6]: cells whose packed collision result differs from the scalar operator's.  All must be 0.
 * Synchronises the stream. */
int  lbm_selftest_math(lbm_ctx *ctx, unsigned long long mismatches[7], void *stream);
/* LBMSolver._compute_macroscopic_quantities legacy/lbm_solver.py:488-535 on the current state
 * (streams g on the fly, writes rho and u_dst; no collision). */
int  lbm_macroscopic(lbm_ctx *ctx, const lbm_fields *fields, void *stream);
/* TopBoundary/BottomBoundary/OutletBoundary boundary_conditions.py:178-324 (observable part:
 * rho on open faces). */
int  lbm_face_bc(lbm_ctx *ctx, const lbm_fields *fields, void *stream);
/* Exact conversion between the device's post-collision populations and the reference's
 * `f[q,i,j,k]` (pre-collision, after streaming) in the SAME [q][zp][y][x] layout. */
int  lbm_export_f(lbm_ctx *ctx, const float *g, const uint8_t *flags, float *f_out, void *stream);
int  lbm_import_f(lbm_ctx *ctx, const float *f_in, const uint8_t *flags, float *g, void *stream);

/* ---- observability ------------------------------------------------------- */
/* One fused, deterministic pass over rho / u of the owned fluid cells, result left in DEVICE memory (no host sync):
 *   out8[0] max |u|   [1] min rho   [2] max rho   [3] sum rho   [4] sum 0.5 rho |u|^2   [5] NaN count   [6] Inf count
 *   [7] fluid cells.
 * Replaces visualizer.compute_statistics / get_statistics (src/visualization/visualizer.py:130-183),
 * NumericalStabilityMonitor.check_field_stability (src/core/numerical_stability.py:52-110) and the host-side
 * reductions of main.py:907-912.  flags may be NULL (every cell is fluid).  When `flags` is the field the four-cell walls kernel's
 * quad list was built for (lbm_pack_flags), the pass walks that list and never visits the solid part of the box; the sums then
 * add up in another (still fixed) order than the dense scan's. */
int  lbm_field_statistics(lbm_ctx *ctx, const float *rho, const float *u, const uint8_t *flags, double *out8, void *stream);

/* ---- neighbours that feed body_force --------------------- */
/* PressureGradientDrive.compute_pressure_gradient + _accumulate_* pressure_gradient_drive.py:124-193,274-279:
 * body_force += scale * clamp(-cs^2 grad(rho)/rho, max_force) on fluid cells. */
int  lbm_pressure_gradient_force(lbm_ctx *ctx, const float *rho, const uint8_t *flags, float *body_force,
                                 float max_force, float scale, void *stream);
/* Same force WRITTEN to body_force on fluid cells (solid cells untouched: the step kernel never reads them): replaces
 * LBMSolver.clear_body_force (legacy/lbm_solver.py:656-660) + the accumulation above when the drive is the only
 * producer of the step, and saves the full-grid clear pass. */
int  lbm_pressure_gradient_force_set(lbm_ctx *ctx, const float *rho, const uint8_t *flags, float *body_force,
                                     float max_force, float scale, void *stream);
/* PressureGradientDrive.apply_density_drive pressure_gradient_drive.py:95-122 ("method A"): rho of every fluid cell moves towards
 * the target profile by rate * (target - rho), at most max_adjust per call, clamped to [rho_min, rho_max].  target_z: device
 * array, one value per z plane of the slab incl. ghost planes (the reference's target_density depends on z only, :54-72). */
int  lbm_density_drive(lbm_ctx *ctx, float *rho, const uint8_t *flags, const float *target_z, float rate, float max_adjust,
                       float rho_min, float rho_max, void *stream);
/* FilterPaperSystem.compute_forchheimer_resistance filter_paper.py:471-536: body_force += F_drag. */
int  lbm_forchheimer_force(lbm_ctx *ctx, const float *u, const uint8_t *flags, float *body_force,
                           float fmax, void *stream);
/* LBMSolver.add_particle_reaction_forces legacy/lbm_solver.py:1478-1483: body_force += reaction on fluid. */
int  lbm_add_reaction_force(lbm_ctx *ctx, const float *reaction, const uint8_t *flags, float *body_force, void *stream);

/* ---- producers next to the step: surface tension, phase field, pouring nozzle ---------------- */
/* MultiphaseFlow3D.accumulate_surface_tension_pre_collision src/core/multiphase_3d.py:409-418 in two launches:
 * compute_gradients :111-132 (grad_phi, normal; grad_mu when mu and grad_mu are given), then compute_curvature :134-149 +
 * compute_surface_tension_force :313-332 (the live definition) + apply_surface_tension :354-363
 * (body_force += surface_force / rho on fluid cells with rho > 1e-10; body_force = NULL computes the fields only, as the
 * first three kernels of MultiphaseFlow3D.step :396-398 do).  Scalars are [zp][y][x], vectors [3][zp][y][x]; the outer
 * cell layer of every output keeps what it held (the reference never writes it).  Single slab (zghost = 0). */
int  lbm_surface_tension(lbm_ctx *ctx, const float *phi, const float *mu_or_null, const float *rho, const uint8_t *flags,
                         float *grad_phi, float *grad_mu_or_null, float *normal, float *curvature, float *surface_force,
                         float *body_force_or_null, float sigma, void *stream);
/* The two launches of lbm_surface_tension as separate calls, for z-slabs: the stencils read the ghost planes of `phi`
 * (and `mu`) in the first call and of `normal` in the second, so a slab refreshes `normal`'s ghost planes between them
 * (slab.exchange_planes over torch.distributed / NCCL; MultiphaseFlow3D does it).  Any zghost. */
int  lbm_surface_tension_gradients(lbm_ctx *ctx, const float *phi, const float *mu_or_null, float *grad_phi,
                                   float *grad_mu_or_null, float *normal, void *stream);
int  lbm_surface_tension_curvature_force(lbm_ctx *ctx, const float *phi, const float *rho, const uint8_t *flags,
                                         const float *grad_phi, const float *normal, float *curvature, float *surface_force,
                                         float *body_force_or_null, float sigma, void *stream);
/* The same chain when only body_force is wanted (what main.py:795-800 needs from accumulate_surface_tension_pre_collision):
 * ONE launch, no intermediate fields.  surface_force vanishes outside the interface band |phi| < 0.9, so a thread reads its
 * cells' phi and flags and leaves unless a fluid cell is in the band; band cells rebuild their neighbours' normals from a
 * 25-point stencil of phi with the statements of the two-launch version: body_force comes out bit-identical to
 * lbm_surface_tension(..., body_force, ...).  The reference never writes the outer cell layer of `normal` and
 * `surface_force`; pass the arrays that hold those layers (read there only) or NULL for zeros.  Single slab. */
int  lbm_surface_tension_body_force(lbm_ctx *ctx, const float *phi, const float *rho, const uint8_t *flags,
                                    const float *normal_outer_or_null, const float *surface_force_outer_or_null,
                                    float *body_force, float sigma, void *stream);
/* MultiphaseFlow3D.compute_chemical_potential multiphase_3d.py:80-109: laplacian_phi (optional output) and
 * mu = phi^3 - phi - kappa * lap(phi) on interior cells; kappa = 3 * sigma * W / 8 folded by the caller.  The reference
 * calls it once, from standardize_initial_state :542-571.  On z-slabs the ghost planes of phi must be current. */
int  lbm_chemical_potential(lbm_ctx *ctx, const float *phi, float *laplacian_phi_or_null, float *mu, float kappa, void *stream);
/* MultiphaseFlow3D.apply_surface_tension multiphase_3d.py:354-363 alone (step() with precollision_applied = False). */
int  lbm_apply_surface_tension(lbm_ctx *ctx, const float *surface_force, const float *rho, const uint8_t *flags,
                               float *body_force, void *stream);
/* The phase-field half of MultiphaseFlow3D.step multiphase_3d.py:404-407 in two launches:
 * update_phase_field_cahn_hilliard :151-197 + apply_phase_separation :334-352 (phi -> phi_new, interior cells), then
 * copy_phase_field :383-387 + update_density_from_phase :365-381 (phi = phi_new, rho, phase on every cell).
 * mu = NULL is the all-zero chemical potential the live step() leaves behind.  rho_water / rho_air are doubles because
 * the reference folds (RHO_WATER - RHO_AIR) in f64 before it meets an f32 value.  On z-slabs the ghost planes of phi (and
 * mu) must be current; the owned planes are written. */
int  lbm_phase_field_step(lbm_ctx *ctx, float *phi, float *phi_new, const float *mu_or_null, const float *u, float *rho,
                          float *phase, float mobility, float dt, double rho_water, double rho_air, void *stream);
/* MultiphaseFlow3D.update_density_from_phase multiphase_3d.py:365-381 alone (main.py:624). */
int  lbm_density_from_phase(lbm_ctx *ctx, const float *phi, float *rho, float *phase, double rho_water, double rho_air,
                            void *stream);
/* PrecisePouringSystem (src/physics/precise_pouring.py).  The host advances pour_time and evaluates
 * _get_current_pour_position :81-97 (centre or spiral); the kernels visit the nozzle's bounding box only. */
typedef struct {
    float pour_x, pour_y;         /* current nozzle centre, lattice units */
    float radius;                 /* POUR_DIAMETER_GRID / 2 */
    int   pour_z;                 /* POUR_HEIGHT (global plane index); the stream reaches 4 planes below it */
    float velocity;               /* POUR_VELOCITY, lu/ts */
    float flow_rate;              /* pour_flow_rate[None] */
    float dt;
} lbm_pour;
/* apply_pouring_force precise_pouring.py:131-163: body_force.z -= min(velocity * intensity * flow_rate / dt, 10) on the
 * fluid cells under the nozzle (Gaussian radial profile x exponential vertical decay, _is_in_pouring_region :99-129). */
int  lbm_pouring_force(lbm_ctx *ctx, const lbm_pour *pour, const uint8_t *flags, float *body_force, void *stream);
/* apply_gradual_phase_change precise_pouring.py:165-196: phi relaxes toward +1 under the nozzle (rate limited). */
int  lbm_pouring_phase_change(lbm_ctx *ctx, const lbm_pour *pour, const uint8_t *flags, float *phi, void *stream);

/* ---- coffee particles (src/physics/coffee_particles.py) ------------------------------- */
typedef struct {
    float *pos, *vel;             /* [3][n] SoA, lattice units */
    float *radius, *mass;         /* [n] SI (quirk Q9) */
    int32_t *active;              /* [n] */
    float *drag_new, *drag_old, *drag;  /* [3][n] */
    float *u_fluid;               /* [3][n]  fluid_velocity_at_particle */
    float *reynolds, *cd;         /* [n] */
    int32_t *cell;                /* [3][n] base cell (i,j,k) -- bit-exact parity target */
    int n;
} lbm_particles;
/* CoffeeParticleSystem.compute_two_way_coupling_forces + apply_under_relaxation,
 * coffee_particles.py:1107-1212: trilinear gather of u, Schiller-Naumann drag, warp-aggregated
 * atomic scatter of -drag into `reaction` ([3][zp][y][x], zeroed here first), under-relaxation. */
int  lbm_particles_couple(lbm_ctx *ctx, const float *u, float *reaction, lbm_particles *ps,
                          float water_density, float water_viscosity, float relax, void *stream);

/* The same coupling when `reaction` is a field only this call writes (LBMSolver.step_with_two_way_coupling,
 * legacy/lbm_solver.py:1485-1509: clear_body_force -> coupling -> body_force += reaction, with body_force itself as the target):
 * instead of zeroing the whole field, the deposits of the PREVIOUS call are cleared cell by cell from ps->cell (which that call
 * recorded) -- 24 M stores for 1 M particles instead of 1.6 GB at 512^3.  Precondition: `reaction` is zero before the first call
 * and nothing else writes it; ps->cell is not modified between calls.  On z-slabs the interface planes are cleared whole. */
int  lbm_particles_couple_sparse(lbm_ctx *ctx, const float *u, float *reaction, lbm_particles *ps,
                                 float water_density, float water_viscosity, float relax, void *stream);
/* lbm_particles_couple_sparse for z-slabs with a say on the interface planes: they only have to be cleared whole when a neighbour's
 * deposits were added to them since the last call (slab.reduce_ghost_up); while the bed stays away from the cuts -- the interface
 * guard of engine.particles_couple_slab -- clear_interface_planes = 0 saves six plane-sized memsets per call.  The kernel itself
 * skips particles whose base cell lies in another slab (replicated particles, owner computes), so `active` needs no masking. */
int  lbm_particles_couple_slab(lbm_ctx *ctx, const float *u, float *reaction, lbm_particles *ps,
                               float water_density, float water_viscosity, float relax, int clear_interface_planes, void *stream);

/* CoffeeParticleSystem.apply_under_relaxation coffee_particles.py:1200-1212 as a separate call
 * (lbm_particles_couple fuses it when relax >= 0; pass relax < 0 there to skip). */
int  lbm_particles_under_relax(lbm_ctx *ctx, lbm_particles *ps, float relax, void *stream);

/* CoffeeParticleSystem.update_particle_physics coffee_particles.py:641-720 (+ validate_coordinate :75-92,
 * validate_velocity :95-108, check_particle_boundary_violation_safe :734-778, constrain_to_boundary_safe :780-831):
 * explicit Euler step of the active particles with the reference's clamps (dt in [1e-8, 1e-2], |a| <= 1000,
 * |dx| <= 1 lu), the V60 cone constraint and velocity damping.  `force` is [3][n] (the reference's self.force; NULL =
 * no force) and is zeroed for the next step.  counters (device, 2 x int32): += coordinate_errors, boundary_violations.
 * The bounds are what FilterPaperSystem.get_coffee_bed_boundary() returns (main.py:672-679), in lattice units. */
typedef struct {
    float center_x, center_y, bottom_z;
    float bottom_radius_lu, top_radius_lu;
    float cup_height_lu;          /* config.CUP_HEIGHT / config.SCALE_LENGTH (50 if <= 0, coffee_particles.py:758-760) */
    float max_coordinate;         /* float(max(NX, NY, NZ)), coffee_particles.py:23 */
    float nz_minus_5;             /* float(NZ - 5), coffee_particles.py:794 */
} lbm_particle_bounds;
int  lbm_particles_advance(lbm_ctx *ctx, lbm_particles *ps, float *force, const lbm_particle_bounds *bounds, float dt,
                           int32_t *counters, void *stream);

/* CoffeeParticleSystem.apply_fluid_forces coffee_particles.py:547-639: force[p] = clamped Stokes drag (nearest-cell fluid
 * velocity) + buoyancy + gravity for every active particle that passes the reference's guards; others keep their force;
 * invalid positions deactivate the particle and count into counters[0] (may be NULL).  `force` is the [3][n] array
 * lbm_particles_advance consumes.  water_viscosity is the dynamic viscosity the reference stores
 * (WATER_VISCOSITY_90C * WATER_DENSITY_90C); doubles because the reference folds max(1e-8, mu) in f64. */
int  lbm_particles_fluid_forces(lbm_ctx *ctx, const float *u, lbm_particles *ps, float *force, double water_density,
                                double water_viscosity, double gravity, int32_t *counters, void *stream);
/* FilterPaperSystem.block_particles_at_filter src/physics/filter_paper.py:616-700: active particles above a filter-zone
 * cell (flags bit LBM_FLAG_FILTER; the particle's own plane and two either side, first hit wins) that move down bounce
 * (v_z <- -0.3 v_z), get a horizontal kick (uniform - 0.5) * noise (reference: noise = 0.01) and add 0.01 to
 * `accumulated` ([zp][y][x], atomics).  The cell index is int(pos / scale_length) as in the reference (quirk Q9).  The
 * reference's kick comes from Taichi's unseeded generator; here it is a pure function of (seed, particle, draw). */
int  lbm_particles_block_at_filter(lbm_ctx *ctx, lbm_particles *ps, const uint8_t *flags, float *accumulated,
                                   float scale_length, float noise, unsigned seed, void *stream);
/* FilterPaperSystem.update_dynamic_resistance filter_paper.py:703-746 on filter-zone cells:
 * blockage <- 0.95 blockage + 0.05 * 0.9 (1 - exp(-0.1 accumulated)); accumulated *= 0.999. */
int  lbm_filter_dynamic_resistance(lbm_ctx *ctx, const uint8_t *flags, float *blockage, float *accumulated, void *stream);

/* ---- multi-GPU slabs -------------------------------------------------------------------- */
/* Attach an NCCL communicator over the ranks of one box (z-slab chain).  unique_id is the
 * 128-byte ncclUniqueId produced by lbm_nccl_unique_id on rank 0 and broadcast by the host
 * (torch.distributed).  Replaces CUDADualGPULBMSolver.exchange_boundary_data,
 * legacy/cuda_dual_gpu_lbm.py:358-386. */
int  lbm_nccl_unique_id(void *out128);
int  lbm_attach_nccl(lbm_ctx *ctx, const void *unique_id128, int rank, int nranks);
/* exchange the 5+5 outgoing populations of the slab's boundary planes of g (and optionally the
 * ghost planes of a 3-component vector field) with the z neighbours. */
int  lbm_halo_exchange(lbm_ctx *ctx, float *g, float *vec3_or_null, void *stream);
/* The same exchange for whole ghost planes of a scalar ([zp][y][x]) and / or a 3-component vector field: rho before the
 * pressure-gradient producers read it across a slab interface (pressure_gradient_drive.py:124-177 reads rho[k -+ 1]), u before the
 * particle gather.  lbm_step keeps rho's ghost planes current by itself when LBM_FEAT_DRIVE is on. */
int  lbm_halo_exchange_field(lbm_ctx *ctx, float *scalar_or_null, float *vec3_or_null, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
