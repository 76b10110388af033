"""CPU ORACLE (test infrastructure, NOT product code) -- NumPy restatement of the
reference's D3Q19 time step.

PARITY STATUS: pinned against the reference's own source code.  The reference
(latteine1217/pour-over-coffee-lbm) is pure Python + Taichi, and Taichi is not
installable in the authoring container (no wheel, no network) -- but its kernels are
plain Python functions, so tests/golden/make_reference_goldens.py imports the
UNMODIFIED reference modules from /root/reference under a small pure-Python stand-in
for the `taichi` package (tests/golden/taichi_shim: fields = NumPy arrays, IEEE f32
scalar arithmetic in source order, constants folded like Taichi folds them) and
records what LBMSolver.step(), FilterPaperSystem (geometry, Forchheimer),
PressureGradientDrive and CoffeeParticleSystem (coupling, under-relaxation,
integrator) compute on seeded 16^3 states.  This module reproduces every recorded
output BIT FOR BIT (tests/test_oracle_vs_reference_run.py; fixtures
tests/golden/reference_run_*.npz).  Not covered by that pin: Taichi's own code
generation on a real back end (fast-math reassociation / FMA contraction), which no
source-level restatement can see.  Also pinned: every known-answer identity the
reference's own tests hold for this path -- tests/test_oracle_known_answers.py.
Each function cites the file:line it restates (paths relative to the reference root).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.  The product path (pour_over_coffee_lbm_b200) never does.

Conventions
-----------
* compat = reference: all arithmetic is IEEE f32 with the reference's
  left-to-right evaluation order and NO fused multiply-add (NumPy never
  contracts).  The CUDA "strict" build (-fmad=false) follows the same order, so
  parity is bit-exact there.
* compat = physical: an explicit operation-by-operation contract that includes
  fused multiply-adds (`_fma`, exact C99 fmaf); see step_physical.
* Arrays use the reference's logical index order: f[q, i, j, k], u[i, j, k, c]
  (x=i, y=j, z=k; Taichi dense layout, k fastest).
* Two modes:
    reference  LBMSolver quirks kept verbatim (Q1 equilibrium velocity table,
               Q3 clamped Guo-like term, lagged finite-difference LES, stale
               inflow, post-step u damping + face density writes).
    physical   consistent velocity set, standard Guo forcing, local
               non-equilibrium-stress Smagorinsky, Guo-Zhao porous drag,
               periodic or bounce-back faces.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------
# Lattice (config/core.py:36-47)
# --------------------------------------------------------------------------
Q = 19
CX = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0], dtype=np.int32)
CY = np.array([0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1], dtype=np.int32)
CZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1], dtype=np.int32)
W = np.array([1.0 / 3.0] + [1.0 / 18.0] * 6 + [1.0 / 36.0] * 12, dtype=np.float32)

# Velocity table used ONLY by the reference's equilibrium (quirk Q1):
# src/core/lbm_algorithms.py:154-165.  Differs from config at q=8<->10, 12<->14, 16<->18.
EQ_CX = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0], dtype=np.int32)
EQ_CY = np.array([0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1], dtype=np.int32)
EQ_CZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1], dtype=np.int32)


def opposite_table() -> np.ndarray:
    """legacy/lbm_solver.py:431-439 -- opp computed from the config velocity set."""
    opp = np.zeros(Q, dtype=np.int32)
    for q in range(Q):
        for p in range(Q):
            if CX[q] == -CX[p] and CY[q] == -CY[p] and CZ[q] == -CZ[p]:
                opp[q] = p
    return opp


OPP = opposite_table()


# --------------------------------------------------------------------------
# Constants (config/core.py, config/physics.py; Appendix B of SURVEY.md)
# --------------------------------------------------------------------------
@dataclass
class RefConfig:
    """The subset of the reference's `config` module the hot path reads."""
    NX: int = 224
    NY: int = 224
    NZ: int = 224
    TAU_WATER: float = 0.53          # config/core.py:64 (TAU_FLUID), config/__init__.py:285
    TAU_AIR: float = 0.8             # config/core.py:65
    GRAVITY_LU: float = 44.145       # config/physics.py:175-177
    PHYSICAL_DOMAIN_SIZE: float = 0.14
    SCALE_VELOCITY: float = 0.01
    TIME_SCALE_OPTIMIZATION_FACTOR: float = 1.2
    TOP_RADIUS: float = 0.058
    BOTTOM_RADIUS: float = 0.01
    CUP_HEIGHT: float = 0.085
    WATER_VISCOSITY_90C: float = 3.15e-07
    WATER_DENSITY_90C: float = 965.3
    PARTICLE_DIAMETER_MM: float = 0.65
    COFFEE_BEAN_DENSITY: float = 1200.0
    DT: float = 1.0
    CS2: float = 1.0 / 3.0
    USE_LES: bool = True             # ENABLE_LES and RE_CHAR(5397) > 500 (legacy/lbm_solver.py:36,94)
    LES_CS: float = 0.18             # les_turbulence.py:95 (hard-coded, not config's 0.17)
    # SCALE_LENGTH = PHYSICAL_DOMAIN_SIZE / NZ, frozen at import (config/core.py:78)
    SCALE_LENGTH: float = field(default=0.0)
    SCALE_TIME: float = field(default=0.0)

    def __post_init__(self):
        if self.SCALE_LENGTH == 0.0:
            self.SCALE_LENGTH = self.PHYSICAL_DOMAIN_SIZE / self.NZ
        if self.SCALE_TIME == 0.0:
            self.SCALE_TIME = (self.SCALE_LENGTH / self.SCALE_VELOCITY) * self.TIME_SCALE_OPTIMIZATION_FACTOR


# --------------------------------------------------------------------------
# State
# --------------------------------------------------------------------------
@dataclass
class State:
    cfg: RefConfig
    f: np.ndarray            # [19,NX,NY,NZ] f32   legacy/lbm_solver.py:229
    f_new: np.ndarray        # [19,NX,NY,NZ] f32   :230
    rho: np.ndarray          # [NX,NY,NZ]
    u: np.ndarray            # [NX,NY,NZ,3]
    u_sq: np.ndarray         # [NX,NY,NZ]
    phase: np.ndarray        # [NX,NY,NZ]
    solid: np.ndarray        # [NX,NY,NZ] u8        :293
    body_force: np.ndarray   # [NX,NY,NZ,3]         :311
    les_mask: np.ndarray     # [NX,NY,NZ] i32       :203-205
    nu_sgs: np.ndarray       # [NX,NY,NZ]           les_turbulence.py:98
    # filter-paper system (optional; physics/filter_paper.py)
    filter_zone: np.ndarray | None = None   # i32
    filter_blockage: np.ndarray | None = None
    K_lu: np.float32 = F32(0.0)
    beta_lu: np.float32 = F32(0.0)
    apply_filter: bool = False   # boundary manager has a filter system attached
    apply_faces: bool = True     # boundary manager top/bottom/outlet strategies


def init_fields(cfg: RefConfig) -> State:
    """LBMSolver.init_fields, legacy/lbm_solver.py:1067-1112: rho=1, u=0, phase=0, F=0, f=f_new=w_q."""
    shp = (cfg.NX, cfg.NY, cfg.NZ)
    f = np.empty((Q,) + shp, dtype=F32)
    for q in range(Q):
        f[q] = W[q] * F32(1.0)
    return State(
        cfg=cfg, f=f, f_new=f.copy(),
        rho=np.ones(shp, F32), u=np.zeros(shp + (3,), F32), u_sq=np.zeros(shp, F32),
        phase=np.zeros(shp, F32), solid=np.zeros(shp, np.uint8),
        body_force=np.zeros(shp + (3,), F32), les_mask=np.ones(shp, np.int32),
        nu_sgs=np.zeros(shp, F32))


# --------------------------------------------------------------------------
# small f32 helpers (explicit evaluation order)
# --------------------------------------------------------------------------
def _dot3(ax, ay, az, bx, by, bz):
    """Taichi Vector.dot: ((a0*b0 + a1*b1) + a2*b2) in f32."""
    return (ax * bx + ay * by) + az * bz


def _edot(ex: int, ey: int, ez: int, vx, vy, vz):
    """e_q . v with e components in {0,+1,-1}: products are exact, zeros add exactly,
    so the result is the (single-rounding) sum of the non-zero terms in x,y,z order."""
    acc = None
    for e, v in ((ex, vx), (ey, vy), (ez, vz)):
        if e == 0:
            continue
        term = v if e > 0 else -v
        acc = term if acc is None else acc + term
    if acc is None:
        return np.zeros_like(vx)
    return acc


def equilibrium_ref(rho, ux, uy, uz, q: int, table: str = "eq"):
    """equilibrium_d3q19_unified, src/core/lbm_algorithms.py:183-218.
    f_eq = w_q*rho*(1 + 3*eu + 4.5*eu*eu - 1.5*u_sq), left to right, with the
    lbm_algorithms velocity table (table="eq", quirk Q1) or the config one."""
    if table == "eq":
        ex, ey, ez = int(EQ_CX[q]), int(EQ_CY[q]), int(EQ_CZ[q])
    else:
        ex, ey, ez = int(CX[q]), int(CY[q]), int(CZ[q])
    eu = _edot(ex, ey, ez, ux, uy, uz)
    u_sq = _dot3(ux, uy, uz, ux, uy, uz)
    return (W[q] * rho) * (((F32(1.0) + F32(3.0) * eu) + (F32(4.5) * eu) * eu) - F32(1.5) * u_sq)


# --------------------------------------------------------------------------
# Step 0: LES pre-pass (lagged finite differences on u)
# --------------------------------------------------------------------------
def les_update(st: State) -> None:
    """LESTurbulenceModel._compute_sgs_from_vector, src/physics/les_turbulence.py:318-380.
    Interior cells only; outer layer forced to 0 (:377-380)."""
    u = st.u
    nu = np.zeros_like(st.nu_sgs)
    c = (slice(1, -1), slice(1, -1), slice(1, -1))
    ip = (slice(2, None), slice(1, -1), slice(1, -1)); im = (slice(0, -2), slice(1, -1), slice(1, -1))
    jp = (slice(1, -1), slice(2, None), slice(1, -1)); jm = (slice(1, -1), slice(0, -2), slice(1, -1))
    kp = (slice(1, -1), slice(1, -1), slice(2, None)); km = (slice(1, -1), slice(1, -1), slice(0, -2))
    h = F32(0.5)
    dudx = (u[ip + (0,)] - u[im + (0,)]) * h
    dudy = (u[jp + (0,)] - u[jm + (0,)]) * h
    dudz = (u[kp + (0,)] - u[km + (0,)]) * h
    dvdx = (u[ip + (1,)] - u[im + (1,)]) * h
    dvdy = (u[jp + (1,)] - u[jm + (1,)]) * h
    dvdz = (u[kp + (1,)] - u[km + (1,)]) * h
    dwdx = (u[ip + (2,)] - u[im + (2,)]) * h
    dwdy = (u[jp + (2,)] - u[jm + (2,)]) * h
    dwdz = (u[kp + (2,)] - u[km + (2,)]) * h
    S11, S22, S33 = dudx, dvdy, dwdz
    S12 = h * (dudy + dvdx)
    S13 = h * (dudz + dwdx)
    S23 = h * (dvdz + dwdy)
    mag = np.sqrt(F32(2.0) * (((S11 * S11 + S22 * S22) + S33 * S33)
                              + F32(2.0) * ((S12 * S12 + S13 * S13) + S23 * S23)))
    cs = F32(st.cfg.LES_CS)
    cs_delta_sqr = (cs * F32(1.0)) * (cs * F32(1.0))
    val = np.minimum(cs_delta_sqr * mag, F32(0.1))
    # cut-offs in the reference's order: mask, low shear, interface band
    val = np.where(np.abs(st.phase[c]) < F32(0.9), F32(0.0), val)
    val = np.where(mag < F32(1e-3), F32(0.0), val)
    val = np.where(st.les_mask[c] == 0, F32(0.0), val)
    nu[c] = val
    st.nu_sgs = nu.astype(F32)


# --------------------------------------------------------------------------
# Step 1: macroscopic moments
# --------------------------------------------------------------------------
def _gravity_z(cfg: RefConfig, phase):
    """_compute_body_force, legacy/lbm_solver.py:581-592: (0,0,-GRAVITY_LU*phase) if phase>0.001."""
    g = F32(cfg.GRAVITY_LU) * phase
    return np.where(phase > F32(0.001), -g, F32(0.0)).astype(F32)


def macroscopic(st: State) -> None:
    """LBMSolver._compute_macroscopic_quantities, legacy/lbm_solver.py:488-535.
    Fluid cells only; solid cells keep their old rho,u."""
    f = st.f
    rho = np.zeros_like(st.rho)
    for q in range(Q):
        rho = rho + f[q]
    mx = np.zeros_like(rho); my = np.zeros_like(rho); mz = np.zeros_like(rho)
    for q in range(Q):
        # mom += f_q * e_q (vector f32); products with 0 add exactly
        if CX[q]:
            mx = mx + f[q] * F32(CX[q])
        if CY[q]:
            my = my + f[q] * F32(CY[q])
        if CZ[q]:
            mz = mz + f[q] * F32(CZ[q])
    Fx = F32(0.0) + st.body_force[..., 0]
    Fy = F32(0.0) + st.body_force[..., 1]
    Fz = _gravity_z(st.cfg, st.phase) + st.body_force[..., 2]
    ok = rho > F32(1e-12)
    safe = np.where(ok, rho, F32(1.0))
    ux = np.where(ok, (mx + F32(0.5) * Fx) / safe, F32(0.0)).astype(F32)
    uy = np.where(ok, (my + F32(0.5) * Fy) / safe, F32(0.0)).astype(F32)
    uz = np.where(ok, (mz + F32(0.5) * Fz) / safe, F32(0.0)).astype(F32)
    fluid = st.solid == 0
    st.rho = np.where(fluid, rho, st.rho).astype(F32)
    for c, v in enumerate((ux, uy, uz)):
        st.u[..., c] = np.where(fluid, v, st.u[..., c])
    st.u_sq = np.where(fluid, _dot3(ux, uy, uz, ux, uy, uz), st.u_sq).astype(F32)


# --------------------------------------------------------------------------
# Step 2: collide + push-stream with halfway bounce-back
# --------------------------------------------------------------------------
def guo_term_ref(q: int, ux, uy, uz, Fx, Fy, Fz, tau):
    """_compute_forcing_term + _compute_stable_guo_forcing + _prepare_forcing_parameters +
    _calculate_forcing_terms, legacy/lbm_solver.py:594-607, 688-764 (quirk Q3)."""
    fnorm = np.sqrt(_dot3(Fx, Fy, Fz, Fx, Fy, Fz))
    active = fnorm > F32(1e-15)
    tau_safe = np.minimum(np.maximum(tau, F32(0.6)), F32(1.5))
    big = fnorm > F32(10.0)
    scale_f = np.where(big, F32(10.0) / np.where(big, fnorm, F32(1.0)), F32(1.0)).astype(F32)
    fsx, fsy, fsz = Fx * scale_f, Fy * scale_f, Fz * scale_f
    unorm = np.sqrt(_dot3(ux, uy, uz, ux, uy, uz))
    fast = unorm > F32(0.2)
    su = (F32(0.2) / np.where(fast, unorm, F32(1.0))).astype(F32)
    usx = np.where(fast, ux * su, ux); usy = np.where(fast, uy * su, uy); usz = np.where(fast, uz * su, uz)
    ex, ey, ez = int(CX[q]), int(CY[q]), int(CZ[q])
    eu = _edot(ex, ey, ez, usx, usy, usz)
    ef = _edot(ex, ey, ez, fsx, fsy, fsz)
    uf = _dot3(usx, usy, usz, fsx, fsy, fsz)
    coeff = W[q] * (F32(1.0) - F32(0.5) / tau_safe)
    term1 = F32(3.0) * ef
    term2 = (F32(9.0) * eu) * uf
    out = coeff * (term1 + term2)
    out = np.maximum(F32(-0.5), np.minimum(F32(0.5), out))
    return np.where(active, out, F32(0.0)).astype(F32)


def collide_stream(st: State) -> None:
    """LBMSolver._apply_collision_and_streaming + _perform_streaming,
    legacy/lbm_solver.py:537-579, 609-628.  Push scheme; writes f_new only where the
    reference writes (everything else keeps its previous content: quirk Q6)."""
    cfg = st.cfg
    NX, NY, NZ = cfg.NX, cfg.NY, cfg.NZ
    fluid = st.solid == 0
    rho = st.rho
    ux, uy, uz = st.u[..., 0], st.u[..., 1], st.u[..., 2]
    Fx = F32(0.0) + st.body_force[..., 0]
    Fy = F32(0.0) + st.body_force[..., 1]
    Fz = _gravity_z(cfg, st.phase) + st.body_force[..., 2]
    tau = np.where(st.phase > F32(0.5), F32(cfg.TAU_WATER), F32(cfg.TAU_AIR)).astype(F32)
    if cfg.USE_LES:
        tau = tau + F32(3.0) * st.nu_sgs
    tau = np.maximum(F32(0.55), np.minimum(F32(1.90), tau)).astype(F32)
    omega = (F32(1.0) / tau).astype(F32)
    solid = st.solid
    for q in range(Q):
        feq = equilibrium_ref(rho, ux, uy, uz, q, "eq")
        Fq = guo_term_ref(q, ux, uy, uz, Fx, Fy, Fz, tau)
        fq = st.f[q]
        f_post = ((fq - omega * (fq - feq)) + Fq).astype(F32)
        ex, ey, ez = int(CX[q]), int(CY[q]), int(CZ[q])
        # source-side window whose target x+e is inside the domain
        sx = slice(max(0, -ex), NX - max(0, ex)); tx = slice(max(0, ex), NX - max(0, -ex))
        sy = slice(max(0, -ey), NY - max(0, ey)); ty = slice(max(0, ey), NY - max(0, -ey))
        sz = slice(max(0, -ez), NZ - max(0, ez)); tz = slice(max(0, ez), NZ - max(0, -ez))
        src_fluid = fluid[sx, sy, sz]
        tgt_solid = solid[tx, ty, tz] != 0
        fp = f_post[sx, sy, sz]
        # target fluid -> f_new[q, x+e] = f_post
        wr = src_fluid & ~tgt_solid
        view = st.f_new[q][tx, ty, tz]
        view[wr] = fp[wr]
        # target solid -> f_new[opp q, x] = f_post   (halfway bounce-back)
        bb = src_fluid & tgt_solid
        view2 = st.f_new[int(OPP[q])][sx, sy, sz]
        view2[bb] = fp[bb]
        # out of domain: dropped


def swap_fields(st: State) -> None:
    """LBMSolver.swap_fields, legacy/lbm_solver.py:630-654 (element-wise exchange)."""
    st.f, st.f_new = st.f_new, st.f


# --------------------------------------------------------------------------
# Step 3: boundary manager (only the parts with an observable effect)
# --------------------------------------------------------------------------
def forchheimer_params(cfg: RefConfig, porosity: float = 0.85):
    """FilterPaperSystem._initialize_forchheimer_parameters, filter_paper.py:423-469.
    Kernel-scope locals are f32 Taichi variables; integer powers expand to multiplies."""
    dp = F32(cfg.PARTICLE_DIAMETER_MM * 1e-3)
    p = F32(porosity)
    one_m = F32(1.0) - p
    K_phys = ((dp * dp) * ((p * p) * p)) / (F32(180.0) * (one_m * one_m))
    beta = (F32(1.75) * one_m) / ((p * p) * p)
    K_lu = K_phys / F32(cfg.SCALE_LENGTH ** 2)
    return F32(K_lu), F32(beta)


def filter_constants(cfg: RefConfig):
    """Python-scope constant folds in apply_filter_effects / compute_forchheimer_resistance,
    filter_paper.py:514-520, 578-586 (f64 folds, then used as f32 literals)."""
    c_darcy = F32(cfg.WATER_VISCOSITY_90C * cfg.SCALE_TIME / (cfg.SCALE_LENGTH ** 2))
    c_forch = F32(cfg.WATER_DENSITY_90C * cfg.SCALE_TIME ** 2 / (cfg.SCALE_LENGTH ** 3))
    return c_darcy, c_forch


def apply_filter_effects(st: State) -> None:
    """FilterPaperSystem.apply_filter_effects, filter_paper.py:538-614 (damps u in the zone)."""
    if st.filter_zone is None:
        return
    cfg = st.cfg
    c = (slice(1, -1), slice(1, -1), slice(1, -1))
    ux = st.u[c + (0,)]; uy = st.u[c + (1,)]; uz = st.u[c + (2,)]
    umag = np.sqrt(_dot3(ux, uy, uz, ux, uy, uz))
    K = st.K_lu; beta = st.beta_lu
    sel = (st.filter_zone[c] == 1) & (st.solid[c] == 0) & (umag > F32(1e-8)) & (K > F32(1e-12))
    if not np.any(sel):
        return
    c_darcy, c_forch = filter_constants(cfg)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        darcy = c_darcy / K
        forch = ((c_forch * beta) * umag) / np.sqrt(K)
        blk = st.filter_blockage[c] if st.filter_blockage is not None else F32(0.0)
        total = (darcy + forch) * (F32(1.0) + blk)
        r = np.exp(((-total) * F32(0.5)).astype(F32)).astype(F32)
    r = np.maximum(F32(0.1), r)
    hf = (r + F32(1.0)) * F32(0.5)
    st.u[c + (2,)] = np.where(sel, uz * r, uz)
    st.u[c + (0,)] = np.where(sel, ux * hf, ux)
    st.u[c + (1,)] = np.where(sel, uy * hf, uy)


def face_bcs(st: State) -> None:
    """TopBoundary/BottomBoundary/OutletBoundary SoA kernels, boundary_conditions.py:178-324.
    LBMSolver has `ux`, so the SoA branch runs: it writes `rho` and the ux/uy/uz copies that
    LBMSolver never reads (quirk Q5) -- only the rho writes are observable.  Order:
    top -> bottom -> outlet(x faces; y faces; z=0), each a serial Taichi offload."""
    rho = st.rho; fl = st.solid == 0
    # top (:264-283)
    rho[:, :, -1] = np.where(fl[:, :, -1], F32(1.0), rho[:, :, -1])
    # bottom (:286-324)
    rho[:, :, 0] = np.where(fl[:, :, 0], rho[:, :, 1], rho[:, :, 0])
    # outlet (:178-261)
    rho[0, :, :] = np.where(fl[0, :, :], rho[1, :, :], rho[0, :, :])
    rho[-1, :, :] = np.where(fl[-1, :, :], rho[-2, :, :], rho[-1, :, :])
    rho[:, 0, :] = np.where(fl[:, 0, :], rho[:, 1, :], rho[:, 0, :])
    rho[:, -1, :] = np.where(fl[:, -1, :], rho[:, -2, :], rho[:, -1, :])
    rho[:, :, 0] = np.where(fl[:, :, 0], rho[:, :, 1], rho[:, :, 0])


def step(st: State) -> None:
    """LBMSolver.step, legacy/lbm_solver.py:817-867 (A.3 recipe of SURVEY.md)."""
    if st.cfg.USE_LES:
        les_update(st)
    macroscopic(st)
    collide_stream(st)
    swap_fields(st)
    # boundary manager: bounce-back strategy is a net identity (each pair swapped twice,
    # boundary_conditions.py:165-175) and is therefore not restated.
    if st.apply_filter:
        apply_filter_effects(st)
    if st.apply_faces:
        face_bcs(st)


# --------------------------------------------------------------------------
# Geometry (init)
# --------------------------------------------------------------------------
def _radius(cfg: RefConfig):
    i = np.arange(cfg.NX, dtype=F32)[:, None, None]
    j = np.arange(cfg.NY, dtype=F32)[None, :, None]
    cx = F32(cfg.NX * 0.5); cy = F32(cfg.NY * 0.5)
    dx = i - cx; dy = j - cy
    return np.sqrt(dx * dx + dy * dy).astype(F32)   # [NX,NY,1]


def v60_solid(cfg: RefConfig) -> np.ndarray:
    """FilterPaperSystem._setup_v60_geometry, src/physics/filter_paper.py:206-286."""
    NX, NY, NZ = cfg.NX, cfg.NY, cfg.NZ
    top_r = F32(cfg.TOP_RADIUS / cfg.SCALE_LENGTH)
    bot_r = F32(cfg.BOTTOM_RADIUS / cfg.SCALE_LENGTH)
    cup_h = F32(cfg.CUP_HEIGHT / cfg.SCALE_LENGTH)
    bottom_z = F32(5.0)
    top_z = bottom_z + cup_h
    wall = F32(2.0)
    gap = F32(0.002 / cfg.SCALE_LENGTH)
    r = np.broadcast_to(_radius(cfg), (NX, NY, NZ))
    z = np.broadcast_to(np.arange(NZ, dtype=F32)[None, None, :], (NX, NY, NZ))
    hr = (z - bottom_z) / cup_h
    inner = bot_r + (top_r - bot_r) * hr
    solid = np.where(z <= bottom_z, r > bot_r,
                     np.where(z <= top_z, r > (inner + gap) + wall, r > top_r + wall))
    i = np.arange(NX)[:, None, None]; j = np.arange(NY)[None, :, None]; k = np.arange(NZ)[None, None, :]
    edge = (i <= 2) | (i >= NX - 3) | (j <= 2) | (j >= NY - 3) | (k <= 2) | (k >= NZ - 3)
    return (solid | edge).astype(np.uint8)


def filter_zones(cfg: RefConfig, paper_thickness: float = 0.0001) -> np.ndarray:
    """FilterPaperSystem._setup_filter_zones, filter_paper.py:288-364."""
    NX, NY, NZ = cfg.NX, cfg.NY, cfg.NZ
    top_r = F32(cfg.TOP_RADIUS / cfg.SCALE_LENGTH)
    bot_r = F32(cfg.BOTTOM_RADIUS / cfg.SCALE_LENGTH)
    cup_h = F32(cfg.CUP_HEIGHT / cfg.SCALE_LENGTH)
    f_top = F32(5.0) + cup_h
    f_bot = F32(5.0)
    thick = np.maximum(F32(1.0), F32(paper_thickness / cfg.SCALE_LENGTH))
    gap = F32(0.002 / cfg.SCALE_LENGTH)
    r = np.broadcast_to(_radius(cfg), (NX, NY, NZ))
    z = np.broadcast_to(np.arange(NZ, dtype=F32)[None, None, :], (NX, NY, NZ))
    hr = (z - f_bot) / cup_h
    hr = np.maximum(F32(0.0), np.minimum(F32(1.0), hr))
    inner_r = bot_r + (top_r - bot_r) * hr
    f_out = inner_r - gap
    f_in = f_out - thick
    side = (z >= f_bot) & (z <= f_top) & (f_in <= r) & (r <= f_out)
    trans = bot_r - gap
    bottom = ~((z >= f_bot) & (z <= f_top)) & (z >= f_bot - thick) & (z < f_bot) & (r <= trans)
    return (side | bottom).astype(np.int32)


def attach_filter_system(st: State) -> None:
    """FilterPaperSystem.initialize_filter_geometry, filter_paper.py:136-197: writes lbm.solid,
    filter_zone, Forchheimer params and punches the zone out of les_mask (:199-204)."""
    cfg = st.cfg
    st.solid = v60_solid(cfg)
    st.filter_zone = filter_zones(cfg)
    st.filter_blockage = np.zeros(st.rho.shape, F32)
    st.K_lu, st.beta_lu = forchheimer_params(cfg)
    st.les_mask = np.where(st.filter_zone == 1, 0, st.les_mask).astype(np.int32)
    st.apply_filter = True


# --------------------------------------------------------------------------
# Neighbour force producers (inputs of the step)
# --------------------------------------------------------------------------
def compute_forchheimer_resistance(st: State, scale_velocity: float = 0.01) -> None:
    """FilterPaperSystem.compute_forchheimer_resistance, filter_paper.py:471-536 (body_force += F)."""
    cfg = st.cfg
    c = (slice(1, -1), slice(1, -1), slice(1, -1))
    ux = st.u[c + (0,)]; uy = st.u[c + (1,)]; uz = st.u[c + (2,)]
    umag = np.sqrt(_dot3(ux, uy, uz, ux, uy, uz))
    K = st.K_lu; beta = st.beta_lu
    sel = (st.filter_zone[c] == 1) & (st.solid[c] == 0) & (umag > F32(1e-8)) & (K > F32(1e-12))
    c_darcy, c_forch = filter_constants(cfg)
    with np.errstate(all="ignore"):
        coeff = c_darcy / K + ((c_forch * beta) * umag) / np.sqrt(K)
        rx, ry, rz = (-coeff) * ux, (-coeff) * uy, (-coeff) * uz
        mag = np.sqrt(_dot3(rx, ry, rz, rx, ry, rz))
        fmax = F32(0.01 * scale_velocity / cfg.DT)
        big = mag > fmax
        s = np.where(big, fmax / np.where(big, mag, F32(1.0)), F32(1.0)).astype(F32)
    for comp, r in enumerate((rx, ry, rz)):
        cur = st.body_force[c + (comp,)]
        st.body_force[c + (comp,)] = np.where(sel, cur + np.where(big, r * s, r), cur)


def density_drive_target(nz: int) -> np.ndarray:
    """PressureGradientDrive.initialize_target_density, pressure_gradient_drive.py:54-72: the profile along z (the field does not
    depend on i, j).  f32; expressions of Python floats fold in f64 before they meet a kernel value (Taichi's constant folding)."""
    zr = np.arange(nz, dtype=F32) / F32(nz)
    hi = F32(1.0) + ((zr - F32(0.8)) / F32(1.0 - 0.8)) * F32(1.8 - 1.0)
    lo = F32(0.4) + (zr / F32(0.2)) * F32(1.0 - 0.4)
    return np.where(zr >= F32(0.8), hi, np.where(zr <= F32(0.2), lo, F32(1.0))).astype(F32)


def density_drive(rho: np.ndarray, solid: np.ndarray, target_z: np.ndarray, rate: float = 0.025, max_adjust: float = 0.001,
                  rho_min: float = 0.5, rho_max: float = 2.0) -> np.ndarray:
    """PressureGradientDrive.apply_density_drive, pressure_gradient_drive.py:95-122 (method A), one call.  rho [NX,NY,NZ]."""
    diff = target_z[None, None, :].astype(F32) - rho
    adj = diff * F32(rate)
    adj = np.where(np.abs(adj) > F32(max_adjust), np.where(adj > 0, F32(max_adjust), F32(-max_adjust)), adj).astype(F32)
    new = np.maximum(F32(rho_min), np.minimum(F32(rho_max), rho + adj)).astype(F32)
    return np.where(solid == 0, new, rho).astype(F32)


def pressure_gradient_force(st: State, max_force: float = 0.12) -> np.ndarray:
    """PressureGradientDrive.compute_pressure_gradient, pressure_gradient_drive.py:124-177.
    Returns pressure_force [NX,NY,NZ,3] (zero on solid cells, which the reference leaves untouched)."""
    rho = st.rho
    g = []
    for ax in range(3):
        d = np.zeros_like(rho)
        sl = [slice(None)] * 3
        lo = list(sl); hi = list(sl); mid = list(sl)
        mid[ax] = slice(1, -1); hi[ax] = slice(2, None); lo[ax] = slice(0, -2)
        d[tuple(mid)] = (rho[tuple(hi)] - rho[tuple(lo)]) * F32(0.5)
        a0 = list(sl); a1 = list(sl); a0[ax] = 0; a1[ax] = 1
        d[tuple(a0)] = rho[tuple(a1)] - rho[tuple(a0)]
        b0 = list(sl); b1 = list(sl); b0[ax] = -1; b1[ax] = -2
        d[tuple(b0)] = rho[tuple(b0)] - rho[tuple(b1)]
        g.append(d)
    cs2 = F32(st.cfg.CS2)
    ok = rho > F32(1e-12)
    safe = np.where(ok, rho, F32(1.0))
    fx = -(g[0] * cs2) / safe; fy = -(g[1] * cs2) / safe; fz = -(g[2] * cs2) / safe
    mag = np.sqrt(_dot3(fx, fy, fz, fx, fy, fz))
    mf = F32(max_force)
    big = mag > mf
    s = (mf / np.where(big, mag, F32(1.0))).astype(F32)
    out = np.zeros(rho.shape + (3,), F32)
    fluid = st.solid == 0
    for comp, v in enumerate((fx, fy, fz)):
        v = np.where(big, v * s, v)
        v = np.where(ok, v, F32(0.0))
        out[..., comp] = np.where(fluid, v, F32(0.0))
    return out


def accumulate_pressure_force(st: State, pf: np.ndarray, scale: float = 1.0) -> None:
    """_accumulate_pressure_force_to_body_force / _accumulate_mixed_pressure_force,
    pressure_gradient_drive.py:188-193, 274-279."""
    fluid = (st.solid == 0)[..., None]
    add = pf if scale == 1.0 else F32(scale) * pf
    st.body_force = np.where(fluid, st.body_force + add, st.body_force).astype(F32)


# --------------------------------------------------------------------------
# Particles: two-way coupling (src/physics/coffee_particles.py)
# --------------------------------------------------------------------------
def particle_cell_and_weights(cfg: RefConfig, pos: np.ndarray):
    """interpolate_fluid_velocity_from_field / distribute_force_to_grid index+weights,
    coffee_particles.py:1048-1076, 1156-1183.  i=int(max(0,min(N-2,x))) (f32 clamp, truncation)."""
    out = []
    for c, n in enumerate((cfg.NX, cfg.NY, cfg.NZ)):
        x = pos[:, c].astype(F32)
        cl = np.maximum(F32(0.0), np.minimum(F32(n - 2), x))
        i = cl.astype(np.int32)
        fr = x - i.astype(F32)
        fr = np.maximum(F32(0.0), np.minimum(F32(1.0), fr))
        out.append((i, fr))
    (i, fx), (j, fy), (k, fz) = out
    one = F32(1.0)
    w = {
        (0, 0, 0): ((one - fx) * (one - fy)) * (one - fz),
        (0, 0, 1): ((one - fx) * (one - fy)) * fz,
        (0, 1, 0): ((one - fx) * fy) * (one - fz),
        (0, 1, 1): ((one - fx) * fy) * fz,
        (1, 0, 0): (fx * (one - fy)) * (one - fz),
        (1, 0, 1): (fx * (one - fy)) * fz,
        (1, 1, 0): (fx * fy) * (one - fz),
        (1, 1, 1): (fx * fy) * fz,
    }
    return i, j, k, w


# reference summation order of the 8 corners in the gather (coffee_particles.py:1186-1196)
_GATHER_ORDER = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]
# order of the 8 atomic adds in the scatter (:1079-1086)
_SCATTER_ORDER = [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1)]


def interpolate_velocity(cfg: RefConfig, u: np.ndarray, pos: np.ndarray) -> np.ndarray:
    """interpolate_fluid_velocity_from_field, coffee_particles.py:1156-1198."""
    i, j, k, w = particle_cell_and_weights(cfg, pos)
    acc = None
    for (a, b, c) in _GATHER_ORDER:
        term = w[(a, b, c)][:, None] * u[i + a, j + b, k + c, :]
        acc = term if acc is None else acc + term
    return acc.astype(F32)


def drag_coefficient(re_p):
    """compute_drag_coefficient, coffee_particles.py:1088-1099."""
    cd = F32(24.0) / np.maximum(F32(0.01), re_p)
    mid = (re_p >= F32(0.1)) & (re_p < F32(1000.0))
    safe = np.where(mid, re_p, F32(1.0))
    sn = (F32(24.0) / safe) * (F32(1.0) + F32(0.15) * np.power(safe, F32(0.687)).astype(F32))
    cd = np.where(mid, sn, cd)
    cd = np.where(re_p >= F32(1000.0), F32(0.44), cd)
    return cd.astype(F32)


def two_way_coupling(cfg: RefConfig, u: np.ndarray, pos, vel, radius, mass, active, sequential: bool = False):
    """compute_two_way_coupling_forces, coffee_particles.py:1107-1154.
    Returns (drag_force_new [P,3], reaction_force_field [NX,NY,NZ,3], u_fluid, re_p, cd, cell[P,3]).
    The scatter is summed in f32, corner by corner over all particles (atomics are order-free in the reference; compare with
    a tolerance) -- or, with sequential=True, particle by particle in the reference's corner order, which is the order a
    serial execution of the reference's loop produces (bit-exact against the recorded runs; slow, small P only)."""
    P = pos.shape[0]
    rho_w = F32(cfg.WATER_DENSITY_90C)
    mu_w = F32(cfg.WATER_VISCOSITY_90C * cfg.WATER_DENSITY_90C)
    act = active != 0
    u_fl = interpolate_velocity(cfg, u, pos)
    rel = (u_fl - vel).astype(F32)
    mag = np.sqrt(_dot3(rel[:, 0], rel[:, 1], rel[:, 2], rel[:, 0], rel[:, 1], rel[:, 2]))
    mov = act & (mag > F32(1e-8))
    safe_mag = np.where(mov, mag, F32(1.0))
    re_p = (((rho_w * safe_mag) * F32(2.0)) * radius) / np.maximum(F32(1e-8), mu_w)
    cd = drag_coefficient(re_p)
    area = (F32(3.14159) * radius) * radius
    dmag = (((F32(0.5) * rho_w) * cd) * area) * safe_mag
    dmag = np.minimum(dmag, mass * F32(100.0))
    drag_new = np.zeros((P, 3), F32)
    for c in range(3):
        drag_new[:, c] = np.where(mov, (dmag * rel[:, c]) / safe_mag, F32(0.0))
    re_out = np.where(mov, re_p, F32(0.0)).astype(F32)
    cd_out = np.where(mov, cd, F32(0.0)).astype(F32)
    i, j, k, w = particle_cell_and_weights(cfg, pos)
    field_ = np.zeros((cfg.NX, cfg.NY, cfg.NZ, 3), F32)
    react = -drag_new
    if sequential:
        for p in np.nonzero(mov)[0]:
            for (a, b, c) in _SCATTER_ORDER:
                cell_ = (i[p] + a, j[p] + b, k[p] + c)
                field_[cell_] = field_[cell_] + (w[(a, b, c)][p] * react[p]).astype(F32)
    else:
        for (a, b, c) in _SCATTER_ORDER:
            contrib = (w[(a, b, c)][:, None] * react).astype(F32)
            contrib[~mov] = 0
            for comp in range(3):
                np.add.at(field_[..., comp], (i + a, j + b, k + c), contrib[:, comp])
    u_fl_out = np.where(act[:, None], u_fl, F32(0.0)).astype(F32)
    cell = np.stack([i, j, k], axis=1).astype(np.int32)
    return drag_new, field_, u_fl_out, re_out, cd_out, cell


def under_relax(drag_new, drag_old, active, alpha: float):
    """apply_under_relaxation, coffee_particles.py:1200-1212: F = a*F_new + (1-a)*F_old; old <- F."""
    a = F32(alpha)
    out = (a * drag_new + (F32(1.0) - a) * drag_old).astype(F32)
    act = (active != 0)[:, None]
    drag = np.where(act, out, F32(0.0)).astype(F32)
    new_old = np.where(act, out, drag_old).astype(F32)
    return drag, new_old


def _valid_coordinate(x, y, z, max_coord) -> bool:
    """CoffeeParticleSystem.validate_coordinate, coffee_particles.py:75-92 (MIN_COORDINATE = 0, :24)."""
    ok = not (x < 0 or x > max_coord or y < 0 or y > max_coord or z < 0 or z > max_coord)
    if not (x == x and y == y and z == z):
        ok = False
    if abs(x) > F32(1e6) or abs(y) > F32(1e6) or abs(z) > F32(1e6):
        ok = False
    return ok


def _valid_velocity(vx, vy, vz) -> bool:
    """validate_velocity, coffee_particles.py:95-108 (MAX_VELOCITY = 10, :25)."""
    s2 = (vx * vx + vy * vy) + vz * vz
    ok = vx == vx and vy == vy and vz == vz
    if s2 > F32(10.0) * F32(10.0):
        ok = False
    return bool(ok)


def update_particle_physics(cfg: RefConfig, pos, vel, force, mass, active, dt, center_x, center_y, bottom_z,
                            bottom_radius_lu, top_radius_lu):
    """CoffeeParticleSystem.update_particle_physics, coffee_particles.py:641-720, with
    check_particle_boundary_violation_safe (:734-778) and constrain_to_boundary_safe (:780-831), statement by
    statement in f32.  Plain Python loop over the particles (test sizes only).  pos/vel/force [P,3], mass [P],
    active [P] int32 are updated in place; returns (coordinate_errors, boundary_violations)."""
    f = F32
    dt_safe = max(f(1e-8), min(f(1e-2), f(dt)))
    cx0, cy0, bz = f(center_x), f(center_y), f(bottom_z)
    br, tr = f(bottom_radius_lu), f(top_radius_lu)
    max_coord = f(max(cfg.NX, cfg.NY, cfg.NZ))
    cup = f(cfg.CUP_HEIGHT / cfg.SCALE_LENGTH)      # Python-scope constant, rounded to f32 when the kernel uses it
    if not cup > 0:
        cup = f(50.0)
    nz5 = f(cfg.NZ - 5)
    coord_err = 0; viol = 0
    norm = lambda a, b, c: np.sqrt((a * a + b * b) + c * c)
    with np.errstate(all="ignore"):
        for i in range(len(mass)):
            if active[i] != 1:
                continue
            px, py, pz = f(pos[i, 0]), f(pos[i, 1]), f(pos[i, 2])
            if not _valid_coordinate(px, py, pz, max_coord):
                active[i] = 0; coord_err += 1
                continue
            vx, vy, vz = f(vel[i, 0]), f(vel[i, 1]), f(vel[i, 2])
            m = f(mass[i])
            if m > f(1e-10):
                ax, ay, az = f(force[i, 0]) / m, f(force[i, 1]) / m, f(force[i, 2]) / m
                amag = norm(ax, ay, az)
                if amag > f(1000.0):
                    sc = f(1000.0) / amag
                    ax, ay, az = ax * sc, ay * sc, az * sc
                nvx, nvy, nvz = vx + ax * dt_safe, vy + ay * dt_safe, vz + az * dt_safe
                if _valid_velocity(nvx, nvy, nvz):
                    vx, vy, vz = nvx, nvy, nvz
                else:
                    vx = vy = vz = f(0.0); coord_err += 1
            dx, dy, dz = vx * dt_safe, vy * dt_safe, vz * dt_safe
            dmag = norm(dx, dy, dz)
            if dmag > f(1.0):
                sc = f(1.0) / dmag
                dx, dy, dz = dx * sc, dy * sc, dz * sc
            nx_, ny_, nz_ = px + dx, py + dy, pz + dz
            # check_particle_boundary_violation_safe
            violation = False
            if not _valid_coordinate(nx_, ny_, nz_, max_coord):
                violation = True
            elif nz_ < bz - f(1.0):
                violation = True
            else:
                ddx, ddy = nx_ - cx0, ny_ - cy0
                d2 = ddx * ddx + ddy * ddy
                if d2 > f(1e6):
                    violation = True
                else:
                    dist = np.sqrt(d2)
                    hd = nz_ - bz
                    if hd >= 0 and hd < cup:
                        hr = max(f(0.0), min(f(1.0), hd / cup))
                        max_r = br + (tr - br) * hr
                        if dist > max_r * f(0.9):
                            violation = True
                    elif hd >= cup:
                        if dist > tr * f(0.9):
                            violation = True
            if violation:
                # constrain_to_boundary_safe
                cx_, cy_, cz_ = nx_, ny_, nz_
                if cz_ < bz:
                    cz_ = bz + f(0.1)
                max_z = min(bz + cup * f(1.5), nz5)
                if cz_ > max_z:
                    cz_ = max_z - f(0.1)
                ddx, ddy = cx_ - cx0, cy_ - cy0
                d2 = ddx * ddx + ddy * ddy
                if d2 < f(1e6):
                    dist = np.sqrt(d2)
                    if dist > f(0.1):
                        hd = max(f(0.0), cz_ - bz)
                        max_r = tr
                        if hd < cup:
                            hr = max(f(0.0), min(f(1.0), hd / cup))
                            max_r = br + (tr - br) * hr
                        if dist > max_r * f(0.8):
                            sf = (max_r * f(0.8)) / dist
                            sf = max(f(0.1), min(f(1.0), sf))
                            cx_ = cx0 + ddx * sf
                            cy_ = cy0 + ddy * sf
                else:
                    cx_, cy_ = cx0, cy0
                if _valid_coordinate(cx_, cy_, cz_, max_coord):
                    nx_, ny_, nz_ = cx_, cy_, cz_
                    vx, vy, vz = vx * f(0.3), vy * f(0.3), vz * f(0.3)
                    viol += 1
                else:
                    nx_, ny_, nz_ = px, py, pz
                    vx = vy = vz = f(0.0); coord_err += 1
            if _valid_coordinate(nx_, ny_, nz_, max_coord):
                pos[i] = (nx_, ny_, nz_)
            else:
                active[i] = 0; coord_err += 1
            vel[i] = (vx, vy, vz)
            force[i] = 0
    return coord_err, viol


def add_particle_reaction_forces(st: State, reaction: np.ndarray) -> None:
    """LBMSolver.add_particle_reaction_forces, legacy/lbm_solver.py:1478-1483."""
    fluid = (st.solid == 0)[..., None]
    st.body_force = np.where(fluid, st.body_force + reaction, st.body_force).astype(F32)


# ==========================================================================
# compat = physical  (new capability; oracle for the periodic / TGV / roofline configs)
# ==========================================================================
@dataclass
class PhysParams:
    nx: int
    ny: int
    nz: int
    tau_water: float = 0.53
    tau_air: float = 0.8
    gravity_lu: float = 0.0
    periodic: tuple = (True, True, True)
    use_force: bool = False          # body_force + gravity*phase through standard Guo forcing
    use_phase: bool = False          # tau by phase; else tau_water everywhere
    les: bool = False                # local Pi^neq Smagorinsky
    cs_smag: float = 0.18
    tau_min: float = 0.55
    tau_max: float = 1.90
    porous: bool = False             # Guo-Zhao drag in filter-zone cells
    porous_darcy: float = 0.0        # nu/K          [1/ts]
    porous_forch: float = 0.0        # F_eps/sqrt(K) [1/lu]
    mrt_magic: float = 0.0           # 0 = BGK; > 0: two-rate MRT, (tau - 1/2)(tau_odd - 1/2) = mrt_magic (include/lbm_b200.h)


def equilibrium_phys(rho, ux, uy, uz, q: int):
    return equilibrium_ref(rho, ux, uy, uz, q, table="config")


def init_equilibrium_phys(rho0, u0):
    """g[q] = f_eq(rho0,u0) (consistent velocity set). rho0 [NX,NY,NZ], u0 [NX,NY,NZ,3]."""
    g = np.empty((Q,) + rho0.shape, F32)
    for q in range(Q):
        g[q] = equilibrium_phys(rho0, u0[..., 0], u0[..., 1], u0[..., 2], q)
    return g


def _pull(g_q, q, solid, g_opp, periodic, w_q):
    """Stream population q into place (pull form): value arriving at x comes from x-e_q.
    Source solid -> halfway bounce-back (own opposite population); source outside a
    non-periodic face -> w_q (the reference's stale-inflow rule)."""
    ex, ey, ez = int(CX[q]), int(CY[q]), int(CZ[q])
    src = np.roll(g_q, shift=(ex, ey, ez), axis=(0, 1, 2))
    if solid is not None:
        src_solid = np.roll(solid, shift=(ex, ey, ez), axis=(0, 1, 2)) != 0
    else:
        src_solid = np.zeros(g_q.shape, bool)
    oob = np.zeros(g_q.shape, bool)
    for ax, e in enumerate((ex, ey, ez)):
        if e != 0 and not periodic[ax]:
            sl = [slice(None)] * 3
            sl[ax] = 0 if e > 0 else -1
            oob[tuple(sl)] = True
    out = np.where(src_solid & ~oob, g_opp, src)
    out = np.where(oob, w_q, out)
    return out.astype(F32)


# opposite-direction pairs (p, pbar), e_pbar = -e_p: (1,2) (3,4) (5,6) (7,10) (9,8) (11,14) (13,12) (15,18) (17,16)
PAIR_P = (1, 3, 5, 7, 9, 11, 13, 15, 17)
PAIR_M = (2, 4, 6, 10, 8, 14, 12, 18, 16)

# f32 lattice constants of the physical operator (csrc/lbm_phys.cuh:phys_const): f32 products of the f32 weights
W1X2 = F32(2.0) * W[1]; W2X2 = F32(2.0) * W[7]
W1X6 = F32(6.0) * W[1]; W2X6 = F32(6.0) * W[7]
W1X18 = F32(18.0) * W[1]; W2X18 = F32(18.0) * W[7]


def _fma(a, b, c):
    """round(a*b + c) with ONE rounding (C99 fmaf through oracle/ref_cpu.c; NumPy has no fused operation)."""
    from . import ref_cpu
    return ref_cpu.fmaf(a, b, c)


def _vedot(ex: int, ey: int, ez: int, vx, vy, vz):
    """e . v for the + member of a pair: x, y, z order, one rounding per add/sub (lbm_phys.cuh:vedot)."""
    acc = None
    for e, v in ((ex, vx), (ey, vy), (ez, vz)):
        if e == 0:
            continue
        if acc is None:
            assert e > 0
            acc = v
        else:
            acc = (acc + v) if e > 0 else (acc - v)
    return acc


def step_physical(g, p: PhysParams, solid=None, body_force=None, phase=None, filter_zone=None,
                  les_mask=None):
    """One fused pull step on post-collision populations g[q,i,j,k].
    Returns (g_next, rho, u) with rho,u the moments of the streamed (pre-collision) state.
    Standard BGK + Guo forcing (Guo, Zheng, Shi 2002) + local Smagorinsky (Hou et al. 1996)
    + Guo-Zhao (2002) porous drag; see SURVEY.md A.3 last paragraph.

    This mode is a NEW capability (the reference has no consistent-lattice / periodic step), so its arithmetic is
    ours to define.  The contract, shared operation by operation with csrc/lbm_phys.cuh:collide_phys: every step is
    an explicit IEEE f32 add / sub / mul / fused multiply-add (`_fma`), a correctly rounded reciprocal (1/x) or a
    correctly rounded square root.  The operator works on the pair sums s_k = f_p + f_m and differences
    d_k = f_p - f_m of opposite directions, and the Smagorinsky stress is the second moment of f minus its
    equilibrium value rho (I/3 + u u).  The CUDA kernels are bit-exact against this function.
    g_next is defined on fluid cells; solid cells keep g (on the device their slots are bounce-back scratch)."""
    one = F32(1.0); half = F32(0.5)
    f = [None] * Q
    for q in range(Q):
        f[q] = _pull(g[q], q, solid, g[int(OPP[q])], p.periodic, W[q])
    s = [f[PAIR_P[k]] + f[PAIR_M[k]] for k in range(9)]
    d = [f[PAIR_P[k]] - f[PAIR_M[k]] for k in range(9)]
    rho = f[0]
    for k in range(9):
        rho = rho + s[k]
    mx = (((d[0] + d[3]) + d[4]) + d[5]) + d[6]
    my = (((d[1] + d[3]) - d[4]) + d[7]) + d[8]
    mz = (((d[2] + d[5]) - d[6]) + d[7]) - d[8]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv_rho = (one / rho).astype(F32)
    zero = np.zeros(rho.shape, F32)
    forced = False
    Fx, Fy, Fz = zero, zero, zero
    has_phase = phase is not None and p.use_phase
    has_force = (p.use_force and body_force is not None) or (has_phase and p.gravity_lu != 0.0)
    if has_force:
        forced = True
        if p.use_force and body_force is not None:
            Fx = body_force[..., 0].astype(F32); Fy = body_force[..., 1].astype(F32); Fz = body_force[..., 2].astype(F32)
        if has_phase and p.gravity_lu != 0.0:
            Fz = _fma(-F32(p.gravity_lu), phase.astype(F32), Fz)
        ux = _fma(half, Fx, mx) * inv_rho
        uy = _fma(half, Fy, my) * inv_rho
        uz = _fma(half, Fz, mz) * inv_rho
    else:
        ux = mx * inv_rho; uy = my * inv_rho; uz = mz * inv_rho
    if p.porous:
        zone = (filter_zone != 0)
        darcy = F32(p.porous_darcy); forch = F32(p.porous_forch)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            vmag = np.sqrt(_fma(uz, uz, _fma(uy, uy, ux * ux)))
            c0 = half * _fma(half, darcy, one)
            c1 = half * forch
            den = c0 + np.sqrt(_fma(c1, vmag, c0 * c0))
            sc = (one / den).astype(F32)
            vx = ux * sc; vy = uy * sc; vz = uz * sc
            umag = vmag * sc
            cdrag = _fma(forch, umag, darcy)
            cr = -(cdrag * rho)
            ddx = _fma(cr, vx, Fx); ddy = _fma(cr, vy, Fy); ddz = _fma(cr, vz, Fz)
        ux = np.where(zone, vx, ux).astype(F32); uy = np.where(zone, vy, uy).astype(F32); uz = np.where(zone, vz, uz).astype(F32)
        Fx = np.where(zone, ddx, Fx).astype(F32)
        Fy = np.where(zone, ddy, Fy).astype(F32)
        Fz = np.where(zone, ddz, Fz).astype(F32)
        forced = True
    if has_phase:
        tau0 = np.where(phase > half, F32(p.tau_water), F32(p.tau_air)).astype(F32)
    else:
        tau0 = np.full(rho.shape, F32(p.tau_water), F32)
    tau = tau0
    if p.les:
        Mxx = (((s[0] + s[3]) + s[4]) + s[5]) + s[6]
        Myy = (((s[1] + s[3]) + s[4]) + s[7]) + s[8]
        Mzz = (((s[2] + s[5]) + s[6]) + s[7]) + s[8]
        Mxy = s[3] - s[4]; Mxz = s[5] - s[6]; Myz = s[7] - s[8]
        nr = F32(-1.0) * rho
        nrux = nr * ux; nruy = nr * uy
        third = W[0]
        pxx = _fma(nr, _fma(ux, ux, third), Mxx)
        pyy = _fma(nr, _fma(uy, uy, third), Myy)
        pzz = _fma(nr, _fma(uz, uz, third), Mzz)
        pxy = _fma(nrux, uy, Mxy); pxz = _fma(nrux, uz, Mxz); pyz = _fma(nruy, uz, Myz)
        qa = _fma(pzz, pzz, _fma(pyy, pyy, pxx * pxx))
        qb = _fma(pyz, pyz, _fma(pxz, pxz, pxy * pxy))
        qsum = _fma(F32(2.0), qb, qa)
        cs = float(F32(p.cs_smag))      # the C ABI carries Cs as f32
        kk = F32(18.0 * np.sqrt(2.0) * cs * cs)
        with np.errstate(invalid="ignore"):
            qn = np.sqrt(qsum)
            arg = _fma(kk * qn, inv_rho, tau0 * tau0)
            tles = half * (tau0 + np.sqrt(arg))
        if les_mask is not None:
            tles = np.where(les_mask != 0, tles, tau0)
        tau = np.maximum(F32(p.tau_min), np.minimum(F32(p.tau_max), tles)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        omega = (one / tau).astype(F32)
    nom = F32(-1.0) * omega
    omega_d, nom_d = omega, nom
    if p.mrt_magic > 0.0:            # the pair differences (odd moments) relax at 1 / tau_odd
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            tau_d = _fma(F32(p.mrt_magic), (one / (tau + F32(-0.5))).astype(F32), half)
            omega_d = (one / tau_d).astype(F32)
        nom_d = F32(-1.0) * omega_d
    u_sq = _fma(uz, uz, _fma(uy, uy, ux * ux))
    base = _fma(F32(-1.5), u_sq, one)
    # "f - w rho (...)" is ONE fma with the negated prefactor: no product feeds an add / sub (lbm_phys.cuh header)
    nws1 = (-W1X2) * rho; nws2 = (-W2X2) * rho
    nwd1 = (-W1X6) * rho; nwd2 = (-W2X6) * rho
    out = [None] * Q
    f0 = _fma(nom, _fma((-W[0]) * rho, base, f[0]), f[0])
    if forced:
        pref = _fma(F32(-0.5), omega, one)
        uF3 = F32(3.0) * _fma(uz, Fz, _fma(uy, Fy, ux * Fx))
        f0 = _fma((-W[0]) * pref, uF3, f0)
        c18 = (W1X18 * pref, W2X18 * pref)
        pref_d = _fma(F32(-0.5), omega_d, one)
        c6 = (W1X6 * pref_d, W2X6 * pref_d)
        nc2 = ((-W1X2) * pref, (-W2X2) * pref)
    out[0] = f0
    for k in range(9):
        pp, pm = PAIR_P[k], PAIR_M[k]
        e = (int(CX[pp]), int(CY[pp]), int(CZ[pp]))
        a = 0 if k < 3 else 1
        eu = _vedot(*e, ux, uy, uz)
        A = _fma(F32(4.5) * eu, eu, base)
        ns = _fma(nws1 if a == 0 else nws2, A, s[k])
        nd = _fma(nwd1 if a == 0 else nwd2, eu, d[k])
        sp = _fma(nom, ns, s[k])
        dp = _fma(nom_d, nd, d[k])
        if forced:
            eF = _vedot(*e, Fx, Fy, Fz)
            sp = _fma(eu * eF, c18[a], sp)
            sp = _fma(nc2[a], uF3, sp)
            dp = _fma(eF, c6[a], dp)
        hs = half * sp
        out[pp] = _fma(half, dp, hs)
        out[pm] = _fma(F32(-0.5), dp, hs)
    fluid = (solid == 0) if solid is not None else None
    g_next = np.empty_like(g)
    for q in range(Q):
        g_next[q] = out[q] if fluid is None else np.where(fluid, out[q], g[q])
    if fluid is not None:
        rho = np.where(fluid, rho, F32(0.0)).astype(F32)
        ux = np.where(fluid, ux, F32(0.0)); uy = np.where(fluid, uy, F32(0.0)); uz = np.where(fluid, uz, F32(0.0))
    u = np.stack([ux, uy, uz], axis=-1).astype(F32)
    return g_next.astype(F32), rho.astype(F32), u


# --------------------------------------------------------------------------
# stream / un-stream between the reference's pre-collision `f` and the device's
# post-collision `g` (pure data movement; used to compare states in tests)
# --------------------------------------------------------------------------
def stream_from_post_collision(g, solid, w_fill=True):
    """f[q,x] = g[q,x-e] (fluid source) | g[opp q,x] (solid source) | w_q (source outside the box)."""
    f = np.empty_like(g)
    for q in range(Q):
        f[q] = _pull(g[q], q, solid, g[int(OPP[q])], (False, False, False), W[q])
    return f
