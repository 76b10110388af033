"""The plain kernels of csrc/lbm_aux.cu -- V60 geometry, flag packing, neighbour masks, exact f <-> g conversion, face density
writes, pressure-gradient and Forchheimer forces -- compiled by g++ and executed thread by thread on the CPU
(tests/emu/emu_aux.cpp), against the recorded runs of the reference and the oracle.  Includes BASELINE's first criterion
(solid / fluid flags bit-exact) at the reference's own 224^3 on product kernel source, and the whole legacy-compatible
pipeline lbm_pack_flags -> lbm_import_f -> lbm_step x N -> lbm_export_f as emulated product kernels against the 1000-step
recording.  Test infrastructure only."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from oracle import d3q19_ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _build(name, deps):
    return H.build_emu(name, deps)


@pytest.fixture(scope="module")
def aux():
    return _build("emu_aux", ["lbm_aux.cu", "lbm_phys.cuh", "lbm_common.cuh"])


@pytest.fixture(scope="module")
def stepper():
    return _build("emu_step_reference", ["lbm_step_kernel.cuh", "lbm_phys.cuh", "lbm_common.cuh"])


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _dims(n):
    return C.c_int(n), C.c_int(n), C.c_int(n)


def _geometry(aux, n):
    from pour_over_coffee_lbm_b200.config import LBMConfig
    cfg = LBMConfig(NX=n, NY=n, NZ=n)
    solid = np.zeros((n, n, n), np.uint8); zone = np.zeros((n, n, n), np.int32)
    geom = np.array(cfg.v60_geometry_constants(), np.float32)
    aux.emu_v60_geometry(*_dims(n), _p(solid), _p(zone), _p(geom))
    return np.transpose(solid, (2, 1, 0)), np.transpose(zone, (2, 1, 0))


@pytest.mark.parametrize("n", [16, 64, 224])
def test_emulated_v60_geometry_flags_bit_exact(aux, n):
    """BASELINE: "solid/fluid flags ... must match bit-exactly" -- v60_geometry_kernel against the recorded reference run
    (16^3) and the oracle's restatement of FilterPaperSystem._setup_v60_geometry / _setup_filter_zones (64^3, 224^3)."""
    solid, zone = _geometry(aux, n)
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    assert np.array_equal(solid, R.v60_solid(cfg)) and np.array_equal(zone, R.filter_zones(cfg))
    if n == 16:
        z = np.load(os.path.join(GOLD, "reference_run_neighbours.npz"))
        assert np.array_equal(solid, z["solid"]) and np.array_equal(zone, z["filter_zone"])
    if n == 224:
        assert 0.30 < float((solid == 0).mean()) < 0.40


def test_emulated_force_producers_reproduce_the_reference_run(aux):
    """pressure_gradient_kernel (force and mixed drive) and forchheimer_force_kernel against the recorded
    PressureGradientDrive / FilterPaperSystem.compute_forchheimer_resistance run."""
    from pour_over_coffee_lbm_b200.config import LBMConfig
    z = np.load(os.path.join(GOLD, "reference_run_neighbours.npz"))
    n = int(z["n"]); cfg = LBMConfig(NX=n, NY=n, NZ=n)
    flags = H.to_dev_scalar((z["solid"] | (2 * (z["filter_zone"] != 0))).astype(np.uint8))
    rho, u = H.to_dev_scalar(z["rho"]), H.to_dev_vec(z["u"])
    bf = np.zeros_like(u)
    aux.emu_pressure_gradient(*_dims(n), _p(rho), _p(flags), _p(bf), C.c_float(0.12), C.c_float(1.0), C.c_int(1))
    assert np.array_equal(np.transpose(bf, (3, 2, 1, 0)), z["bf_force_drive"])
    bf[:] = 0
    aux.emu_pressure_gradient(*_dims(n), _p(rho), _p(flags), _p(bf), C.c_float(0.12), C.c_float(0.5), C.c_int(1))
    assert np.array_equal(np.transpose(bf, (3, 2, 1, 0)), z["bf_mixed_drive"])
    for vec in (1, 2, 4):                                        # the tile-list variant (what a V60 run uses): same values, "set" mode included
        bt = np.full_like(u, 7.0)
        n_items = aux.emu_pressure_gradient_tiles(*_dims(n), C.c_int(vec), _p(rho), _p(flags), _p(bt), C.c_float(0.12), C.c_float(0.5), C.c_int(0))
        fluid_dev = H.to_dev_scalar(z["solid"]) == 0
        assert n_items > 0 and np.array_equal(bt[:, fluid_dev], bf[:, fluid_dev]) and np.all(bt[:, ~fluid_dev] == 7.0)
    k_lu, beta = cfg.forchheimer_parameters(); c_darcy, c_forch = cfg.filter_constants()
    aux.emu_forchheimer(*_dims(n), _p(u), _p(flags), _p(bf), C.c_float(k_lu), C.c_float(beta), C.c_float(c_darcy), C.c_float(c_forch),
                        C.c_float(0.01 * cfg.SCALE_VELOCITY / cfg.DT))
    assert np.array_equal(np.transpose(bf, (3, 2, 1, 0)), z["bf_mixed_plus_forchheimer"])


@pytest.mark.parametrize("name", ["reference_run_long_air_1000", "reference_run_step_water_default_gravity", "reference_run_openbox"])
def test_emulated_pipeline_pack_import_step_export(aux, stepper, name):
    """lbm_pack_flags (+ neighbour masks) -> lbm_import_f -> lbm_step x N (-> lbm_face_bc) -> lbm_export_f, every stage the product's
    kernel source, against the reference's recorded runs -- 1000 steps, the default-gravity scenario with all clamps
    saturated, and the open box with the boundary manager's face writes."""
    from pour_over_coffee_lbm_b200.config import LBMConfig
    z = np.load(os.path.join(GOLD, name + ".npz"))
    n, steps, gravity = int(z["n"]), int(z["steps"]), float(z["gravity"])
    open_box = "filter_zone" not in z.files
    c = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=gravity)
    cfg = LBMConfig(NX=n, NY=n, NZ=n, TAU_FLUID=c.TAU_WATER, TAU_AIR=c.TAU_AIR, GRAVITY_LU=gravity)
    k_lu, beta_lu = cfg.forchheimer_parameters(); c_darcy, c_forch = cfg.filter_constants()
    solid = H.to_dev_scalar(z["solid"]).astype(np.uint8)
    zone = H.to_dev_scalar(z["filter_zone"]).astype(np.int32) if not open_box else np.zeros((n, n, n), np.int32)
    les = H.to_dev_scalar(z["les_mask"]).astype(np.int32)
    flags = np.zeros((n, n, n), np.uint8); nbr = np.zeros((n, n, n), np.uint64)
    aux.emu_pack_flags_and_masks(*_dims(n), C.c_int(0), _p(flags), _p(solid), _p(zone), _p(les), _p(nbr))
    g = [np.empty((19, n, n, n), np.float32), np.empty((19, n, n, n), np.float32)]
    f_in = H.to_dev_pop(z["f"])
    aux.emu_convert_f(*_dims(n), C.c_int(0), _p(f_in), _p(flags), _p(g[0]))
    g[1][:] = g[0]
    force, phase = H.to_dev_vec(z["body_force"]), H.to_dev_scalar(z["phase"])
    rho = np.ones((n, n, n), np.float32); u = [np.zeros((3, n, n, n), np.float32), np.zeros((3, n, n, n), np.float32)]
    blockage = np.zeros((n, n, n), np.float32)
    f32 = lambda v: C.c_float(float(v))
    cur = 0
    for _ in range(steps):
        stepper.emu_step_reference(*_dims(n), _p(g[cur]), _p(g[1 - cur]), _p(rho), _p(u[cur]), _p(u[1 - cur]), _p(force), _p(phase), _p(blockage),
                                   _p(flags), _p(nbr), C.c_int(1), C.c_int(0 if open_box else 1), f32(cfg.TAU_WATER), f32(cfg.TAU_AIR), f32(gravity),
                                   f32(cfg.LES_CS), f32(0.55), f32(1.90), f32(k_lu), f32(beta_lu), f32(c_darcy), f32(c_forch))
        cur = 1 - cur
        if open_box:
            aux.emu_face_bc(*_dims(n), _p(rho), _p(flags))
    f_out = np.empty_like(g[cur])
    aux.emu_convert_f(*_dims(n), C.c_int(1), _p(g[cur]), _p(flags), _p(f_out))
    fluid = z["solid"] == 0
    assert np.array_equal(np.transpose(rho, (2, 1, 0))[fluid], z["rho"][fluid])
    assert np.array_equal(np.transpose(u[cur], (3, 2, 1, 0))[fluid], z["u"][fluid])
    assert np.array_equal(np.transpose(f_out, (0, 3, 2, 1))[:, fluid], z["f_out"][:, fluid])


@pytest.mark.parametrize("shape,seed", [((13, 9, 11), 1), ((8, 20, 6), 2), ((17, 5, 9), 3)])
def test_emulated_pipeline_on_random_ragged_boxes_matches_the_oracle(aux, stepper, shape, seed):
    """Edge cases no recording holds: non-cubic boxes, random solid masks (isolated fluid cells, solids on every face), random
    filter zones with a non-zero blockage field, random LES mask and phase, water-phase relaxation -- the emulated product
    pipeline against oracle.step, six steps, bit for bit on fluid cells."""
    from pour_over_coffee_lbm_b200.config import LBMConfig
    nx, ny, nz = shape
    rng = np.random.default_rng(seed)
    gravity = 3e-5
    cfg_o = R.RefConfig(NX=nx, NY=ny, NZ=nz, GRAVITY_LU=gravity)
    st = R.init_fields(cfg_o)
    st.solid = (rng.random(shape) < 0.3).astype(np.uint8)
    st.filter_zone = ((rng.random(shape) < 0.25) & (st.solid == 0)).astype(np.int32)
    st.filter_blockage = np.where(st.filter_zone == 1, rng.uniform(0, 0.6, shape), 0).astype(np.float32)
    st.K_lu, st.beta_lu = R.forchheimer_params(cfg_o); st.apply_filter = True
    st.les_mask = (rng.random(shape) < 0.7).astype(np.int32)
    st.phase = rng.choice([0.0, 0.3, 0.95, 1.0], shape).astype(np.float32)
    st.body_force = (2e-5 * rng.standard_normal(shape + (3,))).astype(np.float32)
    u0 = (0.01 * rng.standard_normal(shape + (3,))).astype(np.float32); rho0 = (1 + 0.01 * rng.standard_normal(shape)).astype(np.float32)
    for q in range(R.Q):
        st.f[q] = R.equilibrium_ref(rho0, u0[..., 0], u0[..., 1], u0[..., 2], q, "config")
        # a state reachable from init_fields (f = f_new = w_q): slots that only an inflow through a face could write were never
        # written, so they still hold w_q in both buffers (quirk Q6 -- the rule the neighbour masks' high word encodes)
        for ax, e in enumerate((int(R.CX[q]), int(R.CY[q]), int(R.CZ[q]))):
            if e != 0:
                sl = [slice(None)] * 3; sl[ax] = 0 if e > 0 else -1
                st.f[q][tuple(sl)] = R.W[q]
        st.f_new[q] = st.f[q]
    f0 = st.f.copy()
    steps = 6
    for _ in range(steps):
        R.step(st)
    cfg = LBMConfig(NX=nx, NY=ny, NZ=nz, TAU_FLUID=cfg_o.TAU_WATER, TAU_AIR=cfg_o.TAU_AIR, GRAVITY_LU=gravity)
    k_lu, beta_lu = float(st.K_lu), float(st.beta_lu); c_darcy, c_forch = cfg.filter_constants()
    dims = (C.c_int(nx), C.c_int(ny), C.c_int(nz))
    solid = H.to_dev_scalar(st.solid); zone = H.to_dev_scalar(st.filter_zone); les = H.to_dev_scalar(st.les_mask)
    flags = np.zeros((nz, ny, nx), np.uint8); nbr = np.zeros((nz, ny, nx), np.uint64)
    aux.emu_pack_flags_and_masks(*dims, C.c_int(0), _p(flags), _p(solid), _p(zone), _p(les), _p(nbr))
    g = [np.empty((19, nz, ny, nx), np.float32), None]
    aux.emu_convert_f(*dims, C.c_int(0), _p(H.to_dev_pop(f0)), _p(flags), _p(g[0])); g[1] = g[0].copy()
    force, phase, blockage = H.to_dev_vec(st.body_force), H.to_dev_scalar(st.phase), H.to_dev_scalar(st.filter_blockage)
    rho = np.ones((nz, ny, nx), np.float32); u = [np.zeros((3, nz, ny, nx), np.float32), np.zeros((3, nz, ny, nx), np.float32)]
    f32 = lambda v: C.c_float(float(v))
    cur = 0
    for _ in range(steps):
        stepper.emu_step_reference(*dims, _p(g[cur]), _p(g[1 - cur]), _p(rho), _p(u[cur]), _p(u[1 - cur]), _p(force), _p(phase), _p(blockage), _p(flags),
                                   _p(nbr), C.c_int(1), C.c_int(1), f32(cfg.TAU_WATER), f32(cfg.TAU_AIR), f32(gravity), f32(cfg.LES_CS), f32(0.55),
                                   f32(1.90), f32(k_lu), f32(beta_lu), f32(c_darcy), f32(c_forch))
        cur = 1 - cur
        aux.emu_face_bc(*dims, _p(rho), _p(flags))
    f_out = np.empty_like(g[cur])
    aux.emu_convert_f(*dims, C.c_int(1), _p(g[cur]), _p(flags), _p(f_out))
    fluid = st.solid == 0
    assert np.array_equal(np.transpose(u[cur], (3, 2, 1, 0))[fluid], st.u[fluid], equal_nan=True)
    assert np.array_equal(np.transpose(rho, (2, 1, 0))[fluid], st.rho[fluid], equal_nan=True)
    assert np.array_equal(np.transpose(f_out, (0, 3, 2, 1))[:, fluid], st.f[:, fluid], equal_nan=True)
    assert np.isfinite(st.u[fluid]).all() and (st.filter_zone == 1).sum() > 20


# ---- compat = physical behind walls: phys_walls_kernel<VEC = 1> + phys_finish (write-side bounce-back) ------------------------
@pytest.fixture(scope="module")
def walls():
    return _build("emu_step_walls_physical", ["lbm_phys.cuh", "lbm_common.cuh"])


def _walls_case(aux, walls, shape, periodic, solid, zone, les_mask, phase, bf, rho0, u0, steps, p):
    nx, ny, nz = shape
    dims = (C.c_int(nx), C.c_int(ny), C.c_int(nz))
    per = (1 if periodic[0] else 0) | (2 if periodic[1] else 0) | (4 if periodic[2] else 0)
    g = R.init_equilibrium_phys(rho0, u0)
    d_solid = H.to_dev_scalar(solid); d_zone = H.to_dev_scalar(zone.astype(np.int32)); d_les = H.to_dev_scalar(les_mask.astype(np.int32))
    flags = np.zeros((nz, ny, nx), np.uint8); nbr = np.zeros((nz, ny, nx), np.uint64)
    aux.emu_pack_flags_and_masks(*dims, C.c_int(per), _p(flags), _p(d_solid), _p(d_zone), _p(d_les), _p(nbr))
    b = [H.to_dev_pop(g), None]
    aux.emu_bounce_slots(*dims, C.c_int(per), _p(b[0]), _p(flags), _p(nbr))          # what the engine does before the first step
    b[1] = b[0].copy()
    rho = np.zeros((nz, ny, nx), np.float32); u = np.zeros((3, nz, ny, nx), np.float32)
    f32 = lambda v: C.c_float(float(v))
    cur = walls.emu_step_walls_physical(*dims, C.c_int(per), C.c_int(steps), _p(b[0]), _p(b[1]), _p(rho), _p(u),
                                        _p(H.to_dev_vec(bf)) if bf is not None else None, _p(H.to_dev_scalar(phase)) if phase is not None else None,
                                        _p(flags), _p(nbr), C.c_int(int(p.les)), C.c_int(int(p.porous)), f32(p.tau_water), f32(p.tau_air),
                                        f32(p.gravity_lu), f32(p.cs_smag), f32(p.tau_min), f32(p.tau_max), f32(p.porous_darcy), f32(p.porous_forch))
    for _ in range(steps):
        g, rho_o, u_o = R.step_physical(g, p, solid=solid, body_force=bf, phase=phase, filter_zone=zone, les_mask=les_mask)
    fluid = solid == 0
    assert np.array_equal(np.transpose(b[cur], (0, 3, 2, 1))[:, fluid], g[:, fluid])
    assert np.array_equal(np.transpose(rho, (2, 1, 0))[fluid], rho_o[fluid]) and np.array_equal(np.transpose(u, (3, 2, 1, 0))[fluid], u_o[fluid])


def test_emulated_physical_walls_kernel_v60_all_features(aux, walls):
    """The BASELINE configs[2] family at 32^3: V60 mask, force, phase, local-stress LES, Guo-Zhao drag in the filter zone; the
    emulated one-cell walls kernel (pull + write-side bounce-back) against oracle.step_physical, bit for bit, 12 steps."""
    n, steps = 32, 12
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    solid = R.v60_solid(cfg); zone = R.filter_zones(cfg); les_mask = np.where(zone == 1, 0, 1).astype(np.int32)
    rng = np.random.default_rng(11)
    phase = np.zeros((n, n, n), np.float32); phase[:, :, : int(0.6 * n)] = 1.0
    bf = (2e-5 * rng.standard_normal((n, n, n, 3))).astype(np.float32)
    u0 = H.smooth_velocity(n, 0.02, 11); rho0 = H.smooth_density(n, 0.01, 11)
    p = R.PhysParams(nx=n, ny=n, nz=n, tau_water=0.53, tau_air=0.8, gravity_lu=1e-5, periodic=(False, False, False), use_force=True,
                     use_phase=True, les=True, porous=True, porous_darcy=0.37, porous_forch=0.9)
    _walls_case(aux, walls, (n, n, n), (False, False, False), solid, zone, les_mask, phase, bf, rho0, u0, steps, p)


@pytest.mark.parametrize("periodic", [(True, True, True), (True, False, True), (False, False, False)])
def test_emulated_physical_walls_kernel_obstacles_on_open_and_periodic_faces(aux, walls, periodic):
    """Obstacles touching open and periodic faces of a ragged box (nx = 27: a partial warp-tile): wrap in the pull, wrap of the
    write-side bounce-back targets, w_q inflow on open faces."""
    nx, ny, nz, steps = 27, 10, 12, 8
    rng = np.random.default_rng(5)
    solid = np.zeros((nx, ny, nz), np.uint8)
    solid[0:3, 2:5, 3:6] = 1; solid[nx - 2:, 6:9, 0:2] = 1; solid[10:14, 0:2, nz - 2:] = 1; solid[12:15, 4:7, 5:8] = 1
    solid |= (rng.random((nx, ny, nz)) < 0.05).astype(np.uint8)
    zone = np.zeros((nx, ny, nz), np.int32); les_mask = np.ones((nx, ny, nz), np.int32)
    u0 = H.smooth_velocity(nx, 0.03, 8, nz=nz, ny=ny); rho0 = H.smooth_density(nx, 0.01, 8, nz=nz, ny=ny)
    p = R.PhysParams(nx=nx, ny=ny, nz=nz, tau_water=0.6, periodic=periodic, les=True)
    _walls_case(aux, walls, (nx, ny, nz), periodic, solid, zone, les_mask, None, None, rho0, u0, steps, p)


# ---- packed quad list + wall links of the four-cell walls kernel (csrc/lbm_phys_chord.cuh) ----------------------------------------
_CX = [0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0]
_CY = [0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1]
_CZ = [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]
_OPP = [0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15]


def _chord_lists(aux, solid_zyx, periodic):
    nz, ny, nx = solid_zyx.shape
    flags = np.zeros_like(solid_zyx); nbr = np.zeros(solid_zyx.shape, np.uint64)
    aux.emu_pack_flags_and_masks(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(periodic), _p(flags), _p(np.ascontiguousarray(solid_zyx)), None, None, _p(nbr))
    max_t, max_l = nz * (ny * (nx // 4) // 32 + 2), 18 * solid_zyx.size
    quads = np.zeros((max_t, 32), np.uint64); tl = np.zeros((max_t, 2), np.uint32); links = np.zeros(max_l, np.uint32); nl = C.c_int(0)
    tile_off = np.zeros(nz + 1, np.int32)
    nt = aux.emu_chord_lists(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(periodic), _p(flags), _p(nbr), _p(quads), C.c_int(max_t), _p(tl), _p(links),
                             C.c_int(max_l), C.byref(nl), _p(tile_off))
    assert nt >= 0
    return flags, nbr, quads[:nt], tl[:nt], links[:nl.value], tile_off


@pytest.mark.parametrize("case", ["v60_64", "random_40x12x9", "random_periodic_24x10x8", "two_chords_136x6x5"])
def test_emulated_quad_list_holds_every_active_quad_once_and_links_are_the_wall_links(aux, case):
    """build_chord_lists (lbm_aux.cu): the list is every quad whose 32-byte sector (the quad and its even / odd partner) holds a fluid
    cell, once, in memory order, 32 per tile, only the last tile of a plane padded with dead lanes; the adjacency bits say exactly
    when the neighbouring lane of the same tile holds the neighbouring quad of the same row; the links of a tile are exactly the
    (fluid cell, q) pairs of its lanes whose target x + e_q is a solid cell inside the box; each names the stage word the waiting value
    is put into (row opp(q), the solid neighbour's place or the lane's edge word) and the stage word the new value is taken from (row
    q, the cell's own word); wall_values_kernel fills the per-link buffer with the owning cell's own population q."""
    rng = np.random.default_rng(5)
    periodic = 0
    if case == "v60_64":
        solid = np.ascontiguousarray(np.transpose(R.v60_solid(R.RefConfig(NX=64, NY=64, NZ=64)), (2, 1, 0)))
    elif case == "random_40x12x9":
        solid = (rng.random((9, 12, 40)) < 0.35).astype(np.uint8)
    elif case == "random_periodic_24x10x8":
        solid = (rng.random((8, 10, 24)) < 0.3).astype(np.uint8); periodic = 7
    else:      # a row longer than one tile with a solid gap wider than a tile between two chords
        solid = np.ones((5, 6, 136 * 2), np.uint8); solid[:, :, 3:50] = 0; solid[:, :, 200:269] = 0; solid[2, 3, 120] = 0
    nz, ny, nx = solid.shape
    flags, nbr, quads, tl, links, tile_off = _chord_lists(aux, solid, periodic)
    fluid = solid == 0
    has_fluid = fluid.reshape(nz, ny, nx // 4, 4).any(-1)
    quad_active = has_fluid.copy()
    nq = nx // 4
    quad_active[:, :, 0:nq - nq % 2:2] |= has_fluid[:, :, 1:nq:2]
    quad_active[:, :, 1:nq:2] |= has_fluid[:, :, 0:nq - nq % 2:2]
    g = rng.random((19, nz, ny, nx)).astype(np.float32)
    LIVE, LEFT, RIGHT = 1 << 44, 1 << 45, 1 << 46
    per = [(periodic >> d) & 1 for d in range(3)]
    listed = []
    got_links = set()
    link_cells = []
    for t in range(len(quads)):
        z_tile = int(np.searchsorted(tile_off, t, side="right") - 1)
        for l in range(32):
            e = int(quads[t, l])
            if not e & LIVE:
                assert e == 0 and t == tile_off[z_tile + 1] - 1          # padding only at the end of a plane's last tile
                assert all(int(quads[t, m]) == 0 for m in range(l, 32))
                break
            q0, y, z = e & 0xfff, (e >> 12) & 0xffff, (e >> 28) & 0xffff
            assert z == z_tile and quad_active[z, y, q0]
            listed.append((z, y, q0))
            prev = int(quads[t, l - 1]) if l > 0 else 0
            nxt = int(quads[t, l + 1]) if l < 31 else 0
            same_row = lambda o: bool(o & LIVE) and ((o >> 12) & 0xffff, (o >> 28) & 0xffff) == (y, z)
            for dz in (-1, 0, 1):                                    # bits 47-54: neighbouring rows whose quad at this x brings nothing
                for dy in (-1, 0, 1):
                    if dz == 0 and dy == 0:
                        continue
                    bit = (dz + 1) * 3 + (dy + 1); bit -= bit > 4
                    ys, zs = y + dy, z + dz
                    outside = (not per[1] and not 0 <= ys < ny) or (not per[2] and not 0 <= zs < nz)
                    dead = outside or not has_fluid[zs % nz, ys % ny, q0]
                    assert bool((e >> (47 + bit)) & 1) == dead
            assert bool(e & LEFT) == (same_row(prev) and (prev & 0xfff) == q0 - 1)
            assert bool(e & RIGHT) == (same_row(nxt) and (nxt & 0xfff) == q0 + 1)
        lb, n = int(tl[t, 0]), int(tl[t, 1])
        for L in links[lb:lb + n]:
            L = int(L)
            dst, src = L & 0xfff, (L >> 12) & 0xfff
            q, w = divmod(src, 160)
            l, c = divmod(w, 4)
            assert w < 128 and 1 <= q < 19
            e = int(quads[t, l])
            assert e & LIVE
            q0, y, z = e & 0xfff, (e >> 12) & 0xffff, (e >> 28) & 0xffff
            x = 4 * q0 + c
            row, pos = divmod(dst, 160)
            assert row == _OPP[q]
            wn = c + _CX[q]
            if 0 <= wn <= 3 or (wn < 0 and e & LEFT) or (wn > 3 and e & RIGHT):
                assert pos == 4 * l + wn
            else:
                assert pos == 128 + l
            got_links.add((z, y, x, q))
            link_cells.append((q, z, y, x))
    want = [(int(z), int(y), int(q)) for z, y, q in zip(*np.nonzero(quad_active))]
    assert listed == want                                                 # every active quad once, in memory order
    assert tile_off[-1] == len(quads) and all(tile_off[z + 1] - tile_off[z] == -(-int(quad_active[z].sum()) // 32) for z in range(nz))
    per = [(periodic >> d) & 1 for d in range(3)]
    want_links = set()
    for z, y, x in zip(*np.nonzero(fluid)):
        for q in range(1, 19):
            xt, yt, zt = x + _CX[q], y + _CY[q], z + _CZ[q]
            if not per[0] and not 0 <= xt < nx or not per[1] and not 0 <= yt < ny or not per[2] and not 0 <= zt < nz:
                continue
            if solid[zt % nz, yt % ny, xt % nx]:
                want_links.add((int(z), int(y), int(x), q))
    assert got_links == want_links and len(links) == len(want_links)
    wall = np.full(len(links), -1.0, np.float32)
    aux.emu_wall_values(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(periodic), _p(g), _p(np.ascontiguousarray(quads)), _p(np.ascontiguousarray(tl)),
                        _p(links), _p(wall), C.c_int(len(quads)))
    order = [i for t in range(len(quads)) for i in range(int(tl[t, 0]), int(tl[t, 0]) + int(tl[t, 1]))]
    assert order == list(range(len(links)))
    assert np.array_equal(wall, np.array([g[c] for c in link_cells], np.float32))


def test_emulated_quad_list_pressure_gradient_equals_the_grid_kernel(aux):
    """pressure_gradient_chord_kernel (the producer over the four-cell kernel's quad list) against the recorded drive run."""
    z = np.load(os.path.join(GOLD, "reference_run_neighbours.npz"))
    n = int(z["n"])
    solid = np.ascontiguousarray(H.to_dev_scalar(z["solid"]).astype(np.uint8))
    flags, nbr, quads, tl, links, tile_off = _chord_lists(aux, solid, 0)
    rho, u = H.to_dev_scalar(z["rho"]), H.to_dev_vec(z["u"])
    for scale, key in ((1.0, "bf_force_drive"), (0.5, "bf_mixed_drive")):
        bf = np.zeros_like(u)
        aux.emu_pressure_gradient_chord(*_dims(n), _p(rho), _p(flags), _p(bf), C.c_float(0.12), C.c_float(scale), C.c_int(1), _p(np.ascontiguousarray(quads)),
                                        C.c_int(len(quads)))
        assert np.array_equal(np.transpose(bf, (3, 2, 1, 0)), z[key])


@pytest.mark.parametrize("shape,seed", [((20, 9, 7), 5), ((18, 10, 6), 6), ((13, 8, 5), 7)], ids=["vec4", "ragged_nx", "ragged_odd"])
def test_emulated_forchheimer_and_reaction_kernels_on_ragged_boxes(aux, shape, seed):
    """forchheimer_force_kernel<4 | 1> ((x-chunk, y, z) launch grid, flags of a quad as one word) against the oracle's restatement of
    FilterPaperSystem.compute_forchheimer_resistance on non-cubic boxes with random solids / filter zones, velocities large enough to hit
    the force clamp; add_reaction_kernel<4 | 1> (LBMSolver.add_particle_reaction_forces) against body_force + reaction on fluid cells."""
    nx, ny, nz = shape
    rng = np.random.default_rng(seed)
    cfg_o = R.RefConfig(NX=nx, NY=ny, NZ=nz)
    st = R.init_fields(cfg_o)
    st.solid = (rng.random(shape) < 0.3).astype(np.uint8)
    st.filter_zone = ((rng.random(shape) < 0.4) & (rng.random(shape) < 0.9)).astype(np.int32)          # some filter cells are solid too
    st.K_lu, st.beta_lu = R.forchheimer_params(cfg_o)
    st.u = (rng.choice([1e-9, 1e-4, 0.05], shape + (1,)) * rng.standard_normal(shape + (3,))).astype(np.float32)
    st.body_force = (1e-5 * rng.standard_normal(shape + (3,))).astype(np.float32)
    bf = H.to_dev_vec(st.body_force); u = H.to_dev_vec(st.u)
    flags = H.to_dev_scalar((st.solid | (2 * (st.filter_zone != 0))).astype(np.uint8))
    c_darcy, c_forch = R.filter_constants(cfg_o)
    before = st.body_force.copy()
    R.compute_forchheimer_resistance(st)
    changed = np.any(st.body_force != before, axis=-1)
    assert changed.any() and not changed[st.solid == 1].any() and not changed[st.filter_zone == 0].any()      # the scenario is not trivial
    aux.emu_forchheimer(C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(u), _p(flags), _p(bf), C.c_float(st.K_lu), C.c_float(st.beta_lu),
                        C.c_float(c_darcy), C.c_float(c_forch), C.c_float(0.01 * 0.01 / cfg_o.DT))
    assert np.array_equal(np.transpose(bf, (3, 2, 1, 0)), st.body_force)
    # reaction accumulation
    reaction = (1e-4 * rng.standard_normal(bf.shape)).astype(np.float32)
    want = bf.copy()
    fluid = (flags & 1) == 0
    want[:, fluid] = bf[:, fluid] + reaction[:, fluid]
    aux.emu_add_reaction(C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(reaction), _p(flags), _p(bf))
    assert np.array_equal(bf, want)
