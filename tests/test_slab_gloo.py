"""N>1 host logic on CPU: two gloo ranks own z-slabs of one periodic / walled box, exchange only the 5+5 outgoing
populations per interface (pour_over_coffee_lbm_b200.slab.exchange_halo, the torch.distributed mirror of the NCCL
exchange in liblbm_b200) and must reproduce the single-domain oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from oracle import d3q19_ref as R


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _slab_step(g_loc, p_loc, solid_loc):
    """oracle step on a slab with ghost planes: non-periodic in z locally, ghosts supply the neighbours."""
    g2, rho, u = R.step_physical(g_loc, p_loc, solid=solid_loc)
    return g2, rho, u


def _worker(rank, world, port, periodic_z, out, balanced=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pour_over_coffee_lbm_b200 import slab
    n, nzg, steps = 12, 16, 6
    u0 = H.smooth_velocity(n, 0.04, 13, nz=nzg); rho0 = H.smooth_density(n, 0.01, 13, nz=nzg)
    solid = np.zeros((n, n, nzg), np.uint8)
    if not periodic_z:
        solid[:, :, 0] = 1; solid[:, :, -1] = 1; solid[4:7, 4:7, 6:10] = 1      # walls + an obstacle across the interface
    g_glob = R.init_equilibrium_phys(rho0, u0)
    if balanced:      # unequal slab thicknesses: cuts at the prefix sums of the per-plane fluid count (slab.partition_z_balanced)
        part = slab.partition_z_balanced([(z + 1) ** 2 * float((solid[:, :, z] == 0).sum()) for z in range(nzg)], world, min_planes=2)[rank]
    else:
        part = slab.partition_z(nzg, world)[rank]
    z0, nz = part.z0, part.nz
    zs = np.arange(z0 - 1, z0 + nz + 1) % nzg                                      # owned + ghost planes (wrapped indices)
    g = torch.from_numpy(H.to_dev_pop(g_glob[:, :, :, zs]))                        # device layout [q, z, y, x]
    sol = solid[:, :, zs].copy()
    if not periodic_z:                                                             # ghosts outside the global box are solid
        if z0 == 0: sol[:, :, 0] = 1
        if z0 + nz == nzg: sol[:, :, -1] = 1
    p_loc = R.PhysParams(nx=n, ny=n, nz=nz + 2, tau_water=0.6, les=True, periodic=(True, True, False))
    for _ in range(steps):
        slab.exchange_halo(g, rank, world, periodic_z)
        g_log = H.from_dev_pop(g)
        g2, rho, u = R.step_physical(g_log, p_loc, solid=sol)
        g_new = H.to_dev_pop(g2)
        g[:, 1:-1] = torch.from_numpy(g_new)[:, 1:-1]                              # only owned planes are updated
    res = H.from_dev_pop(g[:, 1:-1])
    gathered = [None] * world
    dist.all_gather_object(gathered, (z0, res))
    if rank == 0:
        full = np.concatenate([r for _, r in sorted(gathered, key=lambda t: t[0])], axis=3)
        p = R.PhysParams(nx=n, ny=n, nz=nzg, tau_water=0.6, les=True, periodic=(True, True, periodic_z))
        gg = g_glob
        for _ in range(steps):
            gg, rho, u = R.step_physical(gg, p, solid=solid if not periodic_z else None)
        fluid = solid == 0
        out.put(bool(np.array_equal(full[:, fluid], gg[:, fluid])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("balanced", [False, True])
@pytest.mark.parametrize("periodic_z", [True, False])
def test_two_gloo_ranks_reproduce_single_domain(periodic_z, balanced):
    """balanced: work-balanced slabs of unequal thickness (11 + 5 planes here) exchange the same 5+5 planes."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, periodic_z, out, balanced)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(timeout=240)
    for p in procs:
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


# ---- multiphase producers on z-slabs: device kernel source (CPU-emulated) + slab.exchange_planes over gloo ----------------
def _mp_worker(rank, world, port, out, cuts):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pour_over_coffee_lbm_b200 import slab
    here = os.path.dirname(os.path.abspath(__file__))
    emu = C.CDLL(os.path.join(here, "emu", "_build", "libemu_producers.so"))
    z = np.load(os.path.join(here, "golden", "reference_run_multiphase.npz"))
    n = int(z["n"]); z0, nz = cuts[rank]
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    F = lambda v: C.c_float(float(v))

    def local(dev, ghost_from_neighbours=False):
        """[.., z, y, x] global array -> this slab's [.., nz + 2, y, x]; ghost planes start as zeros (the exchange must fill them)."""
        zax = dev.ndim - 3
        sl = [slice(None)] * dev.ndim; sl[zax] = slice(z0, z0 + nz)
        own = dev[tuple(sl)]
        pad = [(0, 0)] * dev.ndim; pad[zax] = (1, 1)
        return np.ascontiguousarray(np.pad(own, pad))

    phi, phi_new, mu = local(H.to_dev_scalar(z["phi"])), local(H.to_dev_scalar(z["phi_new_in"])), local(H.to_dev_scalar(z["mu"]))
    u, rho, bf = local(H.to_dev_vec(z["u"])), local(H.to_dev_scalar(z["rho"])), local(H.to_dev_vec(z["body_force"]))
    flags = local(H.to_dev_scalar(z["solid"]).astype(np.uint8))
    phase, curv = np.zeros_like(rho), np.zeros_like(rho)
    gphi, gmu, nrm, sf = (np.zeros_like(bf) for _ in range(4))
    dims = (C.c_int(4), C.c_int(n), C.c_int(n), C.c_int(nz), C.c_int(z0), C.c_int(n))
    tt = torch.from_numpy                                   # shares memory with the NumPy array: the exchange writes in place
    sigma, mob, dt = float(z["sigma"]), float(z["mobility"]), float(z["dt"])
    rw, ra = C.c_double(float(z["rho_water"])), C.c_double(float(z["rho_air"]))

    def chain(body_force):
        slab.exchange_planes(tt(phi), rank, world, False); slab.exchange_planes(tt(mu), rank, world, False)
        emu.emu_slab_gradients(*dims, P(phi), P(mu), P(gphi), P(gmu), P(nrm))
        slab.exchange_planes(tt(nrm), rank, world, False)
        emu.emu_slab_curvature_force(*dims, P(phi), P(rho), P(flags), P(gphi), P(nrm), P(curv), P(sf), P(body_force), F(sigma))

    def phase_step():
        emu.emu_slab_phase_field_step(*dims, P(phi), P(phi_new), P(mu), P(u), P(rho), P(phase), F(mob), F(dt), rw, ra)

    own = lambda a: a[..., 1:-1, :, :].copy()
    snaps = {}
    chain(bf)
    snaps["st"] = dict(grad_phi=own(gphi), grad_mu=own(gmu), normal=own(nrm), curvature=own(curv), surface_force=own(sf), body_force=own(bf))
    chain(None); phase_step()
    snaps["s1"] = dict(phi=own(phi), phi_new=own(phi_new), rho=own(rho), phase=own(phase), body_force=own(bf))
    chain(bf); phase_step()
    snaps["s2"] = dict(phi=own(phi), rho=own(rho), phase=own(phase), body_force=own(bf), surface_force=own(sf), curvature=own(curv))
    gathered = [None] * world
    dist.all_gather_object(gathered, (z0, snaps))
    if rank == 0:
        parts = [s for _, s in sorted(gathered, key=lambda t: t[0])]
        ok = True
        for stage in ("st", "s1", "s2"):
            for name in parts[0][stage]:
                full = np.concatenate([p[stage][name] for p in parts], axis=-3)
                want = z[f"{stage}_{name}"]
                want = H.to_dev_vec(want) if want.ndim == 4 else H.to_dev_scalar(want)
                ok &= bool(np.array_equal(full, want))
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cuts", [((0, 8), (8, 8)), ((0, 5), (5, 11))], ids=["equal", "unequal"])
def test_two_gloo_ranks_multiphase_producers_reproduce_the_reference_run(cuts):
    """The surface-tension chain and the phase-field step on two z-slabs (ghost planes of phi, mu and -- between the two
    launches -- normal filled by slab.exchange_planes over gloo): the gathered slabs equal the recorded single-domain run of
    the reference bit for bit.  The device kernels run as CPU-emulated source (tests/emu)."""
    H.build_emu("emu_producers", ["lbm_producers.cu", "lbm_common.cuh"])
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mp_worker, args=(r, 2, port, out, cuts)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(timeout=240)
    for p in procs:
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


# ---- particle coupling on z-slabs: replicated particles, owner computes (emulated kernel source + slab helpers over gloo) ----
def _particle_worker(rank, world, port, out, cuts):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pour_over_coffee_lbm_b200 import slab
    here = os.path.dirname(os.path.abspath(__file__))
    emu = C.CDLL(os.path.join(here, "emu", "_build", "libemu_particles.so"))
    z = np.load(os.path.join(here, "golden", "reference_run_neighbours.npz"))
    n = int(z["n"]); z0, nz = cuts[rank]
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    P = lambda a: a.ctypes.data_as(C.c_void_p)

    class Particles(C.Structure):
        _fields_ = [(k, C.c_void_p) for k in ("pos", "vel", "radius", "mass", "active", "drag_new", "drag_old", "drag", "u_fluid", "reynolds", "cd",
                                              "cell")] + [("n", C.c_int)]
    npart = z["p_pos"].shape[0]
    t = lambda a: np.ascontiguousarray(a.T.astype(np.float32))
    st = dict(pos=t(z["p_pos"]), vel=t(z["p_vel"]), radius=z["p_radius"].copy(), mass=z["p_mass"].copy(), active=z["p_active"].astype(np.int32).copy(),
              drag_new=np.zeros((3, npart), np.float32), drag_old=t(z["p_drag_old_in"]), drag=np.zeros((3, npart), np.float32),
              u_fluid=np.zeros((3, npart), np.float32), reynolds=np.zeros(npart, np.float32), cd=np.zeros(npart, np.float32),
              cell=np.zeros((3, npart), np.int32))
    # this slab's u with one ghost plane per side: owned planes from the global field, ghosts filled by the exchange
    u_glob = H.to_dev_vec(z["u"])
    u = np.ascontiguousarray(np.pad(u_glob[:, z0:z0 + nz], ((0, 0), (1, 1), (0, 0), (0, 0))))
    slab.exchange_planes(torch.from_numpy(u), rank, world, False)
    # owner computes: the kernel sees the other slabs' particles as inactive
    owned = slab.particle_owner_mask(torch.from_numpy(st["pos"][2]), torch.from_numpy(st["active"]), z0, nz, n)
    masked = owned.numpy().copy()
    s = Particles(*[P(masked if k == "active" else st[k]) for k in ("pos", "vel", "radius", "mass", "active", "drag_new", "drag_old", "drag", "u_fluid",
                                                                      "reynolds", "cd", "cell")], npart)
    react = np.zeros_like(u)
    emu.emu_particles_couple_slab(C.c_int(n), C.c_int(n), C.c_int(nz), C.c_int(z0), C.c_int(n), P(u), P(react), C.byref(s),
                                  C.c_float(np.float32(cfg.WATER_DENSITY_90C)), C.c_float(np.float32(cfg.WATER_VISCOSITY_90C * cfg.WATER_DENSITY_90C)),
                                  C.c_float(0.8))
    slab.reduce_ghost_up(torch.from_numpy(react), rank, world, False)
    outs = [torch.from_numpy(st[k]) for k in ("drag_new", "drag_old", "drag", "u_fluid", "reynolds", "cd", "cell")]
    packed = [o.clone() for o in outs]
    slab.allreduce_owned(outs, owned, torch.from_numpy(st["active"]))
    slab.allreduce_owned_packed(packed, owned, torch.from_numpy(st["active"]))        # one collective instead of seven: same bits
    assert all(torch.equal(a, b) for a, b in zip(outs, packed))      # (the float sum turns the owner's -0.0 into +0.0, the packed one keeps its bits)
    gathered = [None] * world
    dist.all_gather_object(gathered, (z0, react[:, 1:-1].copy(), {k: st[k].copy() for k in ("drag_new", "drag_old", "drag", "u_fluid", "reynolds", "cd", "cell")},
                                      int((masked != 0).sum())))
    if rank == 0:
        parts = sorted(gathered, key=lambda g: g[0])
        react_full = np.transpose(np.concatenate([p[1] for p in parts], axis=1), (3, 2, 1, 0))
        act = z["p_active"] != 0
        ok = all(all(np.array_equal(p[2][k], parts[0][2][k]) for k in parts[0][2]) for p in parts)            # replicated state is consistent
        o = parts[0][2]
        ok &= np.array_equal(o["u_fluid"].T[act], z["p_u_fluid"][act]) and np.array_equal(o["reynolds"][act], z["p_reynolds"][act])
        ok &= bool(np.allclose(o["drag_new"].T[act], z["p_drag_new"][act], rtol=1e-6, atol=0))
        ok &= bool(np.allclose(o["drag"].T[act], z["p_drag"][act], rtol=1e-6, atol=1e-16) and np.allclose(o["drag_old"].T[act], z["p_drag_old_out"][act], rtol=1e-6, atol=1e-16))
        ok &= np.array_equal(o["cell"].T[act], np.stack(R.particle_cell_and_weights(cfg, z["p_pos"])[:3], 1)[act])
        ok &= bool(np.allclose(react_full, z["p_reaction"], rtol=1e-5, atol=1e-12))
        ok &= sum(p[3] for p in parts) == int(act.sum()) and all(p[3] > 0 for p in parts)                       # every active particle has exactly one owner
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cuts", [((0, 8), (8, 8)), ((0, 11), (11, 5))], ids=["equal", "unequal"])
def test_two_gloo_ranks_particle_coupling_reproduces_the_reference_run(cuts):
    """Two-way coupling on two z-slabs with replicated particles: each rank runs the coupling kernel (CPU-emulated source) on
    the particles whose base cell it owns (slab.particle_owner_mask hands the kernel a masked `active` array -- the kernel is
    unchanged), gathers u through a ghost plane, deposits reaction into its top ghost plane; slab.reduce_ghost_up moves that
    plane to the rank above, slab.allreduce_owned makes the per-particle outputs identical everywhere.  Result = the recorded
    single-domain run of the reference (scatter sums within rounding)."""
    H.build_emu("emu_particles", ["lbm_particles.cu", "lbm_common.cuh"])
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_particle_worker, args=(r, 2, port, out, cuts)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(timeout=240)
    for p in procs:
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


# ---- apply_fluid_forces / block_particles_at_filter on z-slabs: the product's slab functions over gloo, kernels CPU-emulated -------------------
def _producer_worker(rank, world, port, out, cuts):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pour_over_coffee_lbm_b200 import engine as E
    here = os.path.dirname(os.path.abspath(__file__))
    emu_p = C.CDLL(os.path.join(here, "emu", "_build", "libemu_particles.so"))
    emu_f = C.CDLL(os.path.join(here, "emu", "_build", "libemu_producers.so"))
    n = 16
    z0, nz = cuts[rank]
    rng = np.random.default_rng(23)
    cfg = R.RefConfig(NX=n, NY=n, NZ=n)
    sl = float(np.float32(cfg.SCALE_LENGTH))
    u_glob = (0.05 * rng.standard_normal((3, n, n, n))).astype(np.float32)                   # [3][z][y][x]
    flags_glob = np.where(rng.random((n, n, n)) < 0.25, 2, 0).astype(np.uint8)               # LBM_FLAG_FILTER on a quarter of the cells
    npart = 600

    def particles(scaled):
        r = np.random.default_rng(29)
        ps = E.ParticleState(npart, "cpu")
        pos = r.uniform(0.5, n - 1.5, (3, npart))
        if scaled:
            pos = pos * sl
        else:
            pos[0, ::41] = np.nan; pos[2, 7::43] = 1e9                                       # invalid: deactivated and counted by the kernel
        ps.pos.copy_(torch.from_numpy(pos.astype(np.float32)))
        ps.vel.copy_(torch.from_numpy((0.02 * r.standard_normal((3, npart))).astype(np.float32)))
        rad = np.clip(r.normal(3.25e-4, 1e-4, npart), 1.6e-4, 4.9e-4).astype(np.float32)
        ps.radius.copy_(torch.from_numpy(rad))
        ps.mass.copy_(torch.from_numpy(((np.float32(4 / 3) * np.float32(3.14159)) * rad ** 3 * np.float32(1200.0)).astype(np.float32)))
        ps.active.fill_(1); ps.active[::13] = 0
        return ps

    pad = lambda a, ax: np.ascontiguousarray(np.pad(a, [(1, 1) if i == ax else (0, 0) for i in range(a.ndim)]))

    class Lib:      # the two C-ABI entry points the slab functions call, on the CPU-emulated kernel source
        def lbm_particles_fluid_forces(self, ctx, u, st, force, rho_w, mu_w, grav, counters, stream):
            return emu_p.emu_particles_fluid_forces_slab(n, n, nz, z0, n, u, st, force, C.c_double(rho_w), C.c_double(mu_w), C.c_double(grav), counters)

        def lbm_particles_block_at_filter(self, ctx, st, flags, acc, scale, noise, seed, stream):
            return emu_f.emu_particles_block_at_filter_slab(n, n, nz, z0, n, st, flags, acc, C.c_float(scale), C.c_float(noise), C.c_uint(seed))

    class Eng:
        lib, _ctx, stream = Lib(), None, None
        nx = ny = n
        zghost = 1
        def _check(self, rc, what): assert rc == 0, what
    eng = Eng()
    eng.z0, eng.nz, eng.nz_global, eng.rank, eng.nranks = z0, nz, n, rank, world
    # ghost planes: the neighbour's boundary planes (flags), zeros for u (the kernel reads the base cell only)
    fl = np.zeros((nz + 2, n, n), np.uint8); fl[1:-1] = flags_glob[z0:z0 + nz]
    if z0 > 0: fl[0] = flags_glob[z0 - 1]
    if z0 + nz < n: fl[-1] = flags_glob[z0 + nz]
    eng.flags = torch.from_numpy(fl)
    eng.u = torch.from_numpy(pad(u_glob[:, z0:z0 + nz], 1))

    pa, pb = particles(False), particles(True)
    force = torch.zeros_like(pa.pos); counters = torch.zeros(2, dtype=torch.int32)
    E.particles_fluid_forces_slab(eng, pa, force, counters, 997.0, 1.0e-3, 9.81)
    acc = torch.zeros((nz + 2, n, n), dtype=torch.float32)
    E.particles_block_at_filter_slab(eng, pb, acc, sl, 0.01, 7)
    gathered = [None] * world
    dist.all_gather_object(gathered, (z0, acc[1:-1].numpy().copy(), force.numpy().copy(), pa.vel.numpy().copy(), pa.active.numpy().copy(), counters.numpy().copy(),
                                      pb.vel.numpy().copy()))
    if rank == 0:
        parts = sorted(gathered, key=lambda g: g[0])
        ok = all(np.array_equal(p[k].view(np.int32), parts[0][k].view(np.int32)) for p in parts for k in (2, 3, 4, 5, 6))      # replicated state identical
        # the single-domain kernels on the whole box
        ra, rb = particles(False), particles(True)
        rforce = np.zeros((3, npart), np.float32); rcount = np.zeros(2, np.int32)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        st = ra.struct()
        emu_p.emu_particles_fluid_forces(n, n, n, P(u_glob), C.byref(st), P(rforce), C.c_double(997.0), C.c_double(1.0e-3), C.c_double(9.81), P(rcount))
        racc = np.zeros((n, n, n), np.float32)
        emu_f.emu_particles_block_at_filter(n, n, n, npart, C.c_void_p(rb.pos.data_ptr()), C.c_void_p(rb.vel.data_ptr()), C.c_void_p(rb.active.data_ptr()),
                                            P(flags_glob), P(racc), C.c_float(sl), C.c_float(0.01), C.c_uint(7))
        act = ra.active.numpy() != 0
        same = lambda a, b: np.array_equal(np.ascontiguousarray(a).view(np.int32), np.ascontiguousarray(b).view(np.int32))
        ok &= same(parts[0][2][:, act], rforce[:, act]) and same(parts[0][3], ra.vel.numpy()) and same(parts[0][4], ra.active.numpy()) and same(parts[0][5], rcount)
        ok &= same(parts[0][6], rb.vel.numpy()) and same(np.concatenate([p[1] for p in parts], axis=0), racc)
        ok &= int(rcount[0]) > 0 and float(racc.sum()) > 0.05 and int((rb.vel.numpy() != particles(True).vel.numpy()).any(0).sum()) > 5
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cuts", [((0, 8), (8, 8)), ((0, 11), (11, 5))], ids=["equal", "unequal"])
def test_two_gloo_ranks_particle_producers_on_slabs_equal_the_single_domain_kernels(cuts):
    """engine.particles_fluid_forces_slab (owner of the base cell computes; force, reset velocities, deactivations and the error counter
    all-reduced) and engine.particles_block_at_filter_slab (the ranks agree on the first filter plane of gz - 2 .. gz + 2, a range that
    straddles the cut for many particles here; its owner computes) -- the product's own functions over gloo, their two C-ABI calls routed to
    the CPU-emulated kernel source -- against the same kernels on the undivided box: every array bit for bit."""
    H.build_emu("emu_particles", ["lbm_particles.cu", "lbm_common.cuh"])
    H.build_emu("emu_producers", ["lbm_producers.cu", "lbm_common.cuh"])
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_producer_worker, args=(r, 2, port, out, cuts)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(timeout=240)
    for p in procs:
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


# ---- the legacy-compatible step on z-slabs: product kernel source (CPU-emulated) + slab.exchange_halo over gloo ---------------
def _step_worker(rank, world, port, out, cuts, fixture):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pour_over_coffee_lbm_b200 import slab
    from pour_over_coffee_lbm_b200.config import LBMConfig
    here = os.path.dirname(os.path.abspath(__file__))
    aux = C.CDLL(os.path.join(here, "emu", "_build", "libemu_aux.so")); stepper = C.CDLL(os.path.join(here, "emu", "_build", "libemu_step_reference.so"))
    z = np.load(os.path.join(here, "golden", fixture))
    n, steps, gravity = int(z["n"]), int(z["steps"]), float(z["gravity"])
    z0, nz = cuts[rank]
    c = R.RefConfig(NX=n, NY=n, NZ=n, GRAVITY_LU=gravity)
    cfg = LBMConfig(NX=n, NY=n, NZ=n, TAU_FLUID=c.TAU_WATER, TAU_AIR=c.TAU_AIR, GRAVITY_LU=gravity)
    k_lu, beta_lu = cfg.forchheimer_parameters(); c_darcy, c_forch = cfg.filter_constants()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    f32 = lambda v: C.c_float(float(v))

    def local(dev):
        """global [.., z, y, x] -> this slab's [.., nz + 2, y, x] with the neighbours' planes as ghosts (zeros outside the box)"""
        zax = dev.ndim - 3
        pad = [(0, 0)] * dev.ndim; pad[zax] = (1, 1)
        full = np.pad(dev, pad)
        sl = [slice(None)] * dev.ndim; sl[zax] = slice(z0, z0 + nz + 2)
        return np.ascontiguousarray(full[tuple(sl)])

    solid = local(H.to_dev_scalar(z["solid"]).astype(np.uint8)); zone = local(H.to_dev_scalar(z["filter_zone"]).astype(np.int32))
    les = local(H.to_dev_scalar(z["les_mask"]).astype(np.int32))
    dims = (C.c_int(n), C.c_int(n), C.c_int(nz), C.c_int(z0), C.c_int(n))
    flags = np.zeros_like(solid); nbr = np.zeros(solid.shape, np.uint64)
    aux.emu_slab_pack_flags_and_masks(*dims, P(flags), P(solid), P(zone), P(les), P(nbr))
    g = [np.empty((19, nz + 2, n, n), np.float32), None]
    aux.emu_slab_convert_f(*dims, C.c_int(0), P(local(H.to_dev_pop(z["f"]))), P(flags), P(g[0])); g[1] = g[0].copy()
    force, phase = local(H.to_dev_vec(z["body_force"])), local(H.to_dev_scalar(z["phase"]))
    rho = np.ones((nz + 2, n, n), np.float32); u = [np.zeros((3, nz + 2, n, n), np.float32), np.zeros((3, nz + 2, n, n), np.float32)]
    blockage = np.zeros((nz + 2, n, n), np.float32)
    cur = 0
    for _ in range(steps):
        slab.exchange_halo(torch.from_numpy(g[cur]), rank, world, False, vec3=torch.from_numpy(u[cur]))   # 5 + 5 populations, u for the FD-LES
        stepper.emu_step_reference_slab(*dims, P(g[cur]), P(g[1 - cur]), P(rho), P(u[cur]), P(u[1 - cur]), P(force), P(phase), P(blockage), P(flags),
                                        P(nbr), C.c_int(1), C.c_int(1), f32(cfg.TAU_WATER), f32(cfg.TAU_AIR), f32(gravity), f32(cfg.LES_CS), f32(0.55),
                                        f32(1.90), f32(k_lu), f32(beta_lu), f32(c_darcy), f32(c_forch))
        cur = 1 - cur
    slab.exchange_halo(torch.from_numpy(g[cur]), rank, world, False)
    f_out = np.empty_like(g[cur])
    aux.emu_slab_convert_f(*dims, C.c_int(1), P(g[cur]), P(flags), P(f_out))
    gathered = [None] * world
    dist.all_gather_object(gathered, (z0, rho[1:-1].copy(), u[cur][:, 1:-1].copy(), f_out[:, 1:-1].copy()))
    if rank == 0:
        parts = sorted(gathered, key=lambda t: t[0])
        rho_f = np.transpose(np.concatenate([p[1] for p in parts], 0), (2, 1, 0))
        u_f = np.transpose(np.concatenate([p[2] for p in parts], 1), (3, 2, 1, 0))
        f_f = np.transpose(np.concatenate([p[3] for p in parts], 1), (0, 3, 2, 1))
        fluid = z["solid"] == 0
        out.put(bool(np.array_equal(rho_f[fluid], z["rho"][fluid]) and np.array_equal(u_f[fluid], z["u"][fluid]) and
                     np.array_equal(f_f[:, fluid], z["f_out"][:, fluid])))
    dist.barrier()
    dist.destroy_process_group()


_SLAB_RUNS = [("reference_run_step_split_phase_small_gravity.npz", ((0, 8), (8, 8)), "les_active_equal"),
              ("reference_run_step_water_default_gravity.npz", ((0, 5), (5, 11)), "default_gravity_unequal"),
              ("reference_run_long_air_1000.npz", ((0, 9), (9, 7)), "1000_steps_unequal")]
if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run_long_air_1000_n20.npz")):
    _SLAB_RUNS.append(("reference_run_long_air_1000_n20.npz", ((0, 13), (13, 7)), "1000_steps_20cubed_unequal"))


@pytest.mark.parametrize("fixture,cuts", [r[:2] for r in _SLAB_RUNS], ids=[r[2] for r in _SLAB_RUNS])
def test_two_gloo_ranks_legacy_step_kernel_reproduces_the_reference_run(fixture, cuts):
    """The product's legacy-compatible step kernel source (CPU-emulated, slab geometry: one ghost plane per side, global-z tests)
    on two z-slabs with slab.exchange_halo (5 + 5 outgoing populations per interface, u planes for the lagged FD-LES) over gloo:
    the gathered result equals the reference's own recorded single-domain runs -- 1000 steps included -- bit for bit."""
    H.build_emu("emu_aux", ["lbm_aux.cu", "lbm_phys.cuh", "lbm_common.cuh"])
    H.build_emu("emu_step_reference", ["lbm_step_kernel.cuh", "lbm_phys.cuh", "lbm_common.cuh"])
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, 2, port, out, cuts, fixture)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(timeout=600)
    for p in procs:
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
