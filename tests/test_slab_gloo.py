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
