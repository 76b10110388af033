"""z-slab domain decomposition across the GPUs of one box (replaces legacy/cuda_dual_gpu_lbm.py).

Rank r owns the contiguous planes [z0, z0+nz) of the global box plus one ghost plane on each
side.  Per step and per interface only the populations that cross it travel: cz=+1 go up,
cz=-1 go down -- 5 contiguous x-y planes each (device layout [q][z][y][x]), so no pack kernel.
The production exchange is ncclSend/ncclRecv inside liblbm_b200 (lbm_step / lbm_halo_exchange);
`exchange_halo` below performs the identical plan through torch.distributed P2P so the host-side
logic can be exercised on CPU tensors with the gloo backend (tests/test_slab_gloo.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from .config import CZ_3D

UP_Q: Tuple[int, ...] = tuple(int(q) for q in range(19) if CZ_3D[q] == 1)      # (5, 11, 12, 15, 16)
DOWN_Q: Tuple[int, ...] = tuple(int(q) for q in range(19) if CZ_3D[q] == -1)   # (6, 13, 14, 17, 18)


@dataclass(frozen=True)
class Slab:
    rank: int
    world: int
    z0: int
    nz: int
    nz_global: int

    @property
    def zghost(self) -> int:
        return 1


def partition_z(nz_global: int, world: int) -> List[Slab]:
    """Equal-thickness slabs; the remainder planes go to the lowest ranks."""
    if world < 1 or nz_global < world:
        raise ValueError("need at least one plane per rank")
    base, rem = divmod(nz_global, world)
    out, z0 = [], 0
    for r in range(world):
        nz = base + (1 if r < rem else 0)
        out.append(Slab(r, world, z0, nz, nz_global))
        z0 += nz
    return out


def partition_z_balanced(plane_weight, world: int, min_planes: int = 1) -> List[Slab]:
    """Plane-aligned slabs of (nearly) equal WORK instead of equal thickness: `plane_weight[z]` is the cost
    of plane z -- for a V60 box its fluid-cell count, which grows with z through the cone, so equal-thickness slabs leave
    the lowest rank with a third of the mean load.  Greedy cut at the prefix-sum targets k/world, every slab keeps at
    least `min_planes` planes; deterministic, so every rank computes the same partition from the same weights."""
    w = [float(x) for x in plane_weight]
    nzg = len(w)
    if world < 1 or nzg < world * min_planes:
        raise ValueError("need at least min_planes planes per rank")
    total = sum(w)
    if not total > 0:
        return partition_z(nzg, world)
    cuts, acc, z = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        lo = cuts[-1] + min_planes                       # earliest admissible cut
        hi = nzg - (world - r) * min_planes              # latest admissible cut
        while z < hi and (z < lo or acc + 0.5 * w[z] < target):
            acc += w[z]; z += 1
        cuts.append(z)
    cuts.append(nzg)
    return [Slab(r, world, cuts[r], cuts[r + 1] - cuts[r], nzg) for r in range(world)]


def neighbours(rank: int, world: int, periodic_z: bool) -> Tuple[Optional[int], Optional[int]]:
    """(rank below, rank above) or None at a non-periodic end of the chain."""
    down = rank - 1 if rank > 0 else (world - 1 if periodic_z else None)
    up = rank + 1 if rank < world - 1 else (0 if periodic_z else None)
    return down, up


def halo_bytes_per_step(nx: int, ny: int, interfaces: int = 1) -> int:
    """bytes one rank sends per step: 5 populations x nx*ny x 4 B per direction per interface."""
    return 5 * nx * ny * 4 * 2 * interfaces


def exchange_halo(g: torch.Tensor, rank: int, world: int, periodic_z: bool, group=None,
                  vec3: Optional[torch.Tensor] = None) -> None:
    """Fill the ghost planes of g [19, nz+2, ny, nx] (and of a [3, nz+2, ny, nx] field) from the z neighbours,
    sending only the outgoing populations.  Backend-agnostic mirror of liblbm_b200's NCCL exchange."""
    import torch.distributed as dist
    down, up = neighbours(rank, world, periodic_z)
    nzp = g.shape[1]
    if world == 1:
        if periodic_z:
            for q in UP_Q: g[q, 0].copy_(g[q, nzp - 2])
            for q in DOWN_Q: g[q, nzp - 1].copy_(g[q, 1])
            if vec3 is not None:
                vec3[:, 0].copy_(vec3[:, nzp - 2]); vec3[:, nzp - 1].copy_(vec3[:, 1])
        return
    ops, recv_bufs = [], []
    if up is not None:
        send = torch.stack([g[q, nzp - 2] for q in UP_Q]).contiguous()
        recv = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, up, group), dist.P2POp(dist.irecv, recv, up, group)]
        recv_bufs.append(("hi", recv))
    if down is not None:
        send = torch.stack([g[q, 1] for q in DOWN_Q]).contiguous()
        recv = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, down, group), dist.P2POp(dist.irecv, recv, down, group)]
        recv_bufs.append(("lo", recv))
    vec_bufs = []
    if vec3 is not None:
        if up is not None:
            s = vec3[:, nzp - 2].contiguous(); r = torch.empty_like(s)
            ops += [dist.P2POp(dist.isend, s, up, group), dist.P2POp(dist.irecv, r, up, group)]
            vec_bufs.append(("hi", r))
        if down is not None:
            s = vec3[:, 1].contiguous(); r = torch.empty_like(s)
            ops += [dist.P2POp(dist.isend, s, down, group), dist.P2POp(dist.irecv, r, down, group)]
            vec_bufs.append(("lo", r))
    if world == 2 and periodic_z:
        # both neighbours are the same peer: order the two message streams deterministically
        ops = ops if rank == 0 else [ops[i] for i in _swap_pairs(len(ops))]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for where, buf in recv_bufs:
        if where == "hi":       # from the rank above: its bottom plane's down-going populations
            for i, q in enumerate(DOWN_Q): g[q, nzp - 1].copy_(buf[i])
        else:                   # from the rank below: its top plane's up-going populations
            for i, q in enumerate(UP_Q): g[q, 0].copy_(buf[i])
    for where, buf in vec_bufs:
        if where == "hi": vec3[:, nzp - 1].copy_(buf)
        else: vec3[:, 0].copy_(buf)


def exchange_planes(field: torch.Tensor, rank: int, world: int, periodic_z: bool, group=None) -> None:
    """Fill the two ghost planes of a scalar [nz+2, ny, nx] or vector [C, nz+2, ny, nx] field from the z neighbours' boundary
    planes (whole planes: phi, mu, normal of the multiphase producers -- their 7-point stencils read k -+ 1).  torch.distributed
    point-to-point: NCCL on device tensors, gloo on the CPU tests.  A slab at a non-periodic end keeps its outer ghost plane."""
    import torch.distributed as dist
    zdim = field.dim() - 3
    nzp = field.shape[zdim]
    plane = lambda k: field.select(zdim, k)
    down, up = neighbours(rank, world, periodic_z)
    if world == 1:
        if periodic_z:
            plane(0).copy_(plane(nzp - 2)); plane(nzp - 1).copy_(plane(1))
        return
    ops, bufs = [], []
    for peer, send_k, recv_k in ((up, nzp - 2, nzp - 1), (down, 1, 0)):
        if peer is None:
            continue
        send = plane(send_k).contiguous(); recv = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, peer, group), dist.P2POp(dist.irecv, recv, peer, group)]
        bufs.append((recv_k, recv))
    if world == 2 and periodic_z and rank == 1:
        ops = [ops[i] for i in _swap_pairs(len(ops))]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for k, buf in bufs:
        plane(k).copy_(buf)


def reduce_ghost_up(field: torch.Tensor, rank: int, world: int, periodic_z: bool, group=None) -> None:
    """Reverse halo for scattered quantities (the particles' reaction force): what a rank deposited in its TOP ghost plane
    belongs to the first owned plane of the rank above -- send it up, add what arrives from below to plane 1, clear the ghost.
    (A particle is owned by the slab that holds its base cell, so its eight corners reach one plane up, never down.)"""
    import torch.distributed as dist
    zdim = field.dim() - 3
    nzp = field.shape[zdim]
    plane = lambda k: field.select(zdim, k)
    down, up = neighbours(rank, world, periodic_z)
    if world == 1:
        if periodic_z:
            plane(1).add_(plane(nzp - 1))
        plane(nzp - 1).zero_()
        return
    ops, recv = [], None
    if up is not None:
        ops.append(dist.P2POp(dist.isend, plane(nzp - 1).contiguous(), up, group))
    if down is not None:
        recv = torch.empty_like(plane(1).contiguous())
        ops.append(dist.P2POp(dist.irecv, recv, down, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if recv is not None:
        plane(1).add_(recv)
    plane(nzp - 1).zero_()


def particle_owner_mask(pos_z: torch.Tensor, active: torch.Tensor, z0: int, nz: int, nz_global: int) -> torch.Tensor:
    """Replicated particles, owner computes: a particle belongs to the slab that holds its base cell
    k = int(max(0, min(NZ - 2, z))) (coffee_particles.py:1054-1056, the kernel's own formula).  int32 mask: active AND owned."""
    k = pos_z.clamp(0.0, float(nz_global - 2)).to(torch.int32)
    owned = (k >= z0) & (k < z0 + nz)
    return torch.where(owned, active, torch.zeros_like(active))


def allreduce_owned(tensors, owned_active: torch.Tensor, active: torch.Tensor, group=None) -> None:
    """Make the per-particle outputs of the coupling kernel identical on every rank: the owner's value wins (sum over ranks of
    values masked to the owners), inactive particles keep what they had.  tensors: [n] or [C, n], any dtype all_reduce takes."""
    import torch.distributed as dist
    own = owned_active != 0
    act = active != 0
    for t in tensors:
        tmp = torch.where(own, t, torch.zeros_like(t))
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        t.copy_(torch.where(act, tmp, t))


def allreduce_owned_packed(tensors, owned_active: torch.Tensor, active: torch.Tensor, group=None) -> None:
    """allreduce_owned in ONE collective: the arrays ([n] or [C, n], 4-byte dtypes) are packed row-wise into one int32 buffer by
    bit pattern, masked to the owners, summed over the ranks as integers (exactly one rank contributes a non-zero word per
    particle, so the sum is exact for any dtype) and unpacked.  Seven small all-reduces per coupling call become one."""
    import torch.distributed as dist
    own = owned_active != 0
    act = active != 0
    rows = [t.reshape(-1, t.shape[-1]) for t in tensors]
    if any(r.element_size() != 4 for r in rows):
        raise TypeError("allreduce_owned_packed packs 4-byte element types")
    buf = torch.cat([r.view(torch.int32) for r in rows], dim=0)
    buf = torch.where(own, buf, torch.zeros_like(buf))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    k = 0
    for t, r in zip(tensors, rows):
        got = buf[k:k + r.shape[0]].view(t.dtype).reshape(t.shape)
        t.copy_(torch.where(act, got, t))
        k += r.shape[0]


def _swap_pairs(n: int) -> List[int]:
    """rank 1 of a 2-rank periodic ring posts its (send,recv) pairs in the opposite neighbour order so that
    message k of rank 0 meets message k of rank 1."""
    idx = list(range(n))
    pairs = [idx[i:i + 2] for i in range(0, n, 2)]        # (send,recv) pairs: up, down, [vec up, vec down]
    out = []
    for i in range(0, len(pairs) - 1, 2):
        out += pairs[i + 1] + pairs[i]
    if len(pairs) % 2:
        out += pairs[-1]
    return out
