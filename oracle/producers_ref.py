"""CPU ORACLE (test infrastructure, NOT product code) -- NumPy restatement of the per-step producers of
`body_force`, `phase` and `rho` that sit next to the D3Q19 step in the reference:

  * MultiphaseFlow3D (src/core/multiphase_3d.py): the surface-tension chain main.py runs before the collision
    (`accumulate_surface_tension_pre_collision` :409-418 = `compute_gradients` :111-132, `compute_curvature` :134-149,
    `compute_surface_tension_force` :313-332 -- the second definition, the one Python binds --, `apply_surface_tension`
    :354-363) and the phase-field step after it (`step` :389-407 = the same three kernels, then
    `update_phase_field_cahn_hilliard` :151-197, `apply_phase_separation` :334-352, `copy_phase_field` :383-387,
    `update_density_from_phase` :365-381);
  * PrecisePouringSystem (src/physics/precise_pouring.py): `_get_current_pour_position` :81-97,
    `_is_in_pouring_region` :99-129, `apply_pouring_force` :131-163, `apply_gradual_phase_change` :165-196.

PARITY STATUS: pinned against the reference's own source code -- tests/golden/make_reference_goldens.py runs the
unmodified modules above under the pure-Python Taichi stand-in and records their outputs
(tests/golden/reference_run_multiphase.npz); tests/test_oracle_vs_reference_run.py requires this module to reproduce the
recorded arrays bit for bit (the Gaussian of the nozzle profile goes through exp: NumPy on both sides here, `expf` on
the device, so the GPU test allows the 2 ulp of CUDA's expf there and nowhere else).

Only tests/ may import this module.  Arrays are in the reference's logical index order ([i, j, k], vectors [i, j, k, c]);
all arithmetic is IEEE f32 in source order, one rounding per operation.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
_C = (slice(1, -1), slice(1, -1), slice(1, -1))        # ti.ndrange((1, N-1), (1, N-1), (1, N-1))


def _sh(a, ax, d):
    """a[i+d] along axis ax, restricted to the interior block."""
    sl = [slice(1, -1)] * 3
    sl[ax] = slice(1 + d, a.shape[ax] - 1 + d)
    return a[tuple(sl)]


def _norm(vx, vy, vz):
    return np.sqrt((vx * vx + vy * vy) + vz * vz)       # Vector.norm(): sqrt of the left-to-right dot product


class MultiphaseState:
    """The fields MultiphaseFlow3D owns (multiphase_3d.py:24-38); all start at zero like ti.field."""

    def __init__(self, n):
        sh = (n, n, n) if isinstance(n, int) else tuple(n)
        self.phi = np.zeros(sh, F32); self.phi_new = np.zeros(sh, F32); self.mu = np.zeros(sh, F32)
        self.grad_phi = np.zeros(sh + (3,), F32); self.grad_mu = np.zeros(sh + (3,), F32)
        self.normal = np.zeros(sh + (3,), F32); self.curvature = np.zeros(sh, F32)
        self.surface_force = np.zeros(sh + (3,), F32)


def compute_chemical_potential(m: MultiphaseState, sigma: float, interface_width: float = 2.0):
    """multiphase_3d.py:80-109; returns laplacian_phi (the reference keeps it as a field)."""
    phi = m.phi
    p0 = phi[_C]
    lap = np.zeros_like(phi)
    lap[_C] = (((((_sh(phi, 0, +1) + _sh(phi, 0, -1)) + _sh(phi, 1, +1)) + _sh(phi, 1, -1)) + _sh(phi, 2, +1)) + _sh(phi, 2, -1)) \
        - F32(6.0) * p0
    kappa = F32(3.0 * sigma * interface_width / 8.0)
    m.mu[_C] = ((p0 * p0) * p0 - p0) + (-kappa) * lap[_C]
    return lap


def compute_gradients(m: MultiphaseState) -> None:
    """multiphase_3d.py:111-132 (interior cells; the outer layer keeps what it held)."""
    for field, grad in ((m.phi, m.grad_phi), (m.mu, m.grad_mu)):
        for ax in range(3):
            grad[_C + (ax,)] = (_sh(field, ax, +1) - _sh(field, ax, -1)) * F32(0.5)
    gx, gy, gz = (m.grad_phi[_C + (c,)] for c in range(3))
    mag = _norm(gx, gy, gz)
    ok = mag > F32(1e-10)
    safe = np.where(ok, mag, F32(1.0))
    for c, g in enumerate((gx, gy, gz)):
        m.normal[_C + (c,)] = np.where(ok, g / safe, F32(0.0))


def compute_curvature(m: MultiphaseState) -> None:
    """multiphase_3d.py:134-149."""
    nx_, ny_, nz_ = (m.normal[..., c] for c in range(3))
    has_normal = _norm(nx_[_C], ny_[_C], nz_[_C]) > F32(1e-10)
    dnx = (_sh(nx_, 0, +1) - _sh(nx_, 0, -1)) * F32(0.5)
    dny = (_sh(ny_, 1, +1) - _sh(ny_, 1, -1)) * F32(0.5)
    dnz = (_sh(nz_, 2, +1) - _sh(nz_, 2, -1)) * F32(0.5)
    m.curvature[_C] = np.where(has_normal, (dnx + dny) + dnz, F32(0.0))


def compute_surface_tension_force(m: MultiphaseState, sigma: float) -> None:
    """multiphase_3d.py:313-332 (the definition that is live: it shadows :199-214)."""
    gx, gy, gz = (m.grad_phi[_C + (c,)] for c in range(3))
    grad_mag = _norm(gx, gy, gz)
    on = (np.abs(m.phi[_C]) < F32(0.9)) & (grad_mag > F32(1e-10))
    force_magnitude = (F32(sigma) * m.curvature[_C]) * grad_mag
    for c in range(3):
        m.surface_force[_C + (c,)] = np.where(on, force_magnitude * m.normal[_C + (c,)], F32(0.0))


def apply_surface_tension(m: MultiphaseState, rho, solid, body_force) -> None:
    """multiphase_3d.py:354-363: body_force += surface_force / rho on fluid cells with rho > 1e-10 (all cells)."""
    sel = (solid == 0) & (rho > F32(1e-10))
    safe = np.where(sel, rho, F32(1.0))
    for c in range(3):
        body_force[..., c] = np.where(sel, body_force[..., c] + m.surface_force[..., c] / safe, body_force[..., c])


def accumulate_surface_tension_pre_collision(m: MultiphaseState, rho, solid, body_force, sigma: float) -> None:
    """multiphase_3d.py:409-418."""
    compute_gradients(m); compute_curvature(m); compute_surface_tension_force(m, sigma)
    apply_surface_tension(m, rho, solid, body_force)


def update_phase_field_cahn_hilliard(m: MultiphaseState, u, mobility: float, dt: float) -> None:
    """multiphase_3d.py:151-197: first-order upwind advection + M lap(mu), explicit Euler, clamp to [-1, 1]."""
    phi = m.phi
    p0 = phi[_C]
    d = []
    for ax in range(3):
        ua = u[_C + (ax,)]
        back = p0 - _sh(phi, ax, -1)
        fwd = _sh(phi, ax, +1) - p0
        d.append(np.where(ua > 0, back, fwd))
    ux, uy, uz = (u[_C + (c,)] for c in range(3))
    convection = -((ux * d[0] + uy * d[1]) + uz * d[2])
    mu = m.mu
    lap_mu = (((((_sh(mu, 0, +1) + _sh(mu, 0, -1)) + _sh(mu, 1, +1)) + _sh(mu, 1, -1)) + _sh(mu, 2, +1)) + _sh(mu, 2, -1)) \
        - F32(6.0) * mu[_C]
    diffusion = F32(mobility) * lap_mu
    new = p0 + F32(dt) * (convection + diffusion)
    m.phi_new[_C] = np.maximum(F32(-1.0), np.minimum(F32(1.0), new))


def apply_phase_separation(m: MultiphaseState, dt: float) -> None:
    """multiphase_3d.py:334-352."""
    phi = m.phi
    p0 = phi[_C]
    lap = (((((_sh(phi, 0, +1) + _sh(phi, 0, -1)) + _sh(phi, 1, +1)) + _sh(phi, 1, -1)) + _sh(phi, 2, +1)) + _sh(phi, 2, -1)) \
        - F32(6.0) * p0
    chem = p0 * (p0 * p0 - F32(1.0)) - F32(0.01) * lap
    inc = (F32(-0.001) * chem) * F32(dt)
    m.phi_new[_C] = np.where(np.abs(p0) < F32(0.99), m.phi_new[_C] + inc, m.phi_new[_C])


def update_density_from_phase(m: MultiphaseState, rho, phase, rho_water: float, rho_air: float) -> None:
    """multiphase_3d.py:365-381 (all cells)."""
    p = np.maximum(F32(-1.0), np.minimum(F32(1.0), m.phi))
    rho[...] = F32(rho_air) + (F32(rho_water - rho_air) * (p + F32(1.0))) / F32(2.0)
    phase[...] = (p + F32(1.0)) / F32(2.0)


def multiphase_step(m: MultiphaseState, u, rho, phase, solid, body_force, sigma, mobility, dt, rho_water, rho_air,
                    step_count=0, precollision_applied=False) -> None:
    """MultiphaseFlow3D.step, multiphase_3d.py:389-407."""
    compute_gradients(m); compute_curvature(m); compute_surface_tension_force(m, sigma)
    if (not precollision_applied) and step_count > 10:
        apply_surface_tension(m, rho, solid, body_force)
    update_phase_field_cahn_hilliard(m, u, mobility, dt)
    apply_phase_separation(m, dt)
    m.phi[...] = m.phi_new                                      # copy_phase_field :383-387
    update_density_from_phase(m, rho, phase, rho_water, rho_air)


# ---- PrecisePouringSystem -------------------------------------------------------------------------------------------
class PourState:
    """The scalars of PrecisePouringSystem (precise_pouring.py:14-47); f32 fields like the reference's 0-d ti.fields."""

    def __init__(self, n, diameter_grid: float, height: int, velocity: float):
        self.n = n
        self.POUR_DIAMETER_GRID = float(diameter_grid); self.POUR_HEIGHT = int(height); self.POUR_VELOCITY = float(velocity)
        self.active = 0; self.pattern = 0
        self.center_x = F32(0); self.center_y = F32(0); self.flow_rate = F32(0); self.pour_time = F32(0)
        self.spiral_radius = F32(0); self.spiral_speed = F32(0); self.spiral_cx = F32(0); self.spiral_cy = F32(0)

    def start_pouring(self, center_x=None, center_y=None, flow_rate=1.0, pattern="center"):
        """precise_pouring.py:49-74."""
        cx = self.n // 2 if center_x is None else center_x
        cy = self.n // 2 if center_y is None else center_y
        self.center_x, self.center_y, self.flow_rate = F32(cx), F32(cy), F32(flow_rate)
        self.active = 1; self.pour_time = F32(0.0)
        if pattern == "center":
            self.pattern = 0
        elif pattern == "spiral":
            self.pattern = 1
            self.spiral_cx, self.spiral_cy = F32(cx), F32(cy)
            self.spiral_radius, self.spiral_speed = F32(5.0), F32(1.0)

    def position(self):
        """_get_current_pour_position, precise_pouring.py:81-97."""
        x, y = self.center_x, self.center_y
        if self.pattern == 1:
            t = self.pour_time * self.spiral_speed
            r = self.spiral_radius * (F32(1.0) + F32(0.1) * t)
            x = self.spiral_cx + r * np.cos(t)
            y = self.spiral_cy + r * np.sin(t)
            d = self.POUR_DIAMETER_GRID
            x = max(F32(d), min(F32(self.n - d), x))
            y = max(F32(d), min(F32(self.n - d), y))
        return F32(x), F32(y)


def pouring_intensity(p: PourState, shape, pour_x, pour_y):
    """_is_in_pouring_region, precise_pouring.py:99-129, evaluated on the whole grid.  Returns total_intensity [i,j,k]."""
    nx, ny, nz = shape
    i = np.arange(nx, dtype=np.int32)[:, None, None]
    j = np.arange(ny, dtype=np.int32)[None, :, None]
    k = np.arange(nz, dtype=np.int32)[None, None, :]
    dx = i.astype(F32) - F32(pour_x)
    dy = j.astype(F32) - F32(pour_y)
    dist = np.sqrt(dx * dx + dy * dy)
    radius = F32(p.POUR_DIAMETER_GRID / 2.0)
    pour_z = p.POUR_HEIGHT
    inside = (dist <= radius) & (k <= pour_z) & (k.astype(F32) >= F32(pour_z) - F32(4.0))
    t = dist / radius
    intensity = np.exp(F32(-0.5) * (t * t)).astype(F32)
    # the vertical distance is an integer: the constant expression exp(-d / 2.0) folds in f64 and is stored as f32
    decay = np.array([F32(math.exp(-(pour_z - kk) / 2.0)) for kk in range(nz)], F32)[None, None, :]
    return np.where(inside, intensity * decay, F32(0.0)).astype(F32) * np.ones(shape, F32)


def apply_pouring_force(p: PourState, body_force, solid, dt: float) -> None:
    """precise_pouring.py:131-163: body_force.z -= min(POUR_VELOCITY * intensity * flow_rate / dt, 10) under the nozzle."""
    if p.active != 1:
        return
    dt = F32(dt)
    p.pour_time = F32(p.pour_time + dt)
    px, py = p.position()
    tot = pouring_intensity(p, solid.shape, px, py)
    sel = (solid == 0) & (tot > 0)
    if dt > F32(1e-8):
        accel = ((F32(p.POUR_VELOCITY) * tot) * p.flow_rate) / dt
    else:
        accel = np.zeros_like(tot)
    accel = np.minimum(accel, F32(10.0))
    body_force[..., 0] = np.where(sel, body_force[..., 0] + F32(0.0), body_force[..., 0])
    body_force[..., 1] = np.where(sel, body_force[..., 1] + F32(0.0), body_force[..., 1])
    body_force[..., 2] = np.where(sel, body_force[..., 2] + (-accel), body_force[..., 2])


def apply_gradual_phase_change(p: PourState, phi, solid, dt: float) -> None:
    """precise_pouring.py:165-196: relax phi toward +1 under the nozzle (rate limited, intensity weighted)."""
    if p.active != 1:
        return
    dt = F32(dt)
    px, py = p.position()
    tot = pouring_intensity(p, solid.shape, px, py)
    sel = (solid == 0) & (tot > 0)
    rate = (F32(1.0) - phi) / F32(0.05)
    rate = np.maximum(F32(-2.0), np.minimum(F32(2.0), rate))
    change = ((rate * tot) * dt) * p.flow_rate
    new = np.maximum(F32(-1.0), np.minimum(F32(1.0), phi + change))
    phi[...] = np.where(sel, new, phi)


# ---- FilterPaperSystem: particle interception and dynamic resistance (src/physics/filter_paper.py) -------------------
def update_dynamic_resistance(filter_zone, blockage, accumulated) -> None:
    """filter_paper.py:703-746 (filter-zone cells, in place)."""
    z = filter_zone == 1
    nb = F32(0.9) * (F32(1.0) - np.exp(F32(-0.1) * accumulated).astype(F32))
    blockage[...] = np.where(z, F32(0.95) * blockage + F32(0.05) * nb, blockage)
    accumulated[...] = np.where(z, accumulated * F32(0.999), accumulated)


def uniform01(seed: int, p: int, d: int) -> np.float32:
    """The device's counter-based draw (csrc/lbm_producers.cu:uniform01, lowbias32 hash): the reference's ti.random() is an
    unseeded per-thread stream, so the kick is OUR definition -- a pure function of (seed, particle, draw)."""
    m = 0xFFFFFFFF
    h = (seed ^ ((p * 0x9E3779B9) & m) ^ ((d * 0x85EBCA6B) & m)) & m
    h ^= h >> 16; h = (h * 0x7FEB352D) & m; h ^= h >> 15; h = (h * 0x846CA68B) & m; h ^= h >> 16
    return F32(h >> 8) * F32(1.0 / 16777216.0)


def block_particles_at_filter(filter_zone, pos, vel, active, accumulated, scale_length: float, noise: float = 0.01, seed: int = 0) -> None:
    """filter_paper.py:616-700; pos / vel [P,3] in place.  noise = 0 is the reference with ti.random() == 0.5."""
    nx, ny, nz = filter_zone.shape
    sl = F32(scale_length)
    for p in range(pos.shape[0]):
        if active[p] == 0:
            continue
        g = [int(np.trunc(np.clip(pos[p, c] / sl, -2.0e9, 2.0e9))) for c in range(3)]
        if not (0 <= g[0] < nx and 0 <= g[1] < ny and 0 <= g[2] < nz):
            continue
        for off in range(-2, 3):
            k = g[2] + off
            if 0 <= k < nz and filter_zone[g[0], g[1], k] == 1:
                if vel[p, 2] < 0:
                    vel[p, 2] = (-vel[p, 2]) * F32(0.3)
                    vel[p, 0] = vel[p, 0] + (uniform01(seed, p, 0) - F32(0.5)) * F32(noise)
                    vel[p, 1] = vel[p, 1] + (uniform01(seed, p, 1) - F32(0.5)) * F32(noise)
                    accumulated[g[0], g[1], k] = accumulated[g[0], g[1], k] + F32(0.01)
                break


# ---- CoffeeParticleSystem.apply_fluid_forces (src/physics/coffee_particles.py:547-639) ------------------------------
def _valid_coordinate(x, y, z, max_coord):
    ok = not (x < 0 or x > max_coord or y < 0 or y > max_coord or z < 0 or z > max_coord)
    if not (x == x and y == y and z == z):
        ok = False
    if abs(x) > 1e6 or abs(y) > 1e6 or abs(z) > 1e6:
        ok = False
    return ok


def apply_fluid_forces(u, pos, vel, radius, mass, active, force, water_density, water_viscosity, gravity=9.81) -> int:
    """coffee_particles.py:547-639, in place on vel / active / force ([P,3]); returns the coordinate-error count.
    u is the LBM vector field [NX,NY,NZ,3]; water_viscosity = WATER_VISCOSITY_90C * WATER_DENSITY_90C (:70)."""
    nx, ny, nz = u.shape[:3]
    max_coord = F32(max(nx, ny, nz))
    rho_w = F32(water_density); mu_safe = F32(max(1e-8, water_viscosity)); g = F32(gravity)
    vol_k = F32((4.0 / 3.0) * 3.14159)
    errors = 0
    fmax, fmin = (lambda a, b: a if a >= b else b), (lambda a, b: a if a <= b else b)
    for p in range(pos.shape[0]):
        if active[p] != 1:
            continue
        x, y, z = pos[p]
        if not _valid_coordinate(x, y, z, max_coord):
            active[p] = 0; errors += 1
            continue
        gi = int(fmax(F32(0), fmin(F32(nx - 2), x))); gj = int(fmax(F32(0), fmin(F32(ny - 2), y))); gk = int(fmax(F32(0), fmin(F32(nz - 2), z)))
        f = u[gi, gj, gk]
        fs = np.sqrt((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2])
        if not (fs == fs and fs <= F32(100.0)):
            continue
        v = vel[p].copy()
        s2 = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]
        if not (v[0] == v[0] and v[1] == v[1] and v[2] == v[2]) or s2 > F32(100.0):
            vel[p] = 0; v = vel[p].copy()
        r = f - v
        rs = np.sqrt((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2])
        if not (rs > F32(1e-6) and rs < F32(10.0)):
            continue
        rad, m = radius[p], mass[p]
        if (rad < F32(1e-5) or rad > F32(0.01) or rad != rad) or not (m == m and m > 0):
            continue
        re = (((rs * F32(2.0)) * rad) * rho_w) / mu_safe
        re = fmax(F32(0.01), fmin(F32(1000.0), re))
        cd = F32(24.0) / fmax(F32(0.1), re)
        cd = fmax(F32(0.1), fmin(F32(10.0), cd))
        dm = ((((F32(0.5) * cd) * F32(3.14159)) * (rad * rad)) * rho_w) * rs
        dm = fmin(dm, m * F32(100.0))
        drag = dm * (r / rs) if rs > 0 else np.zeros(3, F32)
        volume = vol_k * ((rad * rad) * rad)
        bm = fmin((volume * rho_w) * g, m * F32(20.0))
        gm = m * g
        buoy = np.array([bm * F32(0), bm * F32(0), bm * F32(1)], F32)
        grav = np.array([gm * F32(0), gm * F32(0), gm * F32(-1)], F32)
        total = (drag + buoy) + grav
        fm = np.sqrt((total[0] * total[0] + total[1] * total[1]) + total[2] * total[2])
        force[p] = total if (fm == fm and fm < m * F32(1000.0)) else grav
    return errors
