// Coffee-particle two-way coupling (src/physics/coffee_particles.py:1048-1212):
// one thread per particle -- trilinear gather of the fluid velocity, Schiller-Naumann drag,
// warp-aggregated atomic scatter of the reaction force, under-relaxation.
// Compiled with -fmad=false: gather/drag values follow oracle/d3q19_ref.py bit for bit except
// powf (<= 2 ulp) and the order of the atomic adds.
#include "lbm_common.cuh"

namespace lbm {

struct ParticleArgs {
    Grid g;
    const float *u;        // [3][vol]
    float *reaction;       // [3][vol]
    lbm_particles ps;
    float rho_w, mu_w, relax;
};

// i = int(max(0, min(N-2, x)))  (f32 clamp, truncation) -- coffee_particles.py:1054-1056
__device__ __forceinline__ int base_cell(float x, int n, float &frac) {
    const float cl = fmaxf(0.0f, fminf((float)(n - 2), x));
    const int i = (int)cl;
    float fr = x - (float)i;
    frac = fmaxf(0.0f, fminf(1.0f, fr));
    return i;
}

// coffee_particles.py:1088-1099
__device__ __forceinline__ float drag_coefficient(float re) {
    float cd = 24.0f / fmaxf(0.01f, re);
    if (re >= 0.1f && re < 1000.0f) cd = (24.0f / re) * (1.0f + 0.15f * powf(re, 0.687f));
    else if (re >= 1000.0f) cd = 0.44f;
    return cd;
}

__global__ void __launch_bounds__(256) particles_couple_kernel(const ParticleArgs A) {
    const Grid &G = A.g;
    const lbm_particles &P = A.ps;
    const int n = P.n;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < n;
    bool act = valid && P.active[p] != 0;
    if (act && G.zg) {       // z-slabs, replicated particles: the rank whose slab holds the base cell computes (slab.particle_owner_mask)
        float fz_;
        const int kg_ = base_cell(P.pos[2 * n + p], G.nz_global, fz_);
        act = kg_ >= G.z0 && kg_ < G.z0 + G.nz;
    }

    float rx = 0.0f, ry = 0.0f, rz = 0.0f;     // reaction force of this particle
    float w[8];
    long long base = -1;                       // linear index of the base cell (scatter key)
    bool scatter = false;
    if (act) {
        const float px = P.pos[p], py = P.pos[n + p], pz = P.pos[2 * n + p];
        float fx, fy, fz;
        const int i = base_cell(px, G.nx, fx), j = base_cell(py, G.ny, fy);
        // z is sliced into slabs: global extent for the clamp, local plane for the address
        const int kg = base_cell(pz, G.nz_global, fz);
        P.cell[p] = i; P.cell[n + p] = j; P.cell[2 * n + p] = kg;
        const int kl = kg - G.z0 + G.zg;
        base = ((long long)kl * G.ny + j) * G.nx + i;
        const float gx = 1.0f - fx, gy = 1.0f - fy, gz = 1.0f - fz;
        // corner order (dx,dy,dz): 000 001 010 011 100 101 110 111  (coffee_particles.py:1176-1183)
        w[0] = (gx * gy) * gz; w[1] = (gx * gy) * fz; w[2] = (gx * fy) * gz; w[3] = (gx * fy) * fz;
        w[4] = (fx * gy) * gz; w[5] = (fx * gy) * fz; w[6] = (fx * fy) * gz; w[7] = (fx * fy) * fz;
        float uf[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float *u = A.u + (long long)d * G.vol + base;
            float acc = w[0] * __ldg(u);
            acc = acc + w[1] * __ldg(u + G.plane);
            acc = acc + w[2] * __ldg(u + G.nx);
            acc = acc + w[3] * __ldg(u + G.nx + G.plane);
            acc = acc + w[4] * __ldg(u + 1);
            acc = acc + w[5] * __ldg(u + 1 + G.plane);
            acc = acc + w[6] * __ldg(u + 1 + G.nx);
            acc = acc + w[7] * __ldg(u + 1 + G.nx + G.plane);
            uf[d] = acc;
            P.u_fluid[d * n + p] = acc;
        }
        const float relx = uf[0] - P.vel[p], rely = uf[1] - P.vel[n + p], relz = uf[2] - P.vel[2 * n + p];
        const float mag = sqrtf(dot3(relx, rely, relz, relx, rely, relz));
        float dnx = 0.0f, dny = 0.0f, dnz = 0.0f, re = 0.0f, cd = 0.0f;
        if (mag > 1e-8f) {
            const float radius = P.radius[p];
            re = (((A.rho_w * mag) * 2.0f) * radius) / fmaxf(1e-8f, A.mu_w);
            cd = drag_coefficient(re);
            const float area = (3.14159f * radius) * radius;
            float dmag = (((0.5f * A.rho_w) * cd) * area) * mag;
            dmag = fminf(dmag, P.mass[p] * 100.0f);
            dnx = (dmag * relx) / mag; dny = (dmag * rely) / mag; dnz = (dmag * relz) / mag;
            rx = -dnx; ry = -dny; rz = -dnz;
            scatter = true;
        }
        P.reynolds[p] = re; P.cd[p] = cd;
        P.drag_new[p] = dnx; P.drag_new[n + p] = dny; P.drag_new[2 * n + p] = dnz;
        // under-relaxation (coffee_particles.py:1200-1212), fused when relax >= 0
        if (A.relax >= 0.0f) {
            const float a = A.relax, b = 1.0f - A.relax;
            const float ox = P.drag_old[p], oy = P.drag_old[n + p], oz = P.drag_old[2 * n + p];
            const float fxr = a * dnx + b * ox, fyr = a * dny + b * oy, fzr = a * dnz + b * oz;
            P.drag[p] = fxr; P.drag[n + p] = fyr; P.drag[2 * n + p] = fzr;
            P.drag_old[p] = fxr; P.drag_old[n + p] = fyr; P.drag_old[2 * n + p] = fzr;
        }
    }   // inactive particles keep their previous drag (the reference only touches active ones)

    // Warp-aggregated scatter: lanes whose particles share a base cell are combined with
    // shuffles and only the group leader issues the 24 reductions (red.global.add.f32).
    const long long key = scatter ? base : -1 - (long long)(threadIdx.x & 31);   // unique key for idle lanes
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    float c[24];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float wk = scatter ? w[k] : 0.0f;
        c[3 * k] = wk * rx; c[3 * k + 1] = wk * ry; c[3 * k + 2] = wk * rz;
    }
    if (peers != (1u << lane)) {        // group of 2+ lanes: sum the members' contributions into the leader
        unsigned rest = peers & ~(1u << leader);
        while (rest) {
            const int src = __ffs(rest) - 1;
            rest &= rest - 1;
#pragma unroll
            for (int k = 0; k < 24; ++k) {
                const float v = __shfl_sync(peers, c[k], src);
                if (lane == leader) c[k] += v;
            }
        }
    }
    if (scatter && lane == leader) {
        // corner offsets in the same (dx,dy,dz) order as w[]; the reference's atomics (:1079-1086) are unordered
        const long long off[8] = {0, G.plane, G.nx, G.nx + G.plane, 1, 1 + G.plane, 1 + G.nx, 1 + G.nx + G.plane};
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int d = 0; d < 3; ++d) atomicAdd(A.reaction + (long long)d * G.vol + base + off[k], c[3 * k + d]);
    }
}

// Sparse clear of the reaction field: zero the 8 corners (x 3 components) of every particle's RECORDED base cell, i.e. exactly
// what the previous coupling call deposited (it wrote P.cell).  1 M particles: 24 M four-byte stores instead of a memset of the
// whole field (1.6 GB at 512^3: 0.27 of the coupling's 0.37 ms; 7 GB on a 1024^2 x 600 slab).  Corners outside the slab's planes
// (ghost planes included) are skipped.
__global__ void __launch_bounds__(256) particles_clear_deposits_kernel(Grid G, float *reaction, const int *cell, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = cell[p], j = cell[n + p], kl = cell[2 * n + p] - G.z0 + G.zg;
    if (i < 0 || i > G.nx - 2 || j < 0 || j > G.ny - 2) return;
    const int nzp = G.nz + 2 * G.zg;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz) {
        const int kk = kl + dz;
        if (kk < 0 || kk >= nzp) continue;
        const long long b = ((long long)kk * G.ny + j) * G.nx + i;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float *r = reaction + (long long)d * G.vol + b;
            r[0] = 0.0f; r[1] = 0.0f; r[G.nx] = 0.0f; r[G.nx + 1] = 0.0f;
        }
    }
}

// CoffeeParticleSystem.apply_under_relaxation as a stand-alone call (coffee_particles.py:1200-1212)
__global__ void particles_under_relax_kernel(lbm_particles P, float relax) {
    const int n = P.n;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || P.active[p] == 0) return;
    const float a = relax, b = 1.0f - relax;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float v = a * P.drag_new[d * n + p] + b * P.drag_old[d * n + p];
        P.drag[d * n + p] = v; P.drag_old[d * n + p] = v;
    }
}

#ifndef LBM_EMULATE_ON_HOST      /* tests/emu compiles the kernels with g++ and runs them thread by thread */
cudaError_t launch_particles_clear_deposits(const Grid &G, float *reaction, const lbm_particles &ps, cudaStream_t s) {
    const int b = 256, gr = (ps.n + b - 1) / b;
    if (ps.n > 0) particles_clear_deposits_kernel<<<gr, b, 0, s>>>(G, reaction, ps.cell, ps.n);
    return cudaGetLastError();
}
cudaError_t launch_particles_under_relax(const lbm_particles &ps, float relax, cudaStream_t s) {
    const int b = 256, gr = (ps.n + b - 1) / b;
    if (ps.n > 0) particles_under_relax_kernel<<<gr, b, 0, s>>>(ps, relax);
    return cudaGetLastError();
}
#endif

// ---------------------------------------------------------------------------------------------
// CoffeeParticleSystem.update_particle_physics (coffee_particles.py:641-720) with its helpers validate_coordinate
// (:75-92), validate_velocity (:95-108), check_particle_boundary_violation_safe (:734-778) and
// constrain_to_boundary_safe (:780-831): explicit Euler with a clamped time step, acceleration and displacement caps,
// the cone constraint (0.9 / 0.8 radius factors), velocity damping 0.3.  One thread per particle; every statement in
// the reference's order, f32, no contraction (-fmad=false), so positions / velocities / active flags match
// oracle/d3q19_ref.py:update_particle_physics bit for bit.  `force` ([3][n], may be NULL = no force) is zeroed like
// the reference's self.force; counters[0] += coordinate errors, counters[1] += boundary violations.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool valid_coordinate(float x, float y, float z, float max_coord) {
    bool ok = !(x < 0.0f || x > max_coord || y < 0.0f || y > max_coord || z < 0.0f || z > max_coord);
    if (!(x == x && y == y && z == z)) ok = false;
    if (fabsf(x) > 1e6f || fabsf(y) > 1e6f || fabsf(z) > 1e6f) ok = false;
    return ok;
}
__device__ __forceinline__ bool valid_velocity(float vx, float vy, float vz) {
    const float s2 = (vx * vx + vy * vy) + vz * vz;
    bool ok = vx == vx && vy == vy && vz == vz;
    if (s2 > 10.0f * 10.0f) ok = false;
    return ok;
}
__device__ __forceinline__ float norm3(float x, float y, float z) { return sqrtf((x * x + y * y) + z * z); }

__global__ void __launch_bounds__(256) particles_advance_kernel(lbm_particles P, float *force, lbm_particle_bounds B, float dt, int *counters) {
    const int n = P.n;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || P.active[p] != 1) return;
    const float dt_safe = fmaxf(1e-8f, fminf(1e-2f, dt));
    int coord_err = 0, viol = 0;
    float px = P.pos[p], py = P.pos[n + p], pz = P.pos[2 * n + p];
    if (!valid_coordinate(px, py, pz, B.max_coordinate)) {
        P.active[p] = 0;
        atomicAdd(counters, 1);
        return;                                                    // `continue`: the force is NOT cleared on this path
    }
    float vx = P.vel[p], vy = P.vel[n + p], vz = P.vel[2 * n + p];
    const float mass = P.mass[p];
    if (mass > 1e-10f) {
        float ax = 0.0f, ay = 0.0f, az = 0.0f;
        if (force) { ax = force[p] / mass; ay = force[n + p] / mass; az = force[2 * n + p] / mass; }
        const float amag = norm3(ax, ay, az);
        if (amag > 1000.0f) { const float s = 1000.0f / amag; ax = ax * s; ay = ay * s; az = az * s; }
        const float nvx = vx + ax * dt_safe, nvy = vy + ay * dt_safe, nvz = vz + az * dt_safe;
        if (valid_velocity(nvx, nvy, nvz)) { vx = nvx; vy = nvy; vz = nvz; }
        else { vx = vy = vz = 0.0f; ++coord_err; }
    }
    float dx = vx * dt_safe, dy = vy * dt_safe, dz = vz * dt_safe;
    const float dmag = norm3(dx, dy, dz);
    if (dmag > 1.0f) { const float s = 1.0f / dmag; dx = dx * s; dy = dy * s; dz = dz * s; }
    float nx_ = px + dx, ny_ = py + dy, nz_ = pz + dz;

    // check_particle_boundary_violation_safe
    bool violation = false;
    if (!valid_coordinate(nx_, ny_, nz_, B.max_coordinate)) violation = true;
    else if (nz_ < B.bottom_z - 1.0f) violation = true;
    else {
        const float ddx = nx_ - B.center_x, ddy = ny_ - B.center_y;
        const float d2 = ddx * ddx + ddy * ddy;
        if (d2 > 1e6f) violation = true;
        else {
            const float dist = sqrtf(d2);
            const float hd = nz_ - B.bottom_z;
            if (hd >= 0.0f && hd < B.cup_height_lu) {
                float hr = hd / B.cup_height_lu;
                hr = fmaxf(0.0f, fminf(1.0f, hr));
                const float max_r = B.bottom_radius_lu + (B.top_radius_lu - B.bottom_radius_lu) * hr;
                if (dist > max_r * 0.9f) violation = true;
            } else if (hd >= B.cup_height_lu) {
                if (dist > B.top_radius_lu * 0.9f) violation = true;
            }
        }
    }
    if (violation) {
        // constrain_to_boundary_safe
        float cx_ = nx_, cy_ = ny_, cz_ = nz_;
        if (cz_ < B.bottom_z) cz_ = B.bottom_z + 0.1f;
        const float max_z = fminf(B.bottom_z + B.cup_height_lu * 1.5f, B.nz_minus_5);
        if (cz_ > max_z) cz_ = max_z - 0.1f;
        const float ddx = cx_ - B.center_x, ddy = cy_ - B.center_y;
        const float d2 = ddx * ddx + ddy * ddy;
        if (d2 < 1e6f) {
            const float dist = sqrtf(d2);
            if (dist > 0.1f) {
                float hd = cz_ - B.bottom_z;
                hd = fmaxf(0.0f, hd);
                float max_r = B.top_radius_lu;
                if (hd < B.cup_height_lu) {
                    float hr = hd / B.cup_height_lu;
                    hr = fmaxf(0.0f, fminf(1.0f, hr));
                    max_r = B.bottom_radius_lu + (B.top_radius_lu - B.bottom_radius_lu) * hr;
                }
                if (dist > max_r * 0.8f) {
                    float sf = (max_r * 0.8f) / dist;
                    sf = fmaxf(0.1f, fminf(1.0f, sf));
                    cx_ = B.center_x + ddx * sf;
                    cy_ = B.center_y + ddy * sf;
                }
            }
        } else { cx_ = B.center_x; cy_ = B.center_y; }
        if (valid_coordinate(cx_, cy_, cz_, B.max_coordinate)) {
            nx_ = cx_; ny_ = cy_; nz_ = cz_;
            vx = vx * 0.3f; vy = vy * 0.3f; vz = vz * 0.3f;
            ++viol;
        } else {
            nx_ = px; ny_ = py; nz_ = pz;
            vx = vy = vz = 0.0f;
            ++coord_err;
        }
    }
    if (valid_coordinate(nx_, ny_, nz_, B.max_coordinate)) { P.pos[p] = nx_; P.pos[n + p] = ny_; P.pos[2 * n + p] = nz_; }
    else { P.active[p] = 0; ++coord_err; }
    P.vel[p] = vx; P.vel[n + p] = vy; P.vel[2 * n + p] = vz;
    if (force) { force[p] = 0.0f; force[n + p] = 0.0f; force[2 * n + p] = 0.0f; }
    if (coord_err) atomicAdd(counters, coord_err);
    if (viol) atomicAdd(counters + 1, viol);
}

// ---------------------------------------------------------------------------------------------
// CoffeeParticleSystem.apply_fluid_forces (coffee_particles.py:547-639): the producer of the integrator's `force`:
// nearest-cell fluid velocity, clamped Stokes drag, buoyancy and gravity, each with the reference's guards; a particle
// that fails a guard keeps the force it had.  One thread per particle, reference statement order, -fmad=false.
// vol_k = f32((4/3) * 3.14159), mu_safe = f32(max(1e-8, water_viscosity)): constant expressions the reference folds in f64.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) particles_fluid_forces_kernel(Grid G, const float *__restrict__ u, lbm_particles P, float *__restrict__ force,
                                                                     float rho_w, float mu_safe, float gravity, float vol_k, float max_coord,
                                                                     int *counters) {
    const int n = P.n;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || P.active[p] != 1) return;
    const float px = P.pos[p], py = P.pos[n + p], pz = P.pos[2 * n + p];
    if (!valid_coordinate(px, py, pz, max_coord)) {
        P.active[p] = 0;
        if (counters) atomicAdd(counters, 1);
        return;
    }
    const int gi = (int)fmaxf(0.0f, fminf((float)(G.nx - 2), px));
    const int gj = (int)fmaxf(0.0f, fminf((float)(G.ny - 2), py));
    const int gk = (int)fmaxf(0.0f, fminf((float)(G.nz_global - 2), pz));
    const int kl = gk - G.z0;
    if (kl < -G.zg || kl >= G.nz + G.zg) return;                     // another slab's cell
    const long long vol = G.vol, c = ((long long)(kl + G.zg) * G.ny + gj) * G.nx + gi;
    const float fx = u[c], fy = u[vol + c], fz = u[2 * vol + c];
    const float fspeed = norm3(fx, fy, fz);
    if (!(fspeed == fspeed && fspeed <= 100.0f)) return;
    float vx = P.vel[p], vy = P.vel[n + p], vz = P.vel[2 * n + p];
    if (!valid_velocity(vx, vy, vz)) { P.vel[p] = 0.0f; P.vel[n + p] = 0.0f; P.vel[2 * n + p] = 0.0f; vx = vy = vz = 0.0f; }
    const float rx = fx - vx, ry = fy - vy, rz = fz - vz;
    const float rs = norm3(rx, ry, rz);
    if (!(rs > 1e-6f && rs < 10.0f)) return;
    const float radius = P.radius[p], mass = P.mass[p];
    const bool radius_ok = !(radius < 1e-5f || radius > 0.01f || radius != radius);
    if (!(radius_ok && mass == mass && mass > 0.0f)) return;
    float re = (((rs * 2.0f) * radius) * rho_w) / mu_safe;
    re = fmaxf(0.01f, fminf(1000.0f, re));
    float cd = 24.0f / fmaxf(0.1f, re);
    cd = fmaxf(0.1f, fminf(10.0f, cd));
    float dm = ((((0.5f * cd) * 3.14159f) * (radius * radius)) * rho_w) * rs;
    dm = fminf(dm, mass * 100.0f);
    float dx = 0.0f, dy = 0.0f, dz = 0.0f;
    if (rs > 0.0f) { dx = dm * (rx / rs); dy = dm * (ry / rs); dz = dm * (rz / rs); }
    const float volume = vol_k * ((radius * radius) * radius);
    const float bm = fminf((volume * rho_w) * gravity, mass * 20.0f);
    const float gm = mass * gravity;
    const float tx = (dx + bm * 0.0f) + gm * 0.0f, ty = (dy + bm * 0.0f) + gm * 0.0f, tz = (dz + bm) + (-gm);
    const float fm = norm3(tx, ty, tz);
    if (fm == fm && fm < mass * 1000.0f) { force[p] = tx; force[n + p] = ty; force[2 * n + p] = tz; }
    else { force[p] = gm * 0.0f; force[n + p] = gm * 0.0f; force[2 * n + p] = -gm; }
}

#ifndef LBM_EMULATE_ON_HOST      /* tests/emu compiles the kernels with g++ and runs them thread by thread */
cudaError_t launch_particles_fluid_forces(const Grid &G, const float *u, const lbm_particles &ps, float *force, float rho_w, float mu_safe,
                                          float gravity, float vol_k, float max_coord, int *counters, cudaStream_t s) {
    if (ps.n <= 0) return cudaSuccess;
    particles_fluid_forces_kernel<<<(ps.n + 255) / 256, 256, 0, s>>>(G, u, ps, force, rho_w, mu_safe, gravity, vol_k, max_coord, counters);
    return cudaGetLastError();
}

cudaError_t launch_particles_couple(const Grid &G, const float *u, float *reaction, const lbm_particles &ps,
                                    float rho_w, float mu_w, float relax, cudaStream_t s) {
    ParticleArgs A{G, u, reaction, ps, rho_w, mu_w, relax};
    const int b = 256, gr = (ps.n + b - 1) / b;
    if (ps.n > 0) particles_couple_kernel<<<gr, b, 0, s>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_particles_advance(const lbm_particles &ps, float *force, const lbm_particle_bounds &b, float dt, int *counters, cudaStream_t s) {
    if (ps.n <= 0) return cudaSuccess;
    particles_advance_kernel<<<(ps.n + 255) / 256, 256, 0, s>>>(ps, force, b, dt, counters);
    return cudaGetLastError();
}
#endif

}  // namespace lbm
