// Per-step producers of body_force / phase / rho that run next to the D3Q19 step:
//   MultiphaseFlow3D        src/core/multiphase_3d.py   surface-tension chain (:111-149, :313-332, :354-363) and the
//                                                       phase-field step (:151-197, :334-352, :383-387, :365-381)
//   PrecisePouringSystem    src/physics/precise_pouring.py   nozzle force (:131-163), gradual phase change (:165-196)
// compat = reference arithmetic: IEEE f32, the reference's evaluation order, compiled with -fmad=false (no contraction);
// bit-exact against oracle/producers_ref.py except the Gaussian of the nozzle profile (expf, <= 2 ulp).
//
// The reference runs the chain as 4 + 4 full-grid Taichi kernels (8 fields re-read between them); here it is 2 + 2
// launches, and the nozzle kernels visit the nozzle's bounding box (~10^2 cells) instead of the whole grid.
// All HBM-bound: cell loops on an (x-chunk, y, z) grid, x fastest, coalesced 4-byte accesses, no index divisions.
#include "lbm_common.cuh"

namespace lbm {

namespace {

__device__ __forceinline__ float norm3(float x, float y, float z) { return sqrtf(dot3(x, y, z, x, y, z)); }
// ti.max(-1.0, ti.min(1.0, v)) with the reference's operand order (a NaN in v survives, as in the reference)
__device__ __forceinline__ float clamp_pm1(float v) {
    const float t = (1.0f <= v) ? 1.0f : v;
    return (-1.0f >= t) ? -1.0f : t;
}
__device__ __forceinline__ bool interior(const Grid &G, int x, int y, int k) {
    return x >= 1 && x <= G.nx - 2 && y >= 1 && y <= G.ny - 2 && k >= 1 && k <= G.nz_global - 2;
}

// compute_chemical_potential, multiphase_3d.py:80-109 (run once by standardize_initial_state :542-571; the live step()
// never refreshes mu).  Both loops of the reference touch only the cell itself after the Laplacian, so they fuse.
__global__ void mp_chemical_potential_kernel(Grid G, const float *__restrict__ phi, float *__restrict__ laplacian, float *__restrict__ mu,
                                             float kappa) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, zp = blockIdx.z + G.zg;
    if (x >= G.nx || !interior(G, x, y, G.z0 + zp - G.zg)) return;
    const long long c = ((long long)zp * G.ny + y) * G.nx + x;
    const float p0 = phi[c];
    const float lap = (((((phi[c + 1] + phi[c - 1]) + phi[c + G.nx]) + phi[c - G.nx]) + phi[c + G.plane]) + phi[c - G.plane]) - 6.0f * p0;
    if (laplacian) laplacian[c] = lap;
    mu[c] = ((p0 * p0) * p0 - p0) + (-kappa) * lap;
}

// compute_gradients, multiphase_3d.py:111-132
__global__ void mp_gradients_kernel(Grid G, const float *__restrict__ phi, const float *__restrict__ mu, float *__restrict__ grad_phi,
                                    float *__restrict__ grad_mu, float *__restrict__ normal) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, zp = blockIdx.z + G.zg;
    if (x >= G.nx || !interior(G, x, y, G.z0 + zp - G.zg)) return;
    const long long n = G.vol, c = ((long long)zp * G.ny + y) * G.nx + x;
    const float gx = (phi[c + 1] - phi[c - 1]) * 0.5f;
    const float gy = (phi[c + G.nx] - phi[c - G.nx]) * 0.5f;
    const float gz = (phi[c + G.plane] - phi[c - G.plane]) * 0.5f;
    grad_phi[c] = gx; grad_phi[n + c] = gy; grad_phi[2 * n + c] = gz;
    if (mu && grad_mu) {
        grad_mu[c] = (mu[c + 1] - mu[c - 1]) * 0.5f;
        grad_mu[n + c] = (mu[c + G.nx] - mu[c - G.nx]) * 0.5f;
        grad_mu[2 * n + c] = (mu[c + G.plane] - mu[c - G.plane]) * 0.5f;
    }
    const float mag = norm3(gx, gy, gz);
    const bool ok = mag > 1e-10f;
    normal[c] = ok ? gx / mag : 0.0f; normal[n + c] = ok ? gy / mag : 0.0f; normal[2 * n + c] = ok ? gz / mag : 0.0f;
}

// compute_curvature :134-149 + compute_surface_tension_force :313-332 on interior cells, then apply_surface_tension
// :354-363 on every cell (the outer layer applies whatever surface_force holds there: it is never written).
__global__ void mp_curvature_force_kernel(Grid G, const float *__restrict__ phi, const float *__restrict__ rho, const uint8_t *__restrict__ flags,
                                          const float *__restrict__ grad_phi, const float *__restrict__ normal, float *__restrict__ curvature,
                                          float *__restrict__ surface_force, float *__restrict__ body_force, float sigma) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, zp = blockIdx.z + G.zg;
    if (x >= G.nx) return;
    const long long n = G.vol, c = ((long long)zp * G.ny + y) * G.nx + x;
    float sx, sy, sz;
    if (interior(G, x, y, G.z0 + zp - G.zg)) {
        const float nx0 = normal[c], ny0 = normal[n + c], nz0 = normal[2 * n + c];
        float curv = 0.0f;
        if (norm3(nx0, ny0, nz0) > 1e-10f) {
            const float dnx = (normal[c + 1] - normal[c - 1]) * 0.5f;
            const float dny = (normal[n + c + G.nx] - normal[n + c - G.nx]) * 0.5f;
            const float dnz = (normal[2 * n + c + G.plane] - normal[2 * n + c - G.plane]) * 0.5f;
            curv = (dnx + dny) + dnz;
        }
        curvature[c] = curv;
        sx = sy = sz = 0.0f;
        if (fabsf(phi[c]) < 0.9f) {
            const float grad_mag = norm3(grad_phi[c], grad_phi[n + c], grad_phi[2 * n + c]);
            if (grad_mag > 1e-10f) {
                const float fm = (sigma * curv) * grad_mag;
                sx = fm * nx0; sy = fm * ny0; sz = fm * nz0;
            }
        }
        surface_force[c] = sx; surface_force[n + c] = sy; surface_force[2 * n + c] = sz;
    } else {
        if (!body_force) return;
        sx = surface_force[c]; sy = surface_force[n + c]; sz = surface_force[2 * n + c];
    }
    if (!body_force || (flags[c] & LBM_FLAG_SOLID)) return;
    const float r = rho[c];
    if (r > 1e-10f) {
        body_force[c] = body_force[c] + sx / r; body_force[n + c] = body_force[n + c] + sy / r;
        body_force[2 * n + c] = body_force[2 * n + c] + sz / r;
    }
}

// apply_surface_tension :354-363 alone (MultiphaseFlow3D.step with precollision_applied = False re-applies a stored force)
__global__ void mp_apply_surface_tension_kernel(Grid G, const float *__restrict__ surface_force, const float *__restrict__ rho,
                                                const uint8_t *__restrict__ flags, float *__restrict__ body_force) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, zp = blockIdx.z + G.zg;
    if (x >= G.nx) return;
    const long long n = G.vol, c = ((long long)zp * G.ny + y) * G.nx + x;
    if (flags[c] & LBM_FLAG_SOLID) return;
    const float r = rho[c];
    if (r > 1e-10f) {
#pragma unroll
        for (int d = 0; d < 3; ++d) body_force[d * n + c] = body_force[d * n + c] + surface_force[d * n + c] / r;
    }
}

// update_phase_field_cahn_hilliard :151-197 + apply_phase_separation :334-352: both read phi (old) and write phi_new on
// interior cells, so they fuse; mu == nullptr is the all-zero field the live step() leaves it at.
__global__ void mp_phase_update_kernel(Grid G, const float *__restrict__ phi, const float *__restrict__ mu, const float *__restrict__ u,
                                       float *__restrict__ phi_new, float mobility, float dt) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, zp = blockIdx.z + G.zg;
    if (x >= G.nx || !interior(G, x, y, G.z0 + zp - G.zg)) return;
    const long long n = G.vol, c = ((long long)zp * G.ny + y) * G.nx + x;
    const float p0 = phi[c];
    const float pxm = phi[c - 1], pxp = phi[c + 1], pym = phi[c - G.nx], pyp = phi[c + G.nx], pzm = phi[c - G.plane], pzp = phi[c + G.plane];
    const float ux = u[c], uy = u[n + c], uz = u[2 * n + c];
    const float dx = ux > 0.0f ? p0 - pxm : pxp - p0;
    const float dy = uy > 0.0f ? p0 - pym : pyp - p0;
    const float dz = uz > 0.0f ? p0 - pzm : pzp - p0;
    const float convection = -((ux * dx + uy * dy) + uz * dz);
    float lap_mu = 0.0f;
    if (mu) lap_mu = (((((mu[c + 1] + mu[c - 1]) + mu[c + G.nx]) + mu[c - G.nx]) + mu[c + G.plane]) + mu[c - G.plane]) - 6.0f * mu[c];
    const float diffusion = mobility * lap_mu;
    float pn = clamp_pm1(p0 + dt * (convection + diffusion));
    if (fabsf(p0) < 0.99f) {
        const float lap = (((((pxp + pxm) + pyp) + pym) + pzp) + pzm) - 6.0f * p0;
        const float chem = p0 * (p0 * p0 - 1.0f) - 0.01f * lap;
        pn = pn + (-0.001f * chem) * dt;
    }
    phi_new[c] = pn;
}

// copy_phase_field :383-387 + update_density_from_phase :365-381 over every cell (src == dst: density only).
__global__ void mp_copy_density_kernel(Grid G, const float *src, float *dst, float *__restrict__ rho,
                                       float *__restrict__ phase, float rho_air, float drho) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, zp = blockIdx.z + G.zg;
    if (x >= G.nx) return;
    const long long c = ((long long)zp * G.ny + y) * G.nx + x;
    const float raw = src[c];
    if (dst != src) dst[c] = raw;
    const float p = clamp_pm1(raw);
    const float p1 = p + 1.0f;
    rho[c] = rho_air + (drho * p1) / 2.0f;
    phase[c] = p1 / 2.0f;
}

struct PourArgs {
    float pour_x, pour_y, radius; int pour_z;
    float velocity, flow_rate, dt;
    float decay[5];                 // exp(-d / 2.0), d = pour_z - k = 0..4: a constant expression in the reference (folded in f64)
    int x0, y0, k0, wx, wy, wk;     // bounding box of the nozzle
};
// _is_in_pouring_region, precise_pouring.py:99-129
__device__ __forceinline__ float pour_intensity(const PourArgs &P, int x, int y, int k) {
    const float dx = (float)x - P.pour_x, dy = (float)y - P.pour_y;
    const float dist = sqrtf(dx * dx + dy * dy);
    const int d = P.pour_z - k;
    if (!(dist <= P.radius) || d < 0 || d > 4) return 0.0f;
    const float t = dist / P.radius;
    return expf(-0.5f * (t * t)) * P.decay[d];
}
// apply_pouring_force :131-163 (MODE 0) and apply_gradual_phase_change :165-196 (MODE 1) over the nozzle's bounding box
template <int MODE>
__global__ void pour_kernel(Grid G, PourArgs P, const uint8_t *__restrict__ flags, float *__restrict__ field) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.wx * P.wy * P.wk) return;
    const int x = P.x0 + t % P.wx, y = P.y0 + (t / P.wx) % P.wy, k = P.k0 + t / (P.wx * P.wy);
    const int zp = k - G.z0 + G.zg;
    if (zp < G.zg || zp >= G.zg + G.nz) return;                       // another slab's plane
    const long long n = G.vol, c = ((long long)zp * G.ny + y) * G.nx + x;
    if (flags[c] & LBM_FLAG_SOLID) return;
    const float total = pour_intensity(P, x, y, k);
    if (!(total > 0.0f)) return;
    if (MODE == 0) {
        float accel = 0.0f;
        if (P.dt > 1e-8f) accel = ((P.velocity * total) * P.flow_rate) / P.dt;
        accel = accel <= 10.0f ? accel : 10.0f;
        field[c] = field[c] + 0.0f; field[n + c] = field[n + c] + 0.0f; field[2 * n + c] = field[2 * n + c] + (-accel);
    } else {
        const float cur = field[c];
        float rate = (1.0f - cur) / 0.05f;
        rate = (2.0f <= rate) ? 2.0f : rate;
        rate = (-2.0f >= rate) ? -2.0f : rate;
        const float change = ((rate * total) * P.dt) * P.flow_rate;
        field[c] = clamp_pm1(cur + change);
    }
}

// FilterPaperSystem.update_dynamic_resistance, filter_paper.py:703-746 (filter-zone cells):
// blockage <- 0.95 blockage + 0.05 * 0.9 (1 - exp(-0.1 accumulated)); accumulated *= 0.999.  The blockage field is an
// input of the step kernel's filter damping (apply_filter_effects :578-586).
__global__ void dynamic_resistance_kernel(Grid G, const uint8_t *__restrict__ flags, float *__restrict__ blockage, float *__restrict__ accumulated) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, zp = blockIdx.z + G.zg;
    if (x >= G.nx) return;
    const long long c = ((long long)zp * G.ny + y) * G.nx + x;
    if (!(flags[c] & LBM_FLAG_FILTER)) return;
    const float acc = accumulated[c];
    const float nb = 0.9f * (1.0f - expf(-0.1f * acc));
    blockage[c] = 0.95f * blockage[c] + 0.05f * nb;
    accumulated[c] = acc * 0.999f;
}

// Counter-based uniform [0, 1): the reference draws ti.random() from Taichi's unseeded per-thread generator, which no
// implementation can reproduce; here a draw is a pure function of (seed, particle, draw index) -- lowbias32 hash.
__device__ __forceinline__ float uniform01(unsigned seed, unsigned p, unsigned d) {
    unsigned h = seed ^ (p * 0x9E3779B9u) ^ (d * 0x85EBCA6Bu);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}
// FilterPaperSystem.block_particles_at_filter, filter_paper.py:616-700: a particle over a filter-zone cell (5 planes
// around its own) that moves down bounces with restitution 0.3, gets a small horizontal kick, and leaves 0.01 in
// accumulated_particles.  The reference divides the (lattice-unit) position by SCALE_LENGTH here (quirk Q9: mixed
// units) -- reproduced, the caller passes the divisor.
__global__ void particles_block_at_filter_kernel(Grid G, lbm_particles P, const uint8_t *__restrict__ flags, float *__restrict__ accumulated,
                                                 float scale_length, float noise, unsigned seed) {
    const int n = P.n, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || P.active[p] == 0) return;
    const int gx = (int)(P.pos[p] / scale_length), gy = (int)(P.pos[n + p] / scale_length), gz = (int)(P.pos[2 * n + p] / scale_length);
    if (gx < 0 || gx >= G.nx || gy < 0 || gy >= G.ny || gz < 0 || gz >= G.nz_global) return;
    for (int off = -2; off <= 2; ++off) {
        const int k = gz + off;
        if (k < 0 || k >= G.nz_global) continue;
        const int zp = k - G.z0 + G.zg;
        if (zp < 0 || zp >= G.nz + 2 * G.zg) continue;                 // beyond this slab's planes
        const long long c = ((long long)zp * G.ny + gy) * G.nx + gx;
        if (!(flags[c] & LBM_FLAG_FILTER)) continue;
        const float vz = P.vel[2 * n + p];
        if (vz < 0.0f) {
            P.vel[2 * n + p] = (-vz) * 0.3f;
            P.vel[p] = P.vel[p] + (uniform01(seed, (unsigned)p, 0u) - 0.5f) * noise;
            P.vel[n + p] = P.vel[n + p] + (uniform01(seed, (unsigned)p, 1u) - 0.5f) * noise;
            atomicAdd(accumulated + c, 0.01f);
        }
        break;
    }
}

inline dim3 cell_grid(const Grid &G, int b) { return dim3((unsigned)((G.nx + b - 1) / b), (unsigned)G.ny, (unsigned)G.nz); }
inline int cell_block(const Grid &G) { return G.nx >= 128 ? 128 : 64; }

}  // namespace

cudaError_t launch_chemical_potential(const Grid &G, const float *phi, float *laplacian, float *mu, float kappa, cudaStream_t s) {
    if (G.ny > 65535 || G.nz > 65535) return cudaErrorInvalidValue;
    const int b = cell_block(G);
    mp_chemical_potential_kernel<<<cell_grid(G, b), b, 0, s>>>(G, phi, laplacian, mu, kappa);
    return cudaGetLastError();
}
cudaError_t launch_surface_tension(const Grid &G, const float *phi, const float *mu, const float *rho, const uint8_t *flags, float *grad_phi,
                                   float *grad_mu, float *normal, float *curvature, float *surface_force, float *body_force, float sigma,
                                   cudaStream_t s) {
    if (G.ny > 65535 || G.nz > 65535) return cudaErrorInvalidValue;
    const int b = cell_block(G);
    mp_gradients_kernel<<<cell_grid(G, b), b, 0, s>>>(G, phi, mu, grad_phi, grad_mu, normal);
    mp_curvature_force_kernel<<<cell_grid(G, b), b, 0, s>>>(G, phi, rho, flags, grad_phi, normal, curvature, surface_force, body_force, sigma);
    return cudaGetLastError();
}
cudaError_t launch_apply_surface_tension(const Grid &G, const float *surface_force, const float *rho, const uint8_t *flags, float *body_force,
                                         cudaStream_t s) {
    if (G.ny > 65535 || G.nz > 65535) return cudaErrorInvalidValue;
    const int b = cell_block(G);
    mp_apply_surface_tension_kernel<<<cell_grid(G, b), b, 0, s>>>(G, surface_force, rho, flags, body_force);
    return cudaGetLastError();
}
cudaError_t launch_phase_field_step(const Grid &G, float *phi, float *phi_new, const float *mu, const float *u, float *rho, float *phase,
                                    float mobility, float dt, float rho_air, float drho, cudaStream_t s) {
    if (G.ny > 65535 || G.nz > 65535) return cudaErrorInvalidValue;
    const int b = cell_block(G);
    mp_phase_update_kernel<<<cell_grid(G, b), b, 0, s>>>(G, phi, mu, u, phi_new, mobility, dt);
    mp_copy_density_kernel<<<cell_grid(G, b), b, 0, s>>>(G, phi_new, phi, rho, phase, rho_air, drho);
    return cudaGetLastError();
}
cudaError_t launch_density_from_phase(const Grid &G, const float *phi, float *rho, float *phase, float rho_air, float drho, cudaStream_t s) {
    if (G.ny > 65535 || G.nz > 65535) return cudaErrorInvalidValue;
    const int b = cell_block(G);
    mp_copy_density_kernel<<<cell_grid(G, b), b, 0, s>>>(G, phi, const_cast<float *>(phi), rho, phase, rho_air, drho);
    return cudaGetLastError();
}
cudaError_t launch_dynamic_resistance(const Grid &G, const uint8_t *flags, float *blockage, float *accumulated, cudaStream_t s) {
    if (G.ny > 65535 || G.nz > 65535) return cudaErrorInvalidValue;
    const int b = cell_block(G);
    dynamic_resistance_kernel<<<cell_grid(G, b), b, 0, s>>>(G, flags, blockage, accumulated);
    return cudaGetLastError();
}
cudaError_t launch_particles_block_at_filter(const Grid &G, const lbm_particles &ps, const uint8_t *flags, float *accumulated, float scale_length,
                                             float noise, unsigned seed, cudaStream_t s) {
    const int b = 256, gr = (ps.n + b - 1) / b;
    if (ps.n > 0) particles_block_at_filter_kernel<<<gr, b, 0, s>>>(G, ps, flags, accumulated, scale_length, noise, seed);
    return cudaGetLastError();
}
// mode 0: body_force (3 components), mode 1: phi.  *launched = 0 when the nozzle's box misses this slab.
cudaError_t launch_pour(const Grid &G, const lbm_pour &pr, const float decay[5], int mode, const uint8_t *flags, float *field, cudaStream_t s,
                        int *launched) {
    PourArgs P{};
    P.pour_x = pr.pour_x; P.pour_y = pr.pour_y; P.radius = pr.radius; P.pour_z = pr.pour_z;
    P.velocity = pr.velocity; P.flow_rate = pr.flow_rate; P.dt = pr.dt;
    for (int i = 0; i < 5; ++i) P.decay[i] = decay[i];
    const int x0 = max(0, (int)floorf(pr.pour_x - pr.radius)), x1 = min(G.nx - 1, (int)ceilf(pr.pour_x + pr.radius));
    const int y0 = max(0, (int)floorf(pr.pour_y - pr.radius)), y1 = min(G.ny - 1, (int)ceilf(pr.pour_y + pr.radius));
    const int k0 = max(max(0, G.z0), pr.pour_z - 4), k1 = min(min(G.nz_global - 1, G.z0 + G.nz - 1), pr.pour_z);
    *launched = 0;
    if (x1 < x0 || y1 < y0 || k1 < k0) return cudaSuccess;
    P.x0 = x0; P.y0 = y0; P.k0 = k0; P.wx = x1 - x0 + 1; P.wy = y1 - y0 + 1; P.wk = k1 - k0 + 1;
    const long long cells = (long long)P.wx * P.wy * P.wk;
    if (cells > (1LL << 30)) return cudaErrorInvalidValue;
    const int b = 128, gr = (int)((cells + b - 1) / b);
    if (mode == 0) pour_kernel<0><<<gr, b, 0, s>>>(G, P, flags, field);
    else pour_kernel<1><<<gr, b, 0, s>>>(G, P, flags, field);
    *launched = 1;
    return cudaGetLastError();
}

}  // namespace lbm
