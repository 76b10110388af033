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
    const bool act = valid && P.active[p] != 0;

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

cudaError_t launch_particles_under_relax(const lbm_particles &ps, float relax, cudaStream_t s) {
    const int b = 256, gr = (ps.n + b - 1) / b;
    if (ps.n > 0) particles_under_relax_kernel<<<gr, b, 0, s>>>(ps, relax);
    return cudaGetLastError();
}

cudaError_t launch_particles_couple(const Grid &G, const float *u, float *reaction, const lbm_particles &ps,
                                    float rho_w, float mu_w, float relax, cudaStream_t s) {
    ParticleArgs A{G, u, reaction, ps, rho_w, mu_w, relax};
    const int b = 256, gr = (ps.n + b - 1) / b;
    if (ps.n > 0) particles_couple_kernel<<<gr, b, 0, s>>>(A);
    return cudaGetLastError();
}

}  // namespace lbm
