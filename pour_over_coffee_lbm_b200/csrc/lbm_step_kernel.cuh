// The fused D3Q19 pull-scheme collide-stream kernel (sm_100a).
//
// One launch = one time step of LBMSolver.step() (legacy/lbm_solver.py:817-867):
//   stream-on-read (pull) of the 19 post-collision populations, halfway bounce-back
//   against the solid mask, macroscopic moments, body force (Guo), Smagorinsky LES,
//   filter-paper drag, BGK relaxation, 128-bit coalesced write-back.
// The kernel is HBM-bound (152 B of populations per cell update, ~2.7 flop/B); it uses no
// tensor cores.  Every population element is read exactly once and written exactly once
// per step, so loads/stores carry streaming (evict-first) hints.
//
// Template parameters
//   COMPAT  LBM_COMPAT_PHYSICAL | LBM_COMPAT_REFERENCE   
//   MODE    dense (fully periodic, no flags) | bulk (active-tile list, near-wall cells skipped) |
//           boundary (compact list of near-wall cells: bounce-back, open faces) -- see the kernel
//   FORCED  body_force / phase inputs active (either pointer may still be NULL)
//   LES     Smagorinsky: physical = local Pi^neq closed form, reference = FD on lagged u
//   POROUS  filter-zone drag: physical = Guo-Zhao force, reference = post-step u damping
//   VEC     cells per thread along x (1 or 4): VEC=4 uses aligned 128-bit loads and takes
//           the x+-1 neighbours of the shifted populations from the adjacent lane (shuffle).
// Arithmetic order follows oracle/d3q19_ref.py exactly; the translation unit is compiled
// twice, with -fmad=false ("strict", bit-exact against the oracle) and with FMA contraction.
#pragma once
#include <type_traits>
#include "lbm_common.cuh"
#include "lbm_phys.cuh"
#include "lbm_phys_chord.cuh"
#ifndef LBM_REF_GENERIC_COLLISION
#define LBM_REF_GENERIC_COLLISION 0      /* 1 (tests/emu only): the one-cell legacy collision runs through collide_reference_t<float> */
#endif
#ifdef LBM_EMULATE_ON_HOST
#define LBM_EMU_ALL_EDGE true            /* no neighbouring lane on the host: every thread fetches its own x -+ 1 words */
#else
#define LBM_EMU_ALL_EDGE false
#endif

namespace lbm {

template <int VEC> __device__ __forceinline__ void ld_stream(const float *p, float (&v)[VEC]);
template <> __device__ __forceinline__ void ld_stream<1>(const float *p, float (&v)[1]) { v[0] = __ldcs(p); }
template <> __device__ __forceinline__ void ld_stream<2>(const float *p, float (&v)[2]) {
    float2 t = __ldcs(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld_stream<4>(const float *p, float (&v)[4]) {
    float4 t = __ldcs(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <int VEC> __device__ __forceinline__ void ld_cached(const float *p, float (&v)[VEC]);
template <> __device__ __forceinline__ void ld_cached<1>(const float *p, float (&v)[1]) { v[0] = __ldg(p); }
template <> __device__ __forceinline__ void ld_cached<2>(const float *p, float (&v)[2]) {
    float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld_cached<4>(const float *p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <int VEC> __device__ __forceinline__ void st_stream(float *p, const float (&v)[VEC]);
template <> __device__ __forceinline__ void st_stream<1>(float *p, const float (&v)[1]) { __stcs(p, v[0]); }
template <> __device__ __forceinline__ void st_stream<2>(float *p, const float (&v)[2]) {
    __stcs(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
}
template <> __device__ __forceinline__ void st_stream<4>(float *p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
}

// Predicated scalar load (never a branch): the two edge lanes of a row segment fetch the x-+1 neighbour that no
// adjacent lane holds.  Written in PTX because the compiler otherwise turns the 10 conditional loads into divergent
// regions (+80 issue slots per warp, measured 0.408 -> 0.430 ms on the headline config).
__device__ __forceinline__ float ldg_if(const float *p, bool pred, float other) {
#ifdef LBM_EMULATE_ON_HOST
    return pred ? *p : other;
#else
    float v = other;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p ld.global.nc.f32 %0, [%1];\n\t}" : "+f"(v) : "l"(p), "r"((int)pred));
    return v;
#endif
}

struct CellAux {
    float Fx, Fy, Fz;   // body_force at the cell
    float phase;
    float nu_sgs;       // reference-mode FD eddy viscosity
    float blockage;
    unsigned flag;
    bool interior;      // 1 <= x,y,z <= N-2 in global coordinates
};
struct CellOut { float rho, ux, uy, uz; };

// ---------------------------------------------------------------------------------------------
// compat = reference: legacy/lbm_solver.py:488-628, 688-764; lbm_algorithms.py:183-218;
// filter_paper.py:538-614.  Same order of operations as oracle/ref_cpu.c.
// ---------------------------------------------------------------------------------------------
template <bool FORCED, bool LES, bool POROUS>
__device__ __forceinline__ void collide_reference(float (&f)[Q], const CellAux &a, CellOut &o, const StepArgs &P) {
    float rho = 0.0f;
    static_for<0, Q>([&](auto qq) { constexpr int q = decltype(qq)::value; rho += f[q]; });
    float mx = 0.0f, my = 0.0f, mz = 0.0f;
    static_for<0, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        if constexpr (cx(q) != 0) mx += f[q] * (float)cx(q);
        if constexpr (cy(q) != 0) my += f[q] * (float)cy(q);
        if constexpr (cz(q) != 0) mz += f[q] * (float)cz(q);
    });
    float Fx = 0.0f, Fy = 0.0f, Fz = 0.0f;
    const float ph = FORCED ? a.phase : 0.0f;
    if constexpr (FORCED) {
        const float gz = ph > 0.001f ? -(P.gravity_lu * ph) : 0.0f;
        Fx = 0.0f + a.Fx; Fy = 0.0f + a.Fy; Fz = gz + a.Fz;
    }
    float ux = 0.0f, uy = 0.0f, uz = 0.0f;
    if (rho > 1e-12f) {
        ux = (mx + 0.5f * Fx) / rho; uy = (my + 0.5f * Fy) / rho; uz = (mz + 0.5f * Fz) / rho;
    }
    float tau = ph > 0.5f ? P.tau_water : P.tau_air;
    if constexpr (LES) tau = tau + 3.0f * a.nu_sgs;
    tau = fmaxf(P.tau_min, fminf(P.tau_max, tau));
    const float omega = 1.0f / tau;
    // forcing prerequisites (identical for all q)
    bool forced = false;
    float tau_safe = 0.0f, fsx = 0.0f, fsy = 0.0f, fsz = 0.0f, usx = ux, usy = uy, usz = uz, uf = 0.0f;
    if constexpr (FORCED) {
        const float fnorm = sqrtf(dot3(Fx, Fy, Fz, Fx, Fy, Fz));
        forced = fnorm > 1e-15f;
        tau_safe = fminf(fmaxf(tau, 0.6f), 1.5f);
        const float sf = fnorm > 10.0f ? 10.0f / fnorm : 1.0f;
        fsx = Fx * sf; fsy = Fy * sf; fsz = Fz * sf;
        const float unorm = sqrtf(dot3(ux, uy, uz, ux, uy, uz));
        if (unorm > 0.2f) { const float s = 0.2f / unorm; usx = ux * s; usy = uy * s; usz = uz * s; }
        uf = dot3(usx, usy, usz, fsx, fsy, fsz);
    }
    const float u_sq = dot3(ux, uy, uz, ux, uy, uz);
    static_for<0, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        constexpr float w = wq(q);
        const float eu = edot<ex(q), ey(q), ez(q)>(ux, uy, uz);           // quirk Q1: equilibrium table
        const float feq = (w * rho) * (((1.0f + 3.0f * eu) + (4.5f * eu) * eu) - 1.5f * u_sq);
        float Fq = 0.0f;
        if constexpr (FORCED) {
            if (forced) {
                const float eus = edot<cx(q), cy(q), cz(q)>(usx, usy, usz);
                const float ef = edot<cx(q), cy(q), cz(q)>(fsx, fsy, fsz);
                const float coeff = w * (1.0f - 0.5f / tau_safe);
                Fq = coeff * (3.0f * ef + (9.0f * eus) * uf);
                Fq = fmaxf(-0.5f, fminf(0.5f, Fq));
            }
        }
        f[q] = (f[q] - omega * (f[q] - feq)) + Fq;
    });
    // externally visible u: filter damping applied after the step (quirk Q5)
    if constexpr (POROUS) {
        if ((a.flag & LBM_FLAG_FILTER) && a.interior) {
            const float umag = sqrtf(dot3(ux, uy, uz, ux, uy, uz));
            if (umag > 1e-8f && P.K_lu > 1e-12f) {
                const float darcy = P.c_darcy / P.K_lu;
                const float forch = ((P.c_forch * P.beta_lu) * umag) / sqrtf(P.K_lu);
                const float total = (darcy + forch) * (1.0f + a.blockage);
                float r = expf((-total) * 0.5f);
                r = fmaxf(0.1f, r);
                const float hf = (r + 1.0f) * 0.5f;
                uz = uz * r; ux = ux * hf; uy = uy * hf;
            }
        }
    }
    o.rho = rho; o.ux = ux; o.uy = uy; o.uz = uz;
}

// ---------------------------------------------------------------------------------------------
// The same collision on V = float (one cell) or V = P2 (two x-adjacent cells, packed f32x2): collide_reference above, statement by
// statement, with every product written as Ops<V>::mul0 -- a product that keeps its own rounding in front of the add / sub that
// consumes it (the packed mul.rn would be contracted into it by ptxas; lbm_phys.cuh) -- so both instantiations round exactly where
// the scalar code under -fmad=false does.  What has no packed instruction (divisions, square roots, min / max clamps, the
// per-cell conditions) runs lane by lane.  e . v of a direction with a negative leading component is evaluated as
// sigma * (magnitude) with the sign carried at compile time: -(a) + (-(b)) = -(a + b) and (-a) + b = b - a round identically, the
// products 3 eu and 9 eus uf are odd in it, (4.5 eu) eu is even, and the +-0.5 clamp of the Guo term is symmetric.  Differences
// against collide_reference are confined to the SIGN of exact zeros (and to NaN inputs), which no sum with a non-zero term keeps.
// tests/emu runs both instantiations on the CPU against the recorded runs of the reference (test_step_reference_emulated.py).
// ---------------------------------------------------------------------------------------------
template <class V> struct CellAuxT {
    V Fx, Fy, Fz, phase;
    float nu_sgs[Ops<V>::L], blockage[Ops<V>::L];
    unsigned flag[Ops<V>::L];
    bool interior[Ops<V>::L];
};
// magnitude and compile-time sign of e . v for e in {0, +1, -1}^3 with at most two non-zero components (D3Q19)
template <class V, int X, int Y, int Z> struct SignedDot {
    static constexpr int n = (X != 0) + (Y != 0) + (Z != 0);
    static_assert(n <= 2, "D3Q19: at most two non-zero components");
    static constexpr int first = X != 0 ? X : (Y != 0 ? Y : Z);                                  // sign of the leading term
    static constexpr int second = n < 2 ? 0 : (X != 0 ? (Y != 0 ? Y : Z) : Z);                   // sign of the other one
    static constexpr int sign = (n == 2 && first < 0 && second < 0) || (n == 1 && first < 0) ? -1 : 1;
    static __device__ __forceinline__ V mag(V vx, V vy, V vz) {
        using O = Ops<V>;
        if constexpr (n == 0) return O::bc(0.0f);
        else if constexpr (n == 1) return X != 0 ? vx : (Y != 0 ? vy : vz);
        else {
            const V a = X != 0 ? vx : vy, b = X != 0 ? (Y != 0 ? vy : vz) : vz;
            if constexpr (first > 0 && second > 0) return O::add(a, b);
            else if constexpr (first > 0) return O::sub(a, b);
            else if constexpr (second > 0) return O::sub(b, a);          // (-a) + b
            else return O::add(a, b);                                    // (-a) + (-b) = -(a + b)
        }
    }
};
template <class V, bool FORCED, bool LES, bool POROUS>
__device__ __forceinline__ void collide_reference_t(V (&f)[Q], const CellAuxT<V> &a, CellMacro<V> &o, const StepArgs &P) {
    using O = Ops<V>;
    constexpr int L = O::L;
    V rho = f[0];                                                        // 0 + f[0]
    static_for<1, Q>([&](auto qq) { constexpr int q = decltype(qq)::value; rho = O::add(rho, f[q]); });
    V mx = f[1], my = f[3], mz = f[5];                                   // the first term of each sum: 0 + f * (+1)
    static_for<2, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        if constexpr (cx(q) != 0) mx = cx(q) > 0 ? O::add(mx, f[q]) : O::sub(mx, f[q]);
        if constexpr (cy(q) != 0 && q > 3) my = cy(q) > 0 ? O::add(my, f[q]) : O::sub(my, f[q]);
        if constexpr (cz(q) != 0 && q > 5) mz = cz(q) > 0 ? O::add(mz, f[q]) : O::sub(mz, f[q]);
    });
    V Fx = O::bc(0.0f), Fy = O::bc(0.0f), Fz = O::bc(0.0f);
    float ph[L];
#pragma unroll
    for (int l = 0; l < L; ++l) ph[l] = FORCED ? O::get(a.phase, l) : 0.0f;
    if constexpr (FORCED) {
        float gz[L];
#pragma unroll
        for (int l = 0; l < L; ++l) gz[l] = ph[l] > 0.001f ? -(P.gravity_lu * ph[l]) : 0.0f;
        Fx = O::add(O::bc(0.0f), a.Fx); Fy = O::add(O::bc(0.0f), a.Fy); Fz = O::add(O::make(gz), a.Fz);
    }
    // per cell: velocity, relaxation time, the prerequisites of the clamped Guo term (legacy/lbm_solver.py:521-585)
    float uxl[L], uyl[L], uzl[L], oml[L], usx[L], usy[L], usz[L], fsx[L], fsy[L], fsz[L], ufl[L], c0[L], c1[L], c2[L];
    bool forced[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        const float r = O::get(rho, l), fx = O::get(Fx, l), fy = O::get(Fy, l), fz = O::get(Fz, l);
        float ux = 0.0f, uy = 0.0f, uz = 0.0f;
        if (r > 1e-12f) {
            ux = (O::get(mx, l) + 0.5f * fx) / r; uy = (O::get(my, l) + 0.5f * fy) / r; uz = (O::get(mz, l) + 0.5f * fz) / r;
        }
        float tau = ph[l] > 0.5f ? P.tau_water : P.tau_air;
        if constexpr (LES) tau = tau + 3.0f * a.nu_sgs[l];
        tau = fmaxf(P.tau_min, fminf(P.tau_max, tau));
        oml[l] = 1.0f / tau;
        uxl[l] = ux; uyl[l] = uy; uzl[l] = uz;
        forced[l] = false; usx[l] = ux; usy[l] = uy; usz[l] = uz; fsx[l] = fsy[l] = fsz[l] = 0.0f; ufl[l] = 0.0f; c0[l] = c1[l] = c2[l] = 0.0f;
        if constexpr (FORCED) {
            const float fnorm = sqrtf(dot3(fx, fy, fz, fx, fy, fz));
            forced[l] = fnorm > 1e-15f;
            const float tau_safe = fminf(fmaxf(tau, 0.6f), 1.5f);
            const float sf = fnorm > 10.0f ? 10.0f / fnorm : 1.0f;
            fsx[l] = fx * sf; fsy[l] = fy * sf; fsz[l] = fz * sf;
            const float unorm = sqrtf(dot3(ux, uy, uz, ux, uy, uz));
            if (unorm > 0.2f) { const float s = 0.2f / unorm; usx[l] = ux * s; usy[l] = uy * s; usz[l] = uz * s; }
            ufl[l] = dot3(usx[l], usy[l], usz[l], fsx[l], fsy[l], fsz[l]);
            const float pref = 1.0f - 0.5f / tau_safe;
            c0[l] = wq(0) * pref; c1[l] = wq(1) * pref; c2[l] = wq(7) * pref;
        }
    }
    const V ux = O::make(uxl), uy = O::make(uyl), uz = O::make(uzl), omega = O::make(oml);
    const V vsx = O::make(usx), vsy = O::make(usy), vsz = O::make(usz), vfx = O::make(fsx), vfy = O::make(fsy), vfz = O::make(fsz);
    const V uf = O::make(ufl), k0 = O::make(c0), k1 = O::make(c1), k2 = O::make(c2);
    bool any_forced = false;
#pragma unroll
    for (int l = 0; l < L; ++l) any_forced |= forced[l];
    const V u_sq = O::add(O::add(O::mul0(ux, ux), O::mul0(uy, uy)), O::mul0(uz, uz));
    const V k15 = O::mul0(O::bc(1.5f), u_sq);
    const V wr0 = O::mul0(O::bc(wq(0)), rho), wr1 = O::mul0(O::bc(wq(1)), rho), wr2 = O::mul0(O::bc(wq(7)), rho);
    static_for<0, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        using E = SignedDot<V, ex(q), ey(q), ez(q)>;                     // quirk Q1: the equilibrium's own table
        const V s = E::mag(ux, uy, uz);
        const V t3 = O::mul0(O::bc(3.0f), s);
        const V A = E::sign > 0 ? O::add(O::bc(1.0f), t3) : O::sub(O::bc(1.0f), t3);
        const V B = O::add(A, O::mul0(O::mul0(O::bc(4.5f), s), s));
        const V feq = O::mul0(q == 0 ? wr0 : (q < 7 ? wr1 : wr2), O::sub(B, k15));
        V fn = O::sub(f[q], O::mul0(omega, O::sub(f[q], feq)));
        if constexpr (FORCED) {
            if (any_forced) {
                using C = SignedDot<V, cx(q), cy(q), cz(q)>;
                const V eus = C::mag(vsx, vsy, vsz), ef = C::mag(vfx, vfy, vfz);
                const V g = O::add(O::mul0(O::bc(3.0f), ef), O::mul0(O::mul0(O::bc(9.0f), eus), uf));
                const V Fq = O::mul0(q == 0 ? k0 : (q < 7 ? k1 : k2), g);
                float fl[L];
#pragma unroll
                for (int l = 0; l < L; ++l) fl[l] = forced[l] ? fmaxf(-0.5f, fminf(0.5f, O::get(Fq, l))) : 0.0f;
                fn = C::sign > 0 ? O::add(fn, O::make(fl)) : O::sub(fn, O::make(fl));
            }
        }
        f[q] = fn;
    });
    // externally visible u: filter damping applied after the step (quirk Q5)
    if constexpr (POROUS) {
#pragma unroll
        for (int l = 0; l < L; ++l) {
            if ((a.flag[l] & LBM_FLAG_FILTER) && a.interior[l]) {
                const float umag = sqrtf(dot3(uxl[l], uyl[l], uzl[l], uxl[l], uyl[l], uzl[l]));
                if (umag > 1e-8f && P.K_lu > 1e-12f) {
                    const float darcy = P.c_darcy / P.K_lu;
                    const float forch = ((P.c_forch * P.beta_lu) * umag) / sqrtf(P.K_lu);
                    const float total = (darcy + forch) * (1.0f + a.blockage[l]);
                    float r = expf((-total) * 0.5f);
                    r = fmaxf(0.1f, r);
                    const float hf = (r + 1.0f) * 0.5f;
                    uzl[l] = uzl[l] * r; uxl[l] = uxl[l] * hf; uyl[l] = uyl[l] * hf;
                }
            }
        }
    }
    o.rho = rho; o.ux = O::make(uxl); o.uy = O::make(uyl); o.uz = O::make(uzl);
}

// ---------------------------------------------------------------------------------------------
// reference-mode LES pre-pass fused as a stencil read of the previous step's u
// (les_turbulence.py:318-380).  `c` = linear index of the cell in a scalar volume.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float les_fd_nu(const float *__restrict__ u, long long c, long long sy, long long sz,
                                           long long vol, float phase, float csd) {
    const float *ux = u, *uy = u + vol, *uz = u + 2 * vol;
    const float dudx = (__ldg(ux + c + 1) - __ldg(ux + c - 1)) * 0.5f;
    const float dudy = (__ldg(ux + c + sy) - __ldg(ux + c - sy)) * 0.5f;
    const float dudz = (__ldg(ux + c + sz) - __ldg(ux + c - sz)) * 0.5f;
    const float dvdx = (__ldg(uy + c + 1) - __ldg(uy + c - 1)) * 0.5f;
    const float dvdy = (__ldg(uy + c + sy) - __ldg(uy + c - sy)) * 0.5f;
    const float dvdz = (__ldg(uy + c + sz) - __ldg(uy + c - sz)) * 0.5f;
    const float dwdx = (__ldg(uz + c + 1) - __ldg(uz + c - 1)) * 0.5f;
    const float dwdy = (__ldg(uz + c + sy) - __ldg(uz + c - sy)) * 0.5f;
    const float dwdz = (__ldg(uz + c + sz) - __ldg(uz + c - sz)) * 0.5f;
    const float S11 = dudx, S22 = dvdy, S33 = dwdz;
    const float S12 = 0.5f * (dudy + dvdx), S13 = 0.5f * (dudz + dwdx), S23 = 0.5f * (dvdz + dwdy);
    const float mag = sqrtf(2.0f * (((S11 * S11 + S22 * S22) + S33 * S33) + 2.0f * ((S12 * S12 + S13 * S13) + S23 * S23)));
    if (mag < 1e-3f) return 0.0f;
    if (fabsf(phase) < 0.9f) return 0.0f;
    return fminf(csd * mag, 0.1f);
}
// ---------------------------------------------------------------------------------------------
// the kernel
//
// MODE selects how threads map to cells:
//   MODE_DENSE  every cell of planes [z_begin, z_end) -- fully periodic boxes without a flag field.
//   MODE_BULK   one WARP per entry of the active warp-tile list (32*VEC x-consecutive cells of one row holding at
//               least one fluid cell; the 65 % solid part of a V60 box is never launched).  The population loads
//               do not depend on the flag byte, so they are issued together with it (a dependent flag -> branch ->
//               load chain costs a full memory round trip per thread and halves the throughput at 16 warps/SM).
//               Near-wall cells (flag NEAR) fetch a precomputed 64-bit
//               neighbour mask and replace only the populations whose source is solid (halfway bounce-back: own
//               opposite post-collision population) or outside an open face (w_q) -- legacy/lbm_solver.py:609-628.
//               The coalesced 128-bit loads already brought everything else.
// ---------------------------------------------------------------------------------------------
enum { MODE_DENSE = 0, MODE_BULK = 1 };

// The work of one thread: VEC x-consecutive cells starting at (x0, y, z).  No early exit and no branch on the flag
// byte before the loads (see the comments inside).
template <int COMPAT, int MODE, bool FORCED, bool LES, bool POROUS, int VEC, bool COLLIDE, bool MRT = false>
__device__ __forceinline__ void step_cells(const StepArgs &P, const int x0, const int y, const int z, const bool active,
                                           const unsigned lane) {
    constexpr bool WALLS = MODE != MODE_DENSE;
    const Grid &G = P.g;
    constexpr unsigned FULL = 0xffffffffu;
    const int zp = z + G.zg;
    const long long own = ((long long)zp * G.ny + y) * G.nx + x0;

    // Raw flag word of this thread's cells.  It is only DECODED after every independent load below has been issued:
    // a flag -> branch -> load chain costs one full memory round trip per thread (measured: 0.42 -> 0.70 ms on an
    // all-fluid 256^3 box), so nothing may branch on the flags before the population / force / phase loads.
    unsigned flag_word = 0;
    if constexpr (WALLS) {
        if constexpr (VEC == 4) flag_word = __ldg(reinterpret_cast<const unsigned *>(P.flags + own));
        else if constexpr (VEC == 2) flag_word = __ldg(reinterpret_cast<const unsigned short *>(P.flags + own));
        else flag_word = __ldg(P.flags + own);
    }

    // neighbour rows / columns with periodic wrap (open faces clamp; such cells are NEAR and handled below)
    int ym = y - 1; if (ym < 0) ym = G.per_y ? G.ny - 1 : 0;
    int yq = y + 1; if (yq >= G.ny) yq = G.per_y ? 0 : G.ny - 1;
    int zm, zq;
    if (G.zg) { zm = zp - 1; zq = zp + 1; }
    else {
        zm = z - 1; if (zm < 0) zm = G.per_z ? G.nz - 1 : 0;
        zq = z + 1; if (zq >= G.nz) zq = G.per_z ? 0 : G.nz - 1;
    }
    const bool edge_lo = LBM_EMU_ALL_EDGE || lane == 0 || x0 == 0, edge_hi = LBM_EMU_ALL_EDGE || lane == 31 || x0 == G.nx - VEC;   // no lane holds my x-1 / x+VEC
    int xm = x0 - 1; if (xm < 0) xm = G.per_x ? G.nx - 1 : 0;
    int xq = x0 + VEC; if (xq >= G.nx) xq = G.per_x ? 0 : G.nx - 1;

    float f[Q][VEC];
    // (1) all loads first, straight-line: 19 aligned vector loads + (VEC > 1) the predicated scalar loads of the two
    //     edge lanes.  Nothing here may branch or consume a loaded value, or the memory-level parallelism collapses.
    float edge[Q];
    static_for<0, Q>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        const int rz = cz(q) > 0 ? zm : (cz(q) < 0 ? zq : zp);
        const int ry = cy(q) > 0 ? ym : (cy(q) < 0 ? yq : y);
        const float *row = P.src + (long long)q * G.vol + ((long long)rz * G.ny + ry) * G.nx;
        if constexpr (VEC == 1) {
            f[q][0] = __ldcs(row + (cx(q) > 0 ? xm : (cx(q) < 0 ? xq : x0)));
        } else {
            ld_stream<VEC>(row + x0, f[q]);
            if constexpr (cx(q) > 0) edge[q] = ldg_if(row + xm, edge_lo, 0.0f);
            if constexpr (cx(q) < 0) edge[q] = ldg_if(row + xq, edge_hi, 0.0f);
        }
    });
    // (2) shift the populations with cx != 0 by one cell: the x-+1 neighbour comes from the adjacent lane
    if constexpr (VEC > 1) {
        static_for<0, Q>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if constexpr (cx(q) > 0) {             // source is x-1: lane-1 holds it in its last element
                const float t = __shfl_up_sync(FULL, f[q][VEC - 1], 1);
#pragma unroll
                for (int c = VEC - 1; c > 0; --c) f[q][c] = f[q][c - 1];
                f[q][0] = edge_lo ? edge[q] : t;
            } else if constexpr (cx(q) < 0) {      // source is x+1: lane+1 holds it in its first element
                const float t = __shfl_down_sync(FULL, f[q][0], 1);
#pragma unroll
                for (int c = 0; c < VEC - 1; ++c) f[q][c] = f[q][c + 1];
                f[q][VEC - 1] = edge_hi ? edge[q] : t;
            }
        });
    }

    // auxiliary inputs: issued together with the populations (independent of the flags)
    float bf[3][VEC], ph[VEC];
    bool has_force = false, has_phase = false;
    if constexpr (FORCED) {
        has_phase = P.phase != nullptr;
        has_force = P.force != nullptr || (has_phase && P.gravity_lu != 0.0f);
        if (P.force != nullptr) {
#pragma unroll
            for (int d = 0; d < 3; ++d) ld_cached<VEC>(P.force + (long long)d * G.vol + own, bf[d]);
        } else {
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int c = 0; c < VEC; ++c) bf[d][c] = 0.0f;
        }
        if (has_phase) ld_cached<VEC>(P.phase + own, ph);
        else {
#pragma unroll
            for (int c = 0; c < VEC; ++c) ph[c] = 0.0f;
        }
    }

    // decode the flags: which of this thread's cells does this launch update?  (no early exit: lanes without fluid
    // cells run through with their stores predicated off -- an exit here would be hoisted above the loads)
    unsigned fl[VEC];
    bool mine[VEC];
    bool all_mine = true, any_near = false;
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
        fl[c] = WALLS ? ((flag_word >> (8 * c)) & 0xffu) : (unsigned)LBM_FLAG_LES;
        mine[c] = active && !(fl[c] & LBM_FLAG_SOLID);
        all_mine &= mine[c];
        any_near |= mine[c] && (fl[c] & LBM_FLAG_NEAR);
    }

    // halfway bounce-back + open-face inflow (legacy/lbm_solver.py:609-628) for near-wall cells: the neighbour mask
    // (low word: source x - e_q is solid, high word: source outside an open face) says which populations to replace.
    if constexpr (WALLS) {
        if (any_near) {
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                if (mine[c] && (fl[c] & LBM_FLAG_NEAR)) {
                    const unsigned long long m = __ldg(P.nbr + own + c);
                    const unsigned solid_bits = (unsigned)m, oob_bits = (unsigned)(m >> 32);
                    static_for<1, Q>([&](auto qq) {
                        constexpr int q = decltype(qq)::value;
                        if (oob_bits & (1u << q)) f[q][c] = wq(q);                    // stale inflow, SURVEY.md A.2-Q6
                        else if (solid_bits & (1u << q)) f[q][c] = __ldg(P.src + (long long)opp(q) * G.vol + own + c);
                    });
                }
            }
        }
    }

    CellOut out[VEC];
    if constexpr (COMPAT == LBM_COMPAT_PHYSICAL) {
        // compat = physical: lbm_phys.cuh (explicitly rounded operations; packed f32x2 on cell pairs when VEC is even)
        if constexpr (VEC % 2 == 0) {
#pragma unroll
            for (int c = 0; c < VEC; c += 2) {
                P2 fp[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) fp[q] = p2_make(f[q][c], f[q][c + 1]);
                CellIn<P2> in;
                in.Fx = FORCED ? p2_make(bf[0][c], bf[0][c + 1]) : p2_make(0.0f, 0.0f);
                in.Fy = FORCED ? p2_make(bf[1][c], bf[1][c + 1]) : p2_make(0.0f, 0.0f);
                in.Fz = FORCED ? p2_make(bf[2][c], bf[2][c + 1]) : p2_make(0.0f, 0.0f);
                in.phase = FORCED ? p2_make(ph[c], ph[c + 1]) : p2_make(0.0f, 0.0f);
                in.flag[0] = fl[c]; in.flag[1] = fl[c + 1];
                CellMacro<P2> m;
                collide_phys<P2, FORCED, LES && COLLIDE, POROUS, COLLIDE, false, MRT>(fp, in, m, P, has_phase, has_force);
                if constexpr (COLLIDE) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) { f[q][c] = p2_lo(fp[q]); f[q][c + 1] = p2_hi(fp[q]); }
                }
                out[c].rho = p2_lo(m.rho); out[c].ux = p2_lo(m.ux); out[c].uy = p2_lo(m.uy); out[c].uz = p2_lo(m.uz);
                out[c + 1].rho = p2_hi(m.rho); out[c + 1].ux = p2_hi(m.ux); out[c + 1].uy = p2_hi(m.uy); out[c + 1].uz = p2_hi(m.uz);
            }
        } else {
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                float fc[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) fc[q] = f[q][c];
                CellIn<float> in;
                in.Fx = FORCED ? bf[0][c] : 0.0f; in.Fy = FORCED ? bf[1][c] : 0.0f; in.Fz = FORCED ? bf[2][c] : 0.0f;
                in.phase = FORCED ? ph[c] : 0.0f;
                in.flag[0] = fl[c];
                CellMacro<float> m;
                collide_phys<float, FORCED, LES && COLLIDE, POROUS, COLLIDE, false, true>(fc, in, m, P, has_phase, has_force);
                if constexpr (COLLIDE) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) f[q][c] = fc[q];
                }
                out[c].rho = m.rho; out[c].ux = m.ux; out[c].uy = m.uy; out[c].uz = m.uz;
            }
        }
    } else if constexpr (VEC == 2 || LBM_REF_GENERIC_COLLISION) {
        // compat = reference on packed cell pairs (VEC = 2): collide_reference_t<P2>.  LBM_REF_GENERIC_COLLISION (tests/emu only) sends
        // the one-cell form through the same template, V = float, so its logic is checked against the recordings on the CPU.
        using V = std::conditional_t<VEC % 2 == 0, P2, float>;
        constexpr int L = Ops<V>::L;
#pragma unroll
        for (int c = 0; c < VEC; c += L) {
            CellAuxT<V> a;
            float fx[L], fy[L], fz[L], pl[L];
#pragma unroll
            for (int l = 0; l < L; ++l) {
                a.flag[l] = fl[c + l];
                fx[l] = FORCED ? bf[0][c + l] : 0.0f; fy[l] = FORCED ? bf[1][c + l] : 0.0f; fz[l] = FORCED ? bf[2][c + l] : 0.0f;
                pl[l] = FORCED ? ph[c + l] : 0.0f;
                a.blockage[l] = 0.0f; a.nu_sgs[l] = 0.0f;
                const int x = x0 + c + l, zg_ = G.z0 + z;
                a.interior[l] = x >= 1 && x <= G.nx - 2 && y >= 1 && y <= G.ny - 2 && zg_ >= 1 && zg_ <= G.nz_global - 2;
                if constexpr (POROUS) { if (P.blockage) a.blockage[l] = __ldg(P.blockage + own + c + l); }
                if constexpr (LES && COLLIDE) {
                    if (mine[c + l] && a.interior[l] && (fl[c + l] & LBM_FLAG_LES))
                        a.nu_sgs[l] = les_fd_nu(P.u_src, own + c + l, G.nx, G.plane, G.vol, pl[l], P.les_k);
                }
            }
            a.Fx = Ops<V>::make(fx); a.Fy = Ops<V>::make(fy); a.Fz = Ops<V>::make(fz); a.phase = Ops<V>::make(pl);
            V fp[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float t[L];
#pragma unroll
                for (int l = 0; l < L; ++l) t[l] = f[q][c + l];
                fp[q] = Ops<V>::make(t);
            }
            CellMacro<V> m;
            if constexpr (COLLIDE) collide_reference_t<V, FORCED, LES, POROUS>(fp, a, m, P);
            else collide_reference_t<V, FORCED, false, false>(fp, a, m, P);      // moments only: populations untouched below
#pragma unroll
            for (int l = 0; l < L; ++l) {
                if constexpr (COLLIDE) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) f[q][c + l] = Ops<V>::get(fp[q], l);
                }
                out[c + l].rho = Ops<V>::get(m.rho, l); out[c + l].ux = Ops<V>::get(m.ux, l);
                out[c + l].uy = Ops<V>::get(m.uy, l); out[c + l].uz = Ops<V>::get(m.uz, l);
            }
        }
    } else {
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
        CellAux a;
        a.flag = fl[c];
        a.Fx = FORCED ? bf[0][c] : 0.0f; a.Fy = FORCED ? bf[1][c] : 0.0f; a.Fz = FORCED ? bf[2][c] : 0.0f;
        a.phase = FORCED ? ph[c] : 0.0f;
        a.blockage = 0.0f; a.nu_sgs = 0.0f;
        const int x = x0 + c, zg_ = G.z0 + z;
        a.interior = x >= 1 && x <= G.nx - 2 && y >= 1 && y <= G.ny - 2 && zg_ >= 1 && zg_ <= G.nz_global - 2;
        if constexpr (POROUS) { if (P.blockage) a.blockage = __ldg(P.blockage + own + c); }
        if constexpr (LES && COLLIDE) {
            if (mine[c] && a.interior && (fl[c] & LBM_FLAG_LES))
                a.nu_sgs = les_fd_nu(P.u_src, own + c, G.nx, G.plane, G.vol, a.phase, P.les_k);
        }
        float fc[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) fc[q] = f[q][c];
        if constexpr (COLLIDE) {
            collide_reference<FORCED, LES, POROUS>(fc, a, out[c], P);
#pragma unroll
            for (int q = 0; q < Q; ++q) f[q][c] = fc[q];
        } else {
            // moments only (lbm_macroscopic): the collide routine on a scratch copy, populations untouched
            collide_reference<FORCED, false, false>(fc, a, out[c], P);
        }
    }
    }

    // write-back
    if constexpr (COLLIDE) {
        if constexpr (!WALLS) {
            if (active) {
                static_for<0, Q>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    st_stream<VEC>(P.dst + (long long)q * G.vol + own, f[q]);
                });
            }
        } else if (all_mine) {
            static_for<0, Q>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                st_stream<VEC>(P.dst + (long long)q * G.vol + own, f[q]);
            });
        } else {
#pragma unroll
            for (int c = 0; c < VEC; ++c)
                if (mine[c]) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) P.dst[(long long)q * G.vol + own + c] = f[q][c];
                }
        }
    }
    if (P.write_macro) {
        if (WALLS ? all_mine : active) {
            float r[VEC], a0[VEC], a1[VEC], a2[VEC];
#pragma unroll
            for (int c = 0; c < VEC; ++c) { r[c] = out[c].rho; a0[c] = out[c].ux; a1[c] = out[c].uy; a2[c] = out[c].uz; }
            st_stream<VEC>(P.rho + own, r);
            st_stream<VEC>(P.u_dst + own, a0);
            st_stream<VEC>(P.u_dst + G.vol + own, a1);
            st_stream<VEC>(P.u_dst + 2 * G.vol + own, a2);
        } else {
#pragma unroll
            for (int c = 0; c < VEC; ++c)
                if (mine[c]) {
                    P.rho[own + c] = out[c].rho;
                    P.u_dst[own + c] = out[c].ux; P.u_dst[G.vol + own + c] = out[c].uy; P.u_dst[2 * G.vol + own + c] = out[c].uz;
                }
        }
    }
}


// BUILD (0 = fast, 1 = strict/-fmad=false) only makes the two builds distinct symbols: without it the
// linker would merge the identically-named instantiations of the two translation units (ODR).
// MRT = true: the instantiation of the packed (VEC = 4) collision that honours lbm_params.mrt_magic (compat = physical); the BGK
// instantiations, which carry the roofline numbers, are built without it.  The one-cell form decides at run time.
template <int BUILD, int COMPAT, int MODE, bool FORCED, bool LES, bool POROUS, int VEC, int BLOCK, bool COLLIDE = true, int MINB = 1, bool MRT = false>
__global__ void __launch_bounds__(BLOCK, MINB) step_kernel(const __grid_constant__ StepArgs P) {
    const Grid &G = P.g;
    const unsigned lane = threadIdx.x & 31u;
    if constexpr (MODE == MODE_BULK) {
        // one warp per entry of the active warp-tile list; entry = x_segment | y << 8 | z << 20 (4 B, the list is kept
        // L2-resident through an access-policy window: a DRAM miss here would sit in front of every other load)
        const int w = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
        if (w >= P.n_items) return;
        const unsigned e = __ldg(P.items + P.item_begin + w);
        int x0 = (int)(e & 0xffu) * (32 * VEC) + (int)lane * VEC;
        const bool active = x0 < G.nx;
        if (!active) x0 = G.nx - VEC;                                    // duplicate of the last lane: loads stay in bounds
        step_cells<COMPAT, MODE, FORCED, LES, POROUS, VEC, COLLIDE, MRT>(P, x0, (int)((e >> 8) & 0xfffu), (int)(e >> 20), active, lane);
    } else {
        const int nxv = G.nx / VEC;
        const int per_plane = nxv * G.ny;
        int t = blockIdx.x * BLOCK + threadIdx.x;
        const int z = P.z_begin + (int)blockIdx.y * P.z_stride;
        const bool active = t < per_plane;
        if (!active) t = per_plane - 1;
        const int y = t / nxv;
        const int x0 = (t - y * nxv) * VEC;
        step_cells<COMPAT, MODE, FORCED, LES, POROUS, VEC, COLLIDE, MRT>(P, x0, y, z, active, lane);
    }
}

}  // namespace lbm
