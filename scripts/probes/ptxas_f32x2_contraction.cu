// Probe: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 (and sub) into one FFMA2, with or without -fmad=false; the
// scalar mul.rn.f32 + add.rn.f32 pair is left alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -cubin -o t.cubin ptxas_f32x2_contraction.cu
//   cuobjdump -sass t.cubin | grep -E "Function|F(ADD|MUL|FMA)"
// prints FMUL + FADD for k3 and a single FFMA2 for k1 / k2.  Consequence for lbm_phys.cuh: no product may feed an
// add / sub directly; every such place is an explicit fma in the packed, scalar and oracle implementations.
struct P2 { unsigned long long v; };
__device__ __forceinline__ P2 add(P2 a, P2 b) { P2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 sub(P2 a, P2 b) { P2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 mul(P2 a, P2 b) { P2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__global__ void k1(const P2 *a, const P2 *b, const P2 *c, P2 *o) { int i = threadIdx.x; o[i] = add(mul(a[i], b[i]), c[i]); }
__global__ void k2(const P2 *a, const P2 *b, const P2 *c, P2 *o) { int i = threadIdx.x; o[i] = sub(a[i], mul(b[i], c[i])); }
__global__ void k3(const float *a, const float *b, const float *c, float *o) { int i = threadIdx.x; o[i] = __fadd_rn(__fmul_rn(a[i], b[i]), c[i]); }
