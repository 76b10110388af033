// Tool-chain probe (no GPU needed): does ptxas 12.9 contract a packed f32x2 product into the add that consumes it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -c packed_contraction_probe.cu -o p.o && cuobjdump -sass p.o | grep -E "Function|FFMA2|FMUL2|FADD2|LOP3"
// Result (profiles/r02_probe_packed_contraction.txt): mul.rn.f32x2 + add.rn.f32x2 -> ONE FFMA2, also through volatile asm and through
// fma(a, b, -0.0); fma(a, b, +0.0) + add stays two instructions (FFMA2 with RZ, FADD2) but turns a -0 product into +0; a run-time
// XOR with zero between the two blocks the contraction at the price of two LOP3.
#include <cuda_runtime.h>
struct P2 { unsigned long long v; };
__device__ __forceinline__ P2 mul(P2 a, P2 b) { P2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 add(P2 a, P2 b) { P2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 fma(P2 a, P2 b, P2 c) { P2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__global__ void k_muladd(const P2 *a, const P2 *b, const P2 *c, P2 *o) { int i = threadIdx.x; o[i] = add(mul(a[i], b[i]), c[i]); }
__global__ void k_fma0add(const P2 *a, const P2 *b, const P2 *c, P2 *o) {
    int i = threadIdx.x; P2 nz; nz.v = 0x8000000080000000ull;      // (-0.0f, -0.0f)
    o[i] = add(fma(a[i], b[i], nz), c[i]);
}
__global__ void k_fmap0add(const P2 *a, const P2 *b, const P2 *c, P2 *o) {
    int i = threadIdx.x; P2 z; z.v = 0ull;
    o[i] = add(fma(a[i], b[i], z), c[i]);
}
__global__ void k_volatile(const P2 *a, const P2 *b, const P2 *c, P2 *o) {
    int i = threadIdx.x; P2 m = mul(a[i], b[i]);
    asm volatile("" : "+l"(m.v));
    o[i] = add(m, c[i]);
}
__global__ void k_volmul(const P2 *a, const P2 *b, const P2 *c, P2 *o) {
    int i = threadIdx.x; P2 m;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(m.v) : "l"(a[i].v), "l"(b[i].v));
    P2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(m.v), "l"(c[i].v));
    o[i] = r;
}
__global__ void k_xor(const P2 *a, const P2 *b, const P2 *c, P2 *o, unsigned long long zero) {
    int i = threadIdx.x; P2 m = mul(a[i], b[i]);
    m.v ^= zero;                      // a run-time zero: ptxas cannot see through it
    o[i] = add(m, c[i]);
}
