// Probe: which start coordinates does a tiled-mode TMA load (cp.async.bulk.tensor.4d) accept on B200?
//   ./tma_probe <x> <y>    loads the 64 x 4 box at (x, y, 1, 2) of a f32 [3][4][64][256] tensor and prints a checksum.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, float *out) {
    __shared__ __align__(128) float buf[4 * 64];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4 * 64 * 4));
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(d), "l"(&map), "r"(b), "r"(x), "r"(y), "r"(1), "r"(2) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char **argv) {
    const int x = argc > 1 ? atoi(argv[1]) : 0, y = argc > 2 ? atoi(argv[2]) : 0;
    const int nx = 256, ny = 64, nz = 4, nc = 3;
    std::vector<float> h((size_t)nx * ny * nz * nc);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 256 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    cuuint64_t dims[4] = {nx, ny, nz, nc}, strides[3] = {nx * 4ull, nx * ny * 4ull, (cuuint64_t)nx * ny * nz * 4ull};
    cuuint32_t box[4] = {64, 4, 1, 1}, es[4] = {1, 1, 1, 1};
    cuInit(0);
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    probe<<<1, 128>>>(map, x, y, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("x=%d y=%d: %s\n", x, y, cudaGetErrorString(e)); return 0; }
    float res[256]; cudaMemcpy(res, o, sizeof res, cudaMemcpyDeviceToHost);
    // expected value of element (i, j): tensor(x+i, y+j, 1, 2) or 0 outside
    int bad = 0;
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 64; ++i) {
        const int xx = x + i, yy = y + j;
        float want = 0.0f;
        if (xx >= 0 && xx < nx && yy >= 0 && yy < ny) want = h[(((size_t)2 * nz + 1) * ny + yy) * nx + xx];
        bad += res[j * 64 + i] != want;
    }
    printf("x=%d y=%d: ok, %d mismatches\n", x, y, bad);
    return 0;
}
