// Micro-benchmark: L1 data-pipe wavefronts per load instruction for the access shapes of the forest-eval kernel.
// Run under ncu:  ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,
//                     l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum  ./l1_wavefronts
// Build on demand (the binary is a git-ignored artefact): nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o l1_wavefronts l1_wavefronts.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 256
#define W 848

template <int MODE>
__global__ void k_u16(const uint16_t* __restrict__ img, unsigned* out, int n) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned acc = 0;
    unsigned base = (warp * 977u) % (unsigned)(n - 64 * W);
    for (int i = 0; i < ITERS; i++) {
        unsigned idx;
        if (MODE == 0) idx = base + lane;                                   // 32 consecutive u16: 64 contiguous bytes
        else if (MODE == 1) idx = base + (lane & 7) + (lane >> 3) * W;      // 8x4 patch
        else if (MODE == 2) idx = base + (lane & 15) + (lane >> 4) * W;     // 16x2 patch
        else if (MODE == 3) idx = base + lane * 128;                        // 32 different lines
        else if (MODE == 4) idx = base;                                     // broadcast
        else idx = base + lane * 2;                                         // every other u16: 128 contiguous bytes
        acc += __ldg(img + idx);
        base = (base + 4099u + (acc & 1u)) % (unsigned)(n - 64 * W);
    }
    if (acc == 0xffffffffu) out[0] = acc;
}

template <int MODE>
__global__ void k_wide(const uint4* __restrict__ p, unsigned* out, int n16) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned acc = 0;
    unsigned base = (warp * 977u) % (unsigned)(n16 - 4096);
    for (int i = 0; i < ITERS; i++) {
        base &= ~1u;
        if (MODE == 0) {          // LDG.256, all lanes the same 32 bytes
            unsigned r[8];
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p + base));
            acc += r[0] + r[7];
        } else if (MODE == 1) {   // LDG.256, 32 consecutive 32-byte records
            unsigned r[8];
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p + base + 2 * lane));
            acc += r[0] + r[7];
        } else if (MODE == 2) {   // LDG.256, 32 scattered records
            unsigned r[8];
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p + base + 2 * ((lane * 37u) & 2047u)));
            acc += r[0] + r[7];
        } else if (MODE == 3) {   // LDG.128 same address
            const uint4 v = __ldg(p + base);
            acc += v.x + v.w;
        } else if (MODE == 4) {   // LDG.128 + LDG.64 same address (24-byte header)
            const uint4 v = __ldg(p + base);
            const uint2 w2 = __ldg(reinterpret_cast<const uint2*>(p + base + 1));
            acc += v.x + w2.y;
        } else {                  // LDG.32 same address
            acc += __ldg(reinterpret_cast<const unsigned*>(p + base));
        }
        base = (base + 4099u + (acc & 1u)) % (unsigned)(n16 - 4096);
    }
    if (acc == 0xffffffffu) out[0] = acc;
}

int main() {
    const int n = 64 << 20;                 // 128 MB of u16
    uint16_t* img; unsigned* out;
    cudaMalloc(&img, (size_t)n * 2); cudaMemset(img, 1, (size_t)n * 2); cudaMalloc(&out, 4);
    const int blocks = 148 * 4, threads = 256;
    k_u16<0><<<blocks, threads>>>(img, out, n); k_u16<1><<<blocks, threads>>>(img, out, n); k_u16<2><<<blocks, threads>>>(img, out, n);
    k_u16<3><<<blocks, threads>>>(img, out, n); k_u16<4><<<blocks, threads>>>(img, out, n); k_u16<5><<<blocks, threads>>>(img, out, n);
    const int n16 = n / 8;
    k_wide<0><<<blocks, threads>>>((const uint4*)img, out, n16); k_wide<1><<<blocks, threads>>>((const uint4*)img, out, n16);
    k_wide<2><<<blocks, threads>>>((const uint4*)img, out, n16); k_wide<3><<<blocks, threads>>>((const uint4*)img, out, n16);
    k_wide<4><<<blocks, threads>>>((const uint4*)img, out, n16); k_wide<5><<<blocks, threads>>>((const uint4*)img, out, n16);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s; %d warps x %d loads per kernel\n", cudaGetErrorString(e), blocks * threads / 32, ITERS);
    return e != cudaSuccess;
}
