// Micro-benchmark 2: L1 data-pipe wavefronts of a node-header load when the 32 lanes of a warp sit on K distinct nodes
// (the forest-eval kernel's situation on smooth frames: 1-4 distinct nodes per warp, tools/uniform_stats.py).
// Run under ncu:  ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,
//                     l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum  ./l1_groups
// Build on demand (git-ignored artefact): nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o l1_groups l1_groups.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 256

// SHAPE 0: one LDG.256 per lane; 1: two LDG.128 per lane; 2: K rounds of two predicated LDG.128 (each round all active lanes read
// the same header); 3: one LDG.128 (first half of the header only); ADJ: the K nodes are adjacent 32-byte records (siblings) or scattered
template <int SHAPE, int K, bool ADJ>
__global__ void k_groups(const uint4* __restrict__ p, unsigned* out, int n16) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned acc = 0;
    unsigned base = (warp * 977u) % (unsigned)(n16 - 8192);
    const unsigned grp = (unsigned)lane % K;                       // lanes interleaved over the K nodes
    for (int i = 0; i < ITERS; i++) {
        base &= ~1u;
        const uint4* q = p + base + 2 * (ADJ ? grp : ((grp * 37u) & 2047u));
        if (SHAPE == 0) {
            unsigned r[8];
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(q));
            acc += r[0] + r[7];
        } else if (SHAPE == 1) {
            const uint4 a = __ldg(q), b = __ldg(q + 1);
            acc += a.x + b.w;
        } else if (SHAPE == 2) {
            uint4 a = make_uint4(0, 0, 0, 0), b = a;
#pragma unroll
            for (int g = 0; g < K; g++) {
                asm volatile("{\n\t.reg .pred pp;\n\tsetp.eq.u32 pp, %8, %9;\n\t@pp ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%10];\n\t"
                             "@pp ld.global.nc.v4.b32 {%4,%5,%6,%7}, [%10+16];\n\t}"
                             : "+r"(a.x), "+r"(a.y), "+r"(a.z), "+r"(a.w), "+r"(b.x), "+r"(b.y), "+r"(b.z), "+r"(b.w)
                             : "r"(grp), "r"((unsigned)g), "l"(q));
            }
            acc += a.x + b.w;
        } else {
            const uint4 a = __ldg(q);
            acc += a.x + a.w;
        }
        base = (base + 4099u + (acc & 1u)) % (unsigned)(n16 - 8192);
    }
    if (acc == 0xffffffffu) out[0] = acc;
}

#define RUN(S, K, A) k_groups<S, K, A><<<blocks, threads>>>((const uint4*)img, out, n16)

int main() {
    const int n = 64 << 20;
    uint16_t* img; unsigned* out;
    cudaMalloc(&img, (size_t)n * 2); cudaMemset(img, 1, (size_t)n * 2); cudaMalloc(&out, 4);
    const int blocks = 148 * 4, threads = 256, n16 = n / 8;
    // launch order = row order of profiles/r02_micro_l1_groups.md
    RUN(0, 1, true); RUN(0, 2, true); RUN(0, 2, false); RUN(0, 4, true); RUN(0, 4, false); RUN(0, 8, false); RUN(0, 32, false);
    RUN(1, 1, true); RUN(1, 2, true); RUN(1, 2, false); RUN(1, 4, true); RUN(1, 4, false); RUN(1, 8, false); RUN(1, 32, false);
    RUN(2, 1, true); RUN(2, 2, true); RUN(2, 2, false); RUN(2, 4, true); RUN(2, 4, false);
    RUN(3, 1, true); RUN(3, 2, false); RUN(3, 4, false); RUN(3, 32, false);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s; %d warps x %d header loads per kernel\n", cudaGetErrorString(e), blocks * threads / 32, ITERS);
    return e != cudaSuccess;
}
