// Shared device/host helpers for librdf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rdf_b200.h"

#define RDF_NO_PIXEL 65535u   // reference: MAX_UINT16, src/cuda/cu_utils.hpp:8

// ---- error plumbing --------------------------------------------------------------------------------------
void rdf_set_error(const char* fmt, ...);

#define RDF_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            rdf_set_error(__VA_ARGS__);             \
            return RDF_ERR_INVALID;                 \
        }                                           \
    } while (0)

#define RDF_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            rdf_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return RDF_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define RDF_LAUNCH_CHECK(name)                                                              \
    do {                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) {                                                           \
            rdf_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));        \
            return RDF_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

// ---- packed forest ---------------------------------------------------------------------------------------
// One 32-byte header per node, in the canonical row order (row = 2^level - 1 + index, src/cuda/cu_utils.hpp:32-39):
//   a = (ux, uy, vx, vy)      b = (thresh, flags, 0, 0);  flags bit0: left child continues (floor(l_next) == -1),
//   bit1: right child continues.  Leaf pdfs live in a separate table pdf[t][row][side][CP], CP = C rounded up to 4
//   floats so each leaf row is 16-byte aligned.
struct __align__(32) rdf_node_hdr {
    float4 a;
    float thresh;
    int flags;
    int pad0, pad1;
};

struct rdf_forest {
    int T, D, C, CP;
    int64_t nodes_per_tree;       // 2^D - 1
    rdf_node_hdr* hdr;            // [T][nodes_per_tree]
    float* pdf;                   // [T][nodes_per_tree][2][CP]
    size_t packed_bytes;
    int device;
};

// ---- reference arithmetic --------------------------------------------------------------------------------
// compute_feature (src/cuda/decision_tree_common.hpp:8-28): offsets are floor_rd( (scale*u) / float(d) ) with one fp32
// multiply, one IEEE fp32 divide and cvt.rmi.s32.f32 - __fmul_rn/__fdiv_rn keep that true under any compiler flag.
__device__ __forceinline__ int rdf_offset(float su, float df) { return __float2int_rd(__fdiv_rn(su, df)); }

// Array3d<uint16>::get with default 65535 (src/cuda/cu_utils.hpp:58-62,79-86): bounds are per image.
__device__ __forceinline__ unsigned rdf_probe(const uint16_t* __restrict__ img, int W, int H, int x, int y) {
    unsigned v = RDF_NO_PIXEL;
    if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) v = __ldg(img + (size_t)y * W + x);
    return v;
}

__device__ __forceinline__ float rdf_feature(const uint16_t* __restrict__ img, int W, int H, int X, int Y, float df,
                                             float sux, float suy, float svx, float svy) {
    const int ux = (int)((unsigned)X + (unsigned)rdf_offset(sux, df));
    const int uy = (int)((unsigned)Y + (unsigned)rdf_offset(suy, df));
    const int vx = (int)((unsigned)X + (unsigned)rdf_offset(svx, df));
    const int vy = (int)((unsigned)Y + (unsigned)rdf_offset(svy, df));
    const float pu = (float)rdf_probe(img, W, H, ux, uy);
    const float pv = (float)rdf_probe(img, W, H, vx, vy);
    return __fsub_rn(pu, pv);
}

static inline cudaStream_t rdf_stream(void* s) { return (cudaStream_t)s; }
