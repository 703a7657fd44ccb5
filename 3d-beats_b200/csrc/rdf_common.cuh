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
// multiply, one IEEE fp32 divide (div.rn.f32) and cvt.rmi.s32.f32.
//
// Exact path: __fdiv_rn (about 15 SASS instructions per divide incl. the slow-path check; 4 divides per node-step
// were 55 % of all issued instructions in the first ncu capture, profiles/r01_*).
//
// Fast path, bit-identical to div.rn.f32 on its domain: the divisor is the same for all four divides of a pixel and
// for every node the pixel visits, so its correctly rounded reciprocal y = RN(1/d) is computed ONCE per pixel and each
// quotient is q0 = RN(a*y); r = fma(-d, q0, a) (exact); q1 = RN(q0 + r*y)   (Markstein's correction step).
// Why q1 == RN(a/d) exactly: d is an integer in [1, 65535] (a uint16 depth), a is any normal fp32 with
// 2^-60 <= |a| <= 2^60 (or +-0).  q0 is within 2 ulp of a/d, r is exactly representable, and the value rounded in the
// last step is a/d + delta with |delta| <= 2^-23 ulp.  For any rounding midpoint m (between adjacent floats near a/d),
// a and m*d are both multiples of ulp(a/d)/2, and a != m*d (a quotient of two 24-bit floats is never a midpoint), so
// |a/d - m| >= ulp/(2d) > 2^-17 ulp > |delta|: the perturbation cannot cross a midpoint and RN(a/d + delta) == RN(a/d).
// Values outside that domain (NaN, inf, denormal-range or huge offsets) are flagged per node at pack time
// (RDF_FLAG_EXACT_DIV) and take the exact path.  rdf_selftest_fastdiv() checks the identity on the GPU over billions of
// (a, d) pairs including the adversarial neighbourhood a ~ n*d +- few ulp.
#define RDF_FLAG_LEFT_CONT 1
#define RDF_FLAG_RIGHT_CONT 2
#define RDF_FLAG_EXACT_DIV 4

__host__ __device__ __forceinline__ bool rdf_fastdiv_domain(float a) {
    const float m = fabsf(a);
    return a == 0.f || (m >= 8.6736174e-19f && m <= 1.1529215e18f);     // 2^-60 .. 2^60; false for NaN / inf
}

__device__ __forceinline__ int rdf_offset_exact(float su, float df) { return __float2int_rd(__fdiv_rn(su, df)); }

__device__ __forceinline__ float rdf_div_fast(float a, float df, float rcp) {
    const float q0 = __fmul_rn(a, rcp);
    const float r = __fmaf_rn(-df, q0, a);
    return __fmaf_rn(r, rcp, q0);
}

__device__ __forceinline__ int rdf_offset_fast(float su, float df, float rcp) {
    return __float2int_rd(rdf_div_fast(su, df, rcp));
}

// Array3d<uint16>::get with default 65535 (src/cuda/cu_utils.hpp:58-62,79-86): bounds are per image.
__device__ __forceinline__ unsigned rdf_probe(const uint16_t* __restrict__ img, int W, int H, int x, int y) {
    unsigned v = RDF_NO_PIXEL;
    if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) v = __ldg(img + (size_t)y * W + x);
    return v;
}

// EXACT = true: __fdiv_rn; false: reciprocal path (caller guarantees the domain above).  rcp = __frcp_rn(df).
template <bool EXACT>
__device__ __forceinline__ float rdf_feature(const uint16_t* __restrict__ img, int W, int H, int X, int Y, float df, float rcp,
                                             float sux, float suy, float svx, float svy) {
    int oux, ouy, ovx, ovy;
    if (EXACT) {
        oux = rdf_offset_exact(sux, df); ouy = rdf_offset_exact(suy, df);
        ovx = rdf_offset_exact(svx, df); ovy = rdf_offset_exact(svy, df);
    } else {
        oux = rdf_offset_fast(sux, df, rcp); ouy = rdf_offset_fast(suy, df, rcp);
        ovx = rdf_offset_fast(svx, df, rcp); ovy = rdf_offset_fast(svy, df, rcp);
    }
    const int ux = (int)((unsigned)X + (unsigned)oux);
    const int uy = (int)((unsigned)Y + (unsigned)ouy);
    const int vx = (int)((unsigned)X + (unsigned)ovx);
    const int vy = (int)((unsigned)Y + (unsigned)ovy);
    const float pu = (float)rdf_probe(img, W, H, ux, uy);
    const float pv = (float)rdf_probe(img, W, H, vx, vy);
    return __fsub_rn(pu, pv);
}

// scale domain for the fast path: |scale*u| stays normal when 2^-30 <= |scale| <= 2^30 and u is in its own domain
static inline bool rdf_scale_fast_ok(float s) {
    const float m = s < 0 ? -s : s;
    return m >= 9.3132257e-10f && m <= 1.0737418e9f;
}

static inline cudaStream_t rdf_stream(void* s) { return (cudaStream_t)s; }
