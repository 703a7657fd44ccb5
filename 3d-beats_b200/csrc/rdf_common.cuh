// Shared device/host helpers for librdf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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
// One 32-byte header per node (two 128-bit loads), indexed by a GLOBAL node id g = t * nodes_per_tree + row with the
// canonical row order (row = 2^level - 1 + index, src/cuda/cu_utils.hpp:32-39):
//   a       = (ux, uy, vx, vy)
//   ithresh = ceil(thresh) as int32: the feature is an exact integer (difference of two uint16 probes), so
//             f < thresh  <=>  int(f) < ceil(thresh); NaN -> INT_MIN (never left), +-huge saturate.
//   left / right = where the walk goes for f < thresh / otherwise:
//             >= 0        global id of the child node (the reference's flag floor(l_next) == -1, tree_eval.cu:101-102)
//             <  0        ~leaf_id, leaf_id = 2 * g + side: the walk ends, its pdf row is pdf[leaf_id][CP]
//             RDF_NO_LEAF the walk falls off level D-1 with a "continue" flag and adds nothing (tree_eval.cu:95-128)
//   flags   = RDF_FLAG_EXACT_DIV when an offset is outside the domain of the reciprocal-based divide (below).
// Explicit child ids make a level cost one select instead of index arithmetic + flag tests, and leave the node order
// free (nodes may later be re-ordered for locality without touching the kernels).
// Leaf pdfs live in a separate table pdf[T * nodes_per_tree * 2][CP], CP = C rounded up to 4 floats (16-byte rows).
struct __align__(32) rdf_node_hdr {
    float4 a;
    int ithresh;
    int left, right;
    int flags;
};

#define RDF_NO_LEAF ((int)0x80000000)

// Node order inside a tree's slice of the packed arrays (free to choose: the headers carry explicit child ids).
//   RDF_LAYOUT_HEAP    canonical heap order, row = 2^level - 1 + index.
//   RDF_LAYOUT_BLOCKS  levels 0 .. RDF_PACK_TOP_LEVELS-1 in heap order (the part the kernels stage in shared memory), below that
//                      the levels are paired and every node of the upper level of a pair sits in ONE 128-byte line with its two
//                      children (slots 0, 1, 2; slot 3 unused), so the two consecutive levels of a divergent walk touch one line
//                      instead of two; an unpaired last level is stored densely.  The pdf table uses the same ids.
#define RDF_LAYOUT_HEAP 0
#define RDF_LAYOUT_BLOCKS 1
#ifndef RDF_PACK_TOP_LEVELS
#define RDF_PACK_TOP_LEVELS 6
#endif
#ifndef RDF_LAYOUT_DEFAULT
#define RDF_LAYOUT_DEFAULT RDF_LAYOUT_HEAP
#endif

// first slot of pair p (upper level RDF_PACK_TOP_LEVELS + 2p): 2^K + 4 * 2^K * (4^p - 1) / 3 - the 2^K - 1 heap rows are padded by
// one slot so that every block starts on a 128-byte line (tree slices are multiples of four slots, the arrays 256-byte aligned)
__host__ __device__ __forceinline__ int64_t rdf_blocks_pair_offset(int p) {
    const int64_t top = (int64_t)1 << RDF_PACK_TOP_LEVELS;
    return top + 4 * top * ((((int64_t)1 << (2 * p)) - 1) / 3);
}
__host__ __device__ __forceinline__ int64_t rdf_blocks_rows_per_tree(int D) {
    if (D <= RDF_PACK_TOP_LEVELS) return ((int64_t)1 << D) - 1;
    const int pairs = (D - RDF_PACK_TOP_LEVELS) / 2;
    const int64_t off = rdf_blocks_pair_offset(pairs);
    return ((D - RDF_PACK_TOP_LEVELS) & 1) ? off + ((int64_t)1 << (D - 1)) : off;
}
// slot of node (level j, index g within the level) inside its tree
__host__ __device__ __forceinline__ int64_t rdf_blocks_row(int j, int64_t g, int D) {
    if (j < RDF_PACK_TOP_LEVELS) return (((int64_t)1 << j) - 1) + g;
    const int p = (j - RDF_PACK_TOP_LEVELS) >> 1;
    const int L = RDF_PACK_TOP_LEVELS + 2 * p;
    const int64_t off = rdf_blocks_pair_offset(p);
    if (j == L) return L == D - 1 ? off + g : off + 4 * g;
    return off + 4 * (g >> 1) + 1 + (g & 1);
}

struct rdf_forest {
    int T, D, C, CP;
    int64_t nodes_per_tree;       // 2^D - 1 (canonical rows of a tree)
    int64_t rows_per_tree;        // header slots of a tree in the packed arrays (== nodes_per_tree in heap order, more with blocks)
    int layout;                   // RDF_LAYOUT_HEAP / RDF_LAYOUT_BLOCKS
    rdf_node_hdr* hdr;            // [T * rows_per_tree]
    float* pdf;                   // [T * rows_per_tree * 2][CP]
    size_t packed_bytes;
    int device;
    int has_exact_nodes;          // some node carries RDF_FLAG_EXACT_DIV (set by the last pack; read back at create / update)
    int has_early_leaves;         // some walk can end above level D-1 (0: every tree is complete, all walks take exactly D steps)
    int* exact_flag_dev;          // device word the pack kernel ORs into
    rdf_node_hdr* top_host;       // host copy of levels 0 .. top_levels-1 of every tree, [T][2^top_levels - 1], child ids rewritten to
                                  // this array's indexing (rdf_pack); travels to the eval kernel as a kernel parameter
    int top_levels;               // min(D, rdf_top_levels(T)); 0: none (more than RDF_FAST_MAX_TREES trees)
};

// Upper levels that travel as kernel parameters (constant bank): as many as RDF_EVAL_TOP_LEVELS allows and as fit the 32 764-byte
// parameter space next to the other launch parameters.
#ifndef RDF_EVAL_TOP_LEVELS
#define RDF_EVAL_TOP_LEVELS 5
#endif
__host__ __device__ constexpr int rdf_top_levels(int T) {
    int L = RDF_EVAL_TOP_LEVELS;
    while (L > 0 && (size_t)T * ((1u << L) - 1u) * 32u > 30000u) L--;
    return L;
}

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

// Packed-forest path: floor + "add the pixel coordinate" in ONE FMA-pipe instruction, no conversion unit.
// For |q| <= 2^21 and an integer coordinate 0 <= X < 2^16, the exact sum q + X + 1.5*2^23 lies in [2^23, 2^24) where
// floats are the integers, so add.rm.f32 (round toward -inf) returns exactly floor(q) + X + 1.5*2^23 and the low
// mantissa bits hold floor(q) + X: identical to cvt.rmi.s32.f32(q) + X of the reference.  (F2I runs on the
// quarter-rate XU pipe: four of them per node-step were 49 % XU utilisation in profiles/r01_ncu_eval_v1.md.)
#define RDF_MAGIC_F 12582912.0f          // 1.5 * 2^23
#define RDF_MAGIC_BITS 0x4B400000        // __float_as_int(RDF_MAGIC_F)
#define RDF_FASTFLOOR_MAX 2097152.0f     // 2^21: bound on |scale * u| for the path above

__host__ __device__ __forceinline__ bool rdf_fastfloor_domain(float a) {
    const float m = fabsf(a);
    return a == 0.f || (m >= 8.6736174e-19f && m <= RDF_FASTFLOOR_MAX);   // 2^-60 .. 2^21; false for NaN / inf
}

// coordinate + floor(RN(su / d)), cm = (float)coordinate + RDF_MAGIC_F
__device__ __forceinline__ int rdf_coord_fast(float su, float df, float rcp, float cm) {
    return __float_as_int(__fadd_rd(rdf_div_fast(su, df, rcp), cm)) - RDF_MAGIC_BITS;
}

// Two coordinates at once with the packed fp32 instructions of sm_100 (FMUL2 / FFMA2 / FADD2.RM: two independent IEEE fp32
// operations per instruction, so the results are bit-identical to rdf_coord_fast): x + floor(RN(ax / d)), y + floor(RN(ay / d)).
// Halves the issue slots of the divide sequence (16 -> 8 of ~48 instructions per node-step).
__device__ __forceinline__ void rdf_coord_fast2(float ax, float ay, float df, float rcp, float xm, float ym, int& cx, int& cy) {
    unsigned long long a, r, nd, cm, q0, rr, q1, sum;
    float sx, sy;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(ax), "f"(ay));
    asm("mov.b64 %0, {%1,%1};" : "=l"(r) : "f"(rcp));
    asm("mov.b64 %0, {%1,%1};" : "=l"(nd) : "f"(-df));
    asm("mov.b64 %0, {%1,%2};" : "=l"(cm) : "f"(xm), "f"(ym));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(q0) : "l"(a), "l"(r));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rr) : "l"(nd), "l"(q0), "l"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q1) : "l"(rr), "l"(r), "l"(q0));
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(sum) : "l"(q1), "l"(cm));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(sx), "=f"(sy) : "l"(sum));
    cx = __float_as_int(sx) - RDF_MAGIC_BITS;
    cy = __float_as_int(sy) - RDF_MAGIC_BITS;
}

// ceil(thresh) for the integer compare of the packed path (see rdf_node_hdr)
__host__ __device__ __forceinline__ int rdf_int_thresh(float t) {
    if (!(t == t)) return (int)0x80000000;                   // NaN: f < NaN is false for every f
    if (t >= 2147483648.f) return 0x7fffffff;
    if (t <= -2147483648.f) return (int)0x80000000;
    return (int)ceilf(t);
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

// img[idx] through the read-only path with ONE address instruction (mad.wide.u32): the image base stays in a register
// pair and the 32-bit pixel index is scaled and added in the FMA pipe (the compiler's own lowering spent five integer
// instructions per probe on 64-bit adds).
__device__ __forceinline__ unsigned rdf_ldg_u16(const uint16_t* __restrict__ img, unsigned idx) {
    unsigned short v;
    asm("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %1, 2, %2;\n\tld.global.nc.u16 %0, [a];\n\t}" : "=h"(v) : "r"(idx), "l"(img));
    return v;
}

// Packed path: the two probes of a node, then the feature as an exact integer, int(pu) - int(pv).
// xm / ym = (float)X / Y + RDF_MAGIC_F.  Split in two so that a caller can issue the probe loads of speculative nodes early.
struct rdf_probes {
    unsigned pu, pv;
};

template <bool EXACT>
__device__ __forceinline__ rdf_probes rdf_probe_pair(const uint16_t* __restrict__ img, int W, int H, int X, int Y, float df, float rcp,
                                                     float xm, float ym, float sux, float suy, float svx, float svy) {
    int ux, uy, vx, vy;
    if (EXACT) {
        ux = (int)((unsigned)X + (unsigned)rdf_offset_exact(sux, df));
        uy = (int)((unsigned)Y + (unsigned)rdf_offset_exact(suy, df));
        vx = (int)((unsigned)X + (unsigned)rdf_offset_exact(svx, df));
        vy = (int)((unsigned)Y + (unsigned)rdf_offset_exact(svy, df));
    } else {
        rdf_coord_fast2(sux, suy, df, rcp, xm, ym, ux, uy);
        rdf_coord_fast2(svx, svy, df, rcp, xm, ym, vx, vy);
    }
    rdf_probes pr;
    pr.pu = RDF_NO_PIXEL;
    pr.pv = RDF_NO_PIXEL;
    if ((unsigned)ux < (unsigned)W && (unsigned)uy < (unsigned)H) pr.pu = rdf_ldg_u16(img, (unsigned)(uy * W + ux));
    if ((unsigned)vx < (unsigned)W && (unsigned)vy < (unsigned)H) pr.pv = rdf_ldg_u16(img, (unsigned)(vy * W + vx));
    return pr;
}

template <bool EXACT>
__device__ __forceinline__ int rdf_feature_i(const uint16_t* __restrict__ img, int W, int H, int X, int Y, float df, float rcp,
                                             float xm, float ym, float sux, float suy, float svx, float svy) {
    const rdf_probes pr = rdf_probe_pair<EXACT>(img, W, H, X, Y, df, rcp, xm, ym, sux, suy, svx, svy);
    return (int)pr.pu - (int)pr.pv;
}

// scale domain for the fast path: |scale*u| stays normal when 2^-30 <= |scale| <= 2^30 and u is in its own domain
static inline bool rdf_scale_fast_ok(float s) {
    const float m = s < 0 ? -s : s;
    return m >= 9.3132257e-10f && m <= 1.0737418e9f;
}

// packed path (rdf_coord_fast): nodes are flagged for |u| <= 2^21 at pack time, so |scale| <= 1 keeps |scale*u| <= 2^21
static inline bool rdf_scale_fastfloor_ok(float s) {
    const float m = s < 0 ? -s : s;
    return m >= 9.3132257e-10f && m <= 1.0f;
}

static inline cudaStream_t rdf_stream(void* s) { return (cudaStream_t)s; }

// CTAs per SM the register allocation of the per-pixel forest kernels aims for, by number of interleaved trees (see rdf_eval.cu)
#ifndef RDF_EVAL_MIN_BLOCKS_T
#define RDF_EVAL_MIN_BLOCKS_T(T) ((T) <= 5 ? 4 : (T) == 6 ? 3 : 2)
#endif

// Opt-in dynamic shared memory above 48 KB is a PER-DEVICE function attribute: remember what was set on each device so that a
// process driving several GPUs (or switching devices) gets it on all of them.
#define RDF_MAX_DEVICES 64
static inline int rdf_current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= RDF_MAX_DEVICES) d = 0;
    return d;
}
// Number of SMs of the current device (148 on a B200), queried once per device: grid-stride kernels size their grids from it.
static inline int rdf_sm_count() {
    static int n__[RDF_MAX_DEVICES];
    const int d = rdf_current_device();
    if (n__[d] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148;
        n__[d] = v;
    }
    return n__[d];
}

// Environment switches are experiment knobs: each one is read once per process (function-local static of a per-site lambda),
// never on the per-frame launch path.
#define RDF_GETENV_ONCE(name) ([]() -> const char* { static const char* v__ = getenv(name); return v__; }())

#define RDF_ENSURE_DYN_SMEM(func, bytes)                                                                              \
    do {                                                                                                              \
        static size_t set__[RDF_MAX_DEVICES];                                                                         \
        const int d__ = rdf_current_device();                                                                         \
        if ((size_t)(bytes) > set__[d__]) {                                                                           \
            RDF_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));          \
            set__[d__] = (size_t)(bytes);                                                                             \
        }                                                                                                             \
    } while (0)
