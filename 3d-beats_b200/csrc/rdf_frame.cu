// Live-frame conditioning around the forest: what src/3d_bz.py does to a depth frame immediately
// before LayeredDecisionForest.run and to the mean-shift centroids immediately after, without the reference's 8 launches + 2
// full-frame copies per frame, 5 launches + 2 copies per hand, and the per-fingertip host lookups.
//
//   rdf_condition_depth  = deproject_points + transform_points + filter_points_by_plane +
//                          remove_missing_3d_points_from_depth_image + (copy) + gaussian_depth_filter + shrink_image
//                          (src/3d_bz.py:159-220; src/cuda/points_ops.cu:5-36,66-75,131-146,327-373,375-404;
//                           src/cuda/calibrated_plane.cu:30-45)                                       -> ONE launch
//   rdf_stencil_hands    = fill(0) + [grow_groups] + stencil_depth_image_by_group + flip_x / copy + convert_0s_to_maxuint
//                          for every hand (src/3d_bz.py:252-259,390-420; points_ops.cu:118-129,407-483)  -> ONE launch
//   rdf_fingertip_z      = centroid -> raw depth lookup -> deproject -> plane space -> -z, per fingertip
//                          (src/3d_bz.py:503-522)                                                      -> ONE launch
//   rdf_grow_groups, rdf_flip_x, rdf_labels_to_rgba, rdf_depth_to_rgba: the remaining single kernels of that loop
//                          (points_ops.cu:407-438,466-483,258-281,283-325), kept for API parity / display.
//
// Bit-exactness: the 3-D point never has to exist in memory - only its plane-space z and w decide whether a depth sample
// survives - but it has to be computed with the reference's fp32 operation order.  That order is pinned on the SASS nvcc 12.9
// emits for the reference's kernels (with GLM's mat4 * vec4 = (m0*x + m1*y) + (m2*z + m3*w) and w known to be 1):
//     p   = ( RN(RN(d*(x-ppx))/f), RN(RN(d*(y-ppy))/f), d )                       deproject_points
//     z'  = RN( fma(p.y, M[2][1], RN(p.x*M[2][0])) + fma(p.z, M[2][2], M[2][3]) )  transform_points (M row-major, numpy)
//     w'  = same with row 3
// and for the filter:  w_non0 += w (add), sum = fma(float(d), w, sum), taps in (dy, dx) raster order, IEEE divide, floor.
#include "rdf_common.cuh"
#include "rdf_fingertip.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define RF_TILE_W 32
#define RF_TILE_H 16          // output rows per CTA; RF_ROWS_PER_THREAD of them per thread
#define RF_THREADS_Y 8
#define RF_ROWS_PER_THREAD (RF_TILE_H / RF_THREADS_Y)
#define RF_MAX_K 41            // reference: PointsOps.MAX_FILTER_SIZE, src/cuda/points_ops.py:34
#define RF_MAX_HANDS 4

struct rf_condition_params {
    const uint16_t* in;
    uint16_t* out;
    uint16_t* mm;              // nullable
    const float* plane;        // device float[16], row-major
    const float* gauss;        // device float[k*k] or nullptr
    int W, H, k, level;
    float ppx, ppy, focal, thresh;
};

// depth sample after deproject -> plane transform -> plane clip -> remove-missing
__device__ __forceinline__ unsigned rf_clip(unsigned d, int x, int y, const rf_condition_params& p, float m20, float m21, float m22,
                                            float m23, float m30, float m31, float m32, float m33) {
    if (d == 0u) return 0u;   // deproject leaves the (stale) point alone; the depth sample is 0 whatever the point holds
    const float dz = (float)d;
    const float px = __fdiv_rn(__fmul_rn(dz, __fsub_rn((float)x, p.ppx)), p.focal);
    const float py = __fdiv_rn(__fmul_rn(dz, __fsub_rn((float)y, p.ppy)), p.focal);
    const float zt = __fadd_rn(__fmaf_rn(py, m21, __fmul_rn(px, m20)), __fmaf_rn(dz, m22, m23));
    const float wt = __fadd_rn(__fmaf_rn(py, m31, __fmul_rn(px, m30)), __fmaf_rn(dz, m32, m33));
    // filter_points_by_plane only looks at points whose w is still exactly 1 (calibrated_plane.cu:41); remove_missing zeroes the
    // depth where w == 0 (points_ops.cu:142)
    const bool missing = (wt == 1.f) ? (zt > -p.thresh) : (wt == 0.f);
    return missing ? 0u : d;
}

template <int K>   // K > 0: compile-time window; K == 0: run-time p.k; K == -1: no filter
__global__ void __launch_bounds__(RF_TILE_W * RF_THREADS_Y) rdf_condition_kernel(const __grid_constant__ rf_condition_params p) {
    // may have been scheduled early behind the kernel that produces the frame (rdf_upload_frame)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    constexpr int KMAX = K > 0 ? K : (K == 0 ? RF_MAX_K : 1);
    constexpr int RMAX = KMAX / 2;
    __shared__ int tile[(RF_TILE_H + 2 * RMAX) * (RF_TILE_W + 2 * RMAX)];     // -1 = outside the image
    __shared__ float wk[KMAX * KMAX];
    const int k = K > 0 ? K : (K == 0 ? p.k : 1);
    const int R = k / 2;
    const int tw = RF_TILE_W + 2 * R, th = RF_TILE_H + 2 * R;
    const int x0 = blockIdx.x * RF_TILE_W, y0 = blockIdx.y * RF_TILE_H;
    const int tid = threadIdx.y * RF_TILE_W + threadIdx.x;
    const float m20 = __ldg(p.plane + 8), m21 = __ldg(p.plane + 9), m22 = __ldg(p.plane + 10), m23 = __ldg(p.plane + 11);
    const float m30 = __ldg(p.plane + 12), m31 = __ldg(p.plane + 13), m32 = __ldg(p.plane + 14), m33 = __ldg(p.plane + 15);
    if (K >= 0)
        for (int i = tid; i < k * k; i += RF_TILE_W * RF_THREADS_Y) wk[i] = __ldg(p.gauss + i);
    bool any = false;
    if (K != 0) {
        // compile-time window: the tile's loads are issued together (a plain loop pays one memory round trip per trip: the clip
        // arithmetic of a sample depends on its load, and the next load is not hoisted above it)
        constexpr int NTH = RF_TILE_W * RF_THREADS_Y;
        constexpr int NIT = ((RF_TILE_W + 2 * RMAX) * (RF_TILE_H + 2 * RMAX) + NTH - 1) / NTH;
        int dv[NIT];
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int i = tid + it * NTH;
            const int cy = i / tw, cx = i - cy * tw;
            const int gx = x0 - R + cx, gy = y0 - R + cy;
            dv[it] = -1;
            if (i < tw * th && gx >= 0 && gx < p.W && gy >= 0 && gy < p.H) dv[it] = (int)__ldg(p.in + (size_t)gy * p.W + gx);
        }
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int i = tid + it * NTH;
            if (i >= tw * th) break;
            const int cy = i / tw, cx = i - cy * tw;
            int v = dv[it];
            if (v > 0) v = (int)rf_clip((unsigned)v, x0 - R + cx, y0 - R + cy, p, m20, m21, m22, m23, m30, m31, m32, m33);
            tile[i] = v;
            any = any || v > 0;
        }
    } else {
        for (int i = tid; i < tw * th; i += RF_TILE_W * RF_THREADS_Y) {
            const int cy = i / tw, cx = i - cy * tw;
            const int gx = x0 - R + cx, gy = y0 - R + cy;
            int v = -1;
            if (gx >= 0 && gx < p.W && gy >= 0 && gy < p.H)
                v = (int)rf_clip(__ldg(p.in + (size_t)gy * p.W + gx), gx, gy, p, m20, m21, m22, m23, m30, m31, m32, m33);
            tile[i] = v;
            any = any || v > 0;
        }
    }
    // most of a live frame is table (clipped to 0): a tile whose whole neighbourhood is 0 filters to 0 whatever the weights
    // (w_non0 = sum = 0: either w_0 > 0 selects 0, or 0/0 = NaN floors to 0)
    const bool all_zero = __syncthreads_or(any) == 0;
    const int x = x0 + threadIdx.x;
    if (x >= p.W) return;
#pragma unroll
    for (int rr = 0; rr < RF_ROWS_PER_THREAD; rr++) {
        const int ty = threadIdx.y + rr * RF_THREADS_Y, y = y0 + ty;
        if (y >= p.H) break;
        unsigned v;
        if (K < 0) {
            v = (unsigned)tile[ty * tw + threadIdx.x];
        } else if (all_zero) {
            v = 0u;
        } else {
            // branch-free taps: a tap outside the image contributes weight 0 to every sum, a zero sample adds +0 to the non-zero
            // sums and a non-zero sample +0 to the zero sum - x + 0 == x exactly, so the sums equal the reference's skipping loop
            float w0 = 0.f, wn = 0.f, s = 0.f;
#pragma unroll
            for (int dy = 0; dy < k; dy++) {
#pragma unroll
                for (int dx = 0; dx < k; dx++) {
                    const int d = tile[(ty + dy) * tw + threadIdx.x + dx];
                    const float w = wk[dy * k + dx];
                    const float wz = d == 0 ? w : 0.f, wp = d > 0 ? w : 0.f;
                    w0 = __fadd_rn(w0, wz);
                    wn = __fadd_rn(wn, wp);
                    s = __fmaf_rn((float)max(d, 0), wp, s);
                }
            }
            v = w0 > wn ? 0u : (__float2uint_rd(__fdiv_rn(s, wn)) & 0xffffu);
        }
        p.out[(size_t)y * p.W + x] = (uint16_t)v;
        if (p.mm) {
            const int f = 1 << p.level;
            if ((x & (f - 1)) == 0 && (y & (f - 1)) == 0) {
                const int xo = x >> p.level, yo = y >> p.level, wo = p.W >> p.level, ho = p.H >> p.level;
                if (xo < wo && yo < ho) p.mm[(size_t)yo * wo + xo] = (uint16_t)v;
            }
        }
    }
}

extern "C" int rdf_condition_depth(const uint16_t* depth_in_dev, int dim_x, int dim_y, float ppx, float ppy, float focal,
                                   const float* plane_dev, float plane_z_threshold, const float* gauss_dev, int k_size,
                                   int mipmap_level, uint16_t* depth_out_dev, uint16_t* depth_mm_dev, void* stream) {
    RDF_REQUIRE(depth_in_dev && plane_dev && depth_out_dev, "rdf_condition_depth: NULL argument");
    RDF_REQUIRE(depth_in_dev != depth_out_dev, "rdf_condition_depth: the filter reads a neighbourhood, output must not alias input");
    RDF_REQUIRE(dim_x > 0 && dim_y > 0, "rdf_condition_depth: bad shape %dx%d", dim_x, dim_y);
    RDF_REQUIRE(mipmap_level >= 0 && mipmap_level <= 15, "rdf_condition_depth: mipmap_level=%d outside 0..15", mipmap_level);
    if (gauss_dev) RDF_REQUIRE(k_size >= 1 && k_size <= RF_MAX_K && (k_size & 1), "rdf_condition_depth: k_size=%d must be odd, 1..%d", k_size, RF_MAX_K);
    rf_condition_params p;
    p.in = depth_in_dev; p.out = depth_out_dev; p.mm = depth_mm_dev; p.plane = plane_dev; p.gauss = gauss_dev;
    p.W = dim_x; p.H = dim_y; p.k = gauss_dev ? k_size : 1; p.level = mipmap_level;
    p.ppx = ppx; p.ppy = ppy; p.focal = focal; p.thresh = plane_z_threshold;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((dim_x + RF_TILE_W - 1) / RF_TILE_W, (dim_y + RF_TILE_H - 1) / RF_TILE_H, 1);
    cfg.blockDim = dim3(RF_TILE_W, RF_THREADS_Y, 1);
    cfg.stream = rdf_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = RDF_GETENV_ONCE("RDF_NO_PDL") ? 0 : 1;
    if (!gauss_dev) RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_condition_kernel<-1>, p));
    else if (k_size == 5) RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_condition_kernel<5>, p));
    else if (k_size == 3) RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_condition_kernel<3>, p));
    else RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_condition_kernel<0>, p));
    return RDF_OK;
}

// ---- groups: grow by one pixel (points_ops.cu:407-438) -------------------------------------------------------------
// DIRS order of the reference: (x-1,y), (x+1,y), (x,y-1), (x,y+1); outside the image reads 0.
__device__ __forceinline__ unsigned rf_grown(const uint16_t* __restrict__ g, int w, int h, int x, int y) {
    // all five cells are requested at once (a chain of "if zero, look at the next neighbour" pays one load latency per step),
    // then the reference's order decides: centre, left, right, up, down
    const unsigned c = __ldg(g + y * w + x);
    const unsigned l = x > 0 ? (unsigned)__ldg(g + y * w + x - 1) : 0u;
    const unsigned r = x + 1 < w ? (unsigned)__ldg(g + y * w + x + 1) : 0u;
    const unsigned u = y > 0 ? (unsigned)__ldg(g + (y - 1) * w + x) : 0u;
    const unsigned d = y + 1 < h ? (unsigned)__ldg(g + (y + 1) * w + x) : 0u;
    return c ? c : l ? l : r ? r : u ? u : d;
}

__global__ void __launch_bounds__(256) rdf_grow_groups_kernel(const uint16_t* __restrict__ g_in, int w, int h, uint16_t* __restrict__ g_out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= w * h) return;
    const int y = i / w, x = i - y * w;
    g_out[i] = (uint16_t)rf_grown(g_in, w, h, x, y);
}

extern "C" int rdf_grow_groups(const uint16_t* groups_in_dev, int dim_x, int dim_y, uint16_t* groups_out_dev, void* stream) {
    RDF_REQUIRE(groups_in_dev && groups_out_dev && groups_in_dev != groups_out_dev, "rdf_grow_groups: NULL or aliased argument");
    RDF_REQUIRE(dim_x > 0 && dim_y > 0 && (int64_t)dim_x * dim_y < ((int64_t)1 << 31), "rdf_grow_groups: bad shape");
    rdf_grow_groups_kernel<<<(dim_x * dim_y + 255) / 256, 256, 0, rdf_stream(stream)>>>(groups_in_dev, dim_x, dim_y, groups_out_dev);
    RDF_LAUNCH_CHECK("rdf_grow_groups_kernel");
    return RDF_OK;
}

// ---- per-hand stencil (+ flip, + 0 -> 65535) -------------------------------------------------------------------------
struct rf_stencil_params {
    const uint16_t* depth;
    const uint16_t* groups;
    uint16_t* out;             // [num_hands, H, W]
    int W, H, level, grow, num_hands;
    int group[RF_MAX_HANDS];
    int flip[RF_MAX_HANDS];
};

template <bool VEC8>   // VEC8: W % 8 == 0, every thread moves 8 consecutive pixels with 128-bit accesses
__global__ void __launch_bounds__(256) rdf_stencil_hands_kernel(const __grid_constant__ rf_stencil_params p) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    constexpr int PX = VEC8 ? 8 : 1;
    const int x = (blockIdx.x * 64 + (threadIdx.x & 63)) * PX;
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= p.W || y >= p.H) return;
    const int gw = p.W >> p.level, gh = p.H >> p.level;
    const int gy = y >> p.level;
    unsigned d[PX], g[PX];
    if (VEC8) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.depth + (size_t)y * p.W + x));
        d[0] = v.x & 0xffffu; d[1] = v.x >> 16; d[2] = v.y & 0xffffu; d[3] = v.y >> 16;
        d[4] = v.z & 0xffffu; d[5] = v.z >> 16; d[6] = v.w & 0xffffu; d[7] = v.w >> 16;
    } else {
        d[0] = __ldg(p.depth + (size_t)y * p.W + x);
    }
    // stencil_depth_image_by_group reads the group image through Array2d: outside -> 0, which matches no hand
#pragma unroll
    for (int j = 0; j < PX; j++) {
        const int gx = (x + j) >> p.level;
        if (j > 0 && gx == ((x + j - 1) >> p.level)) {
            g[j] = g[j - 1];
            continue;
        }
        g[j] = 0u;
        if (gx < gw && gy < gh) g[j] = p.grow ? rf_grown(p.groups, gw, gh, gx, gy) : (unsigned)__ldg(p.groups + gy * gw + gx);
    }
#pragma unroll
    for (int hnd = 0; hnd < RF_MAX_HANDS; hnd++) {
        if (hnd >= p.num_hands) break;
        unsigned v[PX];
#pragma unroll
        for (int j = 0; j < PX; j++) {
            v[j] = ((int)g[j] == p.group[hnd]) ? d[j] : 0u;
            if (v[j] == 0u) v[j] = RDF_NO_PIXEL;                         // convert_0s_to_maxuint
        }
        uint16_t* row = p.out + ((size_t)hnd * p.H + y) * p.W;
        if (VEC8) {
            uint4 o;
            if (p.flip[hnd]) {                                           // flip_x: pixel x + j lands at W - 1 - x - j
                o = make_uint4(v[7] | (v[6] << 16), v[5] | (v[4] << 16), v[3] | (v[2] << 16), v[1] | (v[0] << 16));
                *reinterpret_cast<uint4*>(row + (p.W - 8 - x)) = o;
            } else {
                o = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
                *reinterpret_cast<uint4*>(row + x) = o;
            }
        } else {
            row[p.flip[hnd] ? p.W - 1 - x : x] = (uint16_t)v[0];
        }
    }
}

extern "C" int rdf_stencil_hands(const uint16_t* depth_dev, int dim_x, int dim_y, const uint16_t* groups_dev, int mipmap_level,
                                 int grow, int num_hands, const int* group_ids, const int* flip_x, uint16_t* out_dev, void* stream) {
    RDF_REQUIRE(depth_dev && groups_dev && group_ids && flip_x && out_dev, "rdf_stencil_hands: NULL argument");
    RDF_REQUIRE(dim_x > 0 && dim_y > 0 && mipmap_level >= 0 && mipmap_level <= 15, "rdf_stencil_hands: bad shape");
    RDF_REQUIRE(num_hands >= 1 && num_hands <= RF_MAX_HANDS, "rdf_stencil_hands: num_hands=%d outside 1..%d", num_hands, RF_MAX_HANDS);
    rf_stencil_params p;
    memset(&p, 0, sizeof(p));
    p.depth = depth_dev; p.groups = groups_dev; p.out = out_dev;
    p.W = dim_x; p.H = dim_y; p.level = mipmap_level; p.grow = grow ? 1 : 0; p.num_hands = num_hands;
    for (int i = 0; i < num_hands; i++) { p.group[i] = group_ids[i]; p.flip[i] = flip_x[i] ? 1 : 0; }
    cudaLaunchConfig_t cfg = {};
    const bool vec8 = (dim_x & 7) == 0 && ((uintptr_t)depth_dev & 15u) == 0 && ((uintptr_t)out_dev & 15u) == 0;
    cfg.gridDim = dim3(((vec8 ? dim_x / 8 : dim_x) + 63) / 64, (dim_y + 3) / 4, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = rdf_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = RDF_GETENV_ONCE("RDF_NO_PDL") ? 0 : 1;
    if (vec8) RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_stencil_hands_kernel<true>, p));
    else RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_stencil_hands_kernel<false>, p));
    return RDF_OK;
}

// ---- flip_x (points_ops.cu:466-483) ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rdf_flip_x_kernel(const uint16_t* __restrict__ in, int w, int h, uint16_t* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= w * h) return;
    const int y = i / w, x = i - y * w;
    out[y * w + (w - 1 - x)] = __ldg(in + i);
}

extern "C" int rdf_flip_x(const uint16_t* in_dev, int dim_x, int dim_y, uint16_t* out_dev, void* stream) {
    RDF_REQUIRE(in_dev && out_dev && in_dev != out_dev, "rdf_flip_x: NULL or aliased argument");
    RDF_REQUIRE(dim_x > 0 && dim_y > 0 && (int64_t)dim_x * dim_y < ((int64_t)1 << 31), "rdf_flip_x: bad shape");
    rdf_flip_x_kernel<<<(dim_x * dim_y + 255) / 256, 256, 0, rdf_stream(stream)>>>(in_dev, dim_x, dim_y, out_dev);
    RDF_LAUNCH_CHECK("rdf_flip_x_kernel");
    return RDF_OK;
}

// ---- display helpers (points_ops.cu:258-325) ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rdf_labels_to_rgba_kernel(const uint16_t* __restrict__ labels, int n, const uchar4* __restrict__ colors,
                                                                 int num_colors, uchar4* __restrict__ rgba) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const unsigned l = __ldg(labels + i);
    if (l == 0u || l == RDF_NO_PIXEL) return;                  // pixel left untouched
    if ((int)l - 1 >= num_colors) return;                      // the reference would dereference a null pointer here
    rgba[i] = __ldg(colors + (l - 1));
}

extern "C" int rdf_labels_to_rgba(const uint16_t* labels_dev, int dim_x, int dim_y, const uint8_t* colors_dev, int num_colors,
                                  uint8_t* rgba_dev, void* stream) {
    RDF_REQUIRE(labels_dev && colors_dev && rgba_dev, "rdf_labels_to_rgba: NULL argument");
    RDF_REQUIRE(dim_x > 0 && dim_y > 0 && num_colors >= 0 && (int64_t)dim_x * dim_y < ((int64_t)1 << 31), "rdf_labels_to_rgba: bad shape");
    RDF_REQUIRE(((uintptr_t)colors_dev & 3u) == 0 && ((uintptr_t)rgba_dev & 3u) == 0, "rdf_labels_to_rgba: colour buffers must be 4-byte aligned");
    const int n = dim_x * dim_y;
    rdf_labels_to_rgba_kernel<<<(n + 255) / 256, 256, 0, rdf_stream(stream)>>>(labels_dev, n, reinterpret_cast<const uchar4*>(colors_dev),
                                                                              num_colors, reinterpret_cast<uchar4*>(rgba_dev));
    RDF_LAUNCH_CHECK("rdf_labels_to_rgba_kernel");
    return RDF_OK;
}

__global__ void __launch_bounds__(256) rdf_depth_to_rgba_kernel(const uint16_t* __restrict__ depth, int n, unsigned d_min, unsigned d_max,
                                                                uchar4* __restrict__ rgba) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const unsigned d = __ldg(depth + i);
    uchar4 c = make_uchar4(0, 0, 0, 255);
    if (d == 0u) {
        c.x = 195; c.y = 157; c.z = 152;
    } else if (d == RDF_NO_PIXEL) {
        c.x = 157; c.y = 195; c.z = 152;
    } else if (d < d_min || d > d_max) {
        c.x = 157; c.y = 152; c.z = 195;
    } else {
        // ((1.0f * d - d_min) * 255.f) / (d_max - d_min), then (uint8)floor(256.f - n_f)
        const float nf = __fdiv_rn(__fmul_rn(__fsub_rn((float)d, (float)d_min), 255.f), (float)((int)d_max - (int)d_min));
        const unsigned char g = (unsigned char)__float2uint_rd(__fsub_rn(256.f, nf));
        c.x = g; c.y = g; c.z = g;
    }
    rgba[i] = c;
}

extern "C" int rdf_depth_to_rgba(const uint16_t* depth_dev, int dim_x, int dim_y, int d_min, int d_max, uint8_t* rgba_dev, void* stream) {
    RDF_REQUIRE(depth_dev && rgba_dev, "rdf_depth_to_rgba: NULL argument");
    RDF_REQUIRE(dim_x > 0 && dim_y > 0 && (int64_t)dim_x * dim_y < ((int64_t)1 << 31), "rdf_depth_to_rgba: bad shape");
    RDF_REQUIRE(d_min >= 0 && d_min <= 65535 && d_max >= 0 && d_max <= 65535, "rdf_depth_to_rgba: d_min / d_max outside uint16");
    RDF_REQUIRE(((uintptr_t)rgba_dev & 3u) == 0, "rdf_depth_to_rgba: rgba buffer must be 4-byte aligned");
    const int n = dim_x * dim_y;
    rdf_depth_to_rgba_kernel<<<(n + 255) / 256, 256, 0, rdf_stream(stream)>>>(depth_dev, n, (unsigned)d_min, (unsigned)d_max,
                                                                             reinterpret_cast<uchar4*>(rgba_dev));
    RDF_LAUNCH_CHECK("rdf_depth_to_rgba_kernel");
    return RDF_OK;
}

// ---- fingertip read-out (src/3d_bz.py:503-522) ----------------------------------------------------------------------
// per fingertip i with label index f = idx[i]:  (px, py) = int32(means[f-1]) * labels_reduce  (NaN -> INT_MIN, as numpy's astype on
// x86 gives); outside the frame -> "reset" (NaN here);  z = raw_depth[py, px];  pt = rs2_deproject_pixel_to_point (no distortion:
// fp32 x = (px-ppx)/fx, y = (py-ppy)/fy, point = (z*x, z*y, z));  -(plane[2,:] . (pt, 1)) accumulated in fp64 like numpy's matmul
// of a float32 matrix with a Python-float vector.
struct rf_fingertip_params {
    const double* means;
    int num_labels;
    rf_fingertip_spec ft;
};

__global__ void __launch_bounds__(RF_MAX_FINGERTIPS) rdf_fingertip_z_kernel(const __grid_constant__ rf_fingertip_params p) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int i = threadIdx.x;
    const double* __restrict__ means = p.means + (size_t)blockIdx.x * 2 * p.num_labels;      // one block per image (hand)
    if (p.ft.means_copy)
        for (int j = i; j < 2 * p.num_labels; j += RF_MAX_FINGERTIPS) p.ft.means_copy[(size_t)blockIdx.x * 2 * p.num_labels + j] = means[j];
    if (i >= p.ft.n) return;
    double out = __longlong_as_double(0x7ff8000000000000ll);
    const int f = p.ft.idx[i];
    if (f >= 1 && f <= p.num_labels) out = rf_fingertip_eval(p.ft, means[2 * (f - 1) + 0], means[2 * (f - 1) + 1]);
    p.ft.z_out[(size_t)blockIdx.x * p.ft.n + i] = out;
}

int rdf_fingertip_fill_spec(rf_fingertip_spec* ft, const char* who, const int* fingertip_labels, int num_fingertips, int labels_reduce,
                            const uint16_t* raw_depth_dev, int dim_x, int dim_y, float ppx, float ppy, float fx, float fy,
                            const float* plane_dev, double* z_out, double* means_copy_out) {
    RDF_REQUIRE(fingertip_labels && raw_depth_dev && plane_dev && z_out, "%s: NULL argument", who);
    RDF_REQUIRE(num_fingertips >= 1 && num_fingertips <= RF_MAX_FINGERTIPS, "%s: num_fingertips=%d outside 1..%d", who, num_fingertips,
                RF_MAX_FINGERTIPS);
    RDF_REQUIRE(labels_reduce >= 1 && dim_x > 0 && dim_y > 0, "%s: bad shape", who);
    memset(ft, 0, sizeof(*ft));
    ft->raw = raw_depth_dev; ft->plane = plane_dev; ft->z_out = z_out; ft->means_copy = means_copy_out;
    ft->n = num_fingertips; ft->r = labels_reduce; ft->W = dim_x; ft->H = dim_y;
    ft->ppx = ppx; ft->ppy = ppy; ft->fx = fx; ft->fy = fy;
    for (int i = 0; i < num_fingertips; i++) ft->idx[i] = fingertip_labels[i];
    return RDF_OK;
}

extern "C" int rdf_fingertip_z(const double* means_dev, int num_images, int num_labels, const int* fingertip_labels, int num_fingertips,
                               int labels_reduce, const uint16_t* raw_depth_dev, int dim_x, int dim_y, float ppx, float ppy, float fx,
                               float fy, const float* plane_dev, double* z_out, double* means_copy_out, void* stream) {
    RDF_REQUIRE(means_dev && num_images >= 1 && num_labels >= 1, "rdf_fingertip_z: bad means argument");
    rf_fingertip_params p;
    p.means = means_dev;
    p.num_labels = num_labels;
    const int rc = rdf_fingertip_fill_spec(&p.ft, "rdf_fingertip_z", fingertip_labels, num_fingertips, labels_reduce, raw_depth_dev, dim_x,
                                           dim_y, ppx, ppy, fx, fy, plane_dev, z_out, means_copy_out);
    if (rc != RDF_OK) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_images, 1, 1);
    cfg.blockDim = dim3(RF_MAX_FINGERTIPS, 1, 1);
    cfg.stream = rdf_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = RDF_GETENV_ONCE("RDF_NO_PDL") ? 0 : 1;
    RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_fingertip_z_kernel, p));
    return RDF_OK;
}
