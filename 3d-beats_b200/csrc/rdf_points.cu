// The reference's pre-forest point kernels ONE BY ONE, for callers that drive them individually through pycuda-style calls:
// run_live.py:86-121 and run_live_layered.py:87-122 (deproject -> [plane calibration] -> transform -> plane filter ->
// setup_depth_image_for_forest) and the product loop src/3d_bz.py:159-259,390-420.  The fused product path is rdf_frame.cu
// (rdf_condition_depth / rdf_stencil_hands); these entry points exist so that the scripts' own call sequences keep working
// against the drop-in (compat/rdf_dropin.py) with a user-visible float4 point image, as the reference has it.
//
//   rdf_deproject_points        deproject_points                              src/cuda/points_ops.cu:5-36
//   rdf_transform_points        transform_points                              src/cuda/points_ops.cu:66-75
//   rdf_filter_points_by_plane  filter_points_by_plane                        src/cuda/calibrated_plane.cu:30-45
//   rdf_remove_missing_points   remove_missing_3d_points_from_depth_image     src/cuda/points_ops.cu:131-146
//   rdf_setup_depth_for_forest  setup_depth_image_for_forest                  src/cuda/points_ops.cu:149-165
//   rdf_zeros_to_no_pixel       convert_0s_to_maxuint                         src/cuda/points_ops.cu:118-127
//   rdf_shrink_image            shrink_image                                  src/cuda/points_ops.cu:375-404
//   rdf_stencil_by_group        stencil_depth_image_by_group                  src/cuda/points_ops.cu:441-463
//   rdf_scatter_groups          write_pixel_groups_to_stencil_image           src/cuda/points_ops.cu:486-503
//
// All of them are pure streaming kernels (16 B per point or 2 B per pixel, no reuse): grid-stride loops sized from the SM
// count, 128-bit point accesses.  fp32 operation order = the reference's compiled kernels (see rdf_frame.cu's header: GLM's
// mat4 * vec4 is (m0*x + m1*y) + (m2*z + m3*w), contracted by nvcc to fma(y, m1, x*m0) + fma(z, m2, m3) for w == 1).
#include "rdf_common.cuh"

static inline unsigned rp_blocks(int64_t n, int threads, int per_sm) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)rdf_sm_count() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// ---- deproject ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rdf_deproject_points_kernel(const uint16_t* __restrict__ depth, int64_t total, int W, int H,
                                                                   float ppx, float ppy, float f, float4* __restrict__ pts) {
    const int per = W * H;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned d = __ldg(depth + i);
        if (d == 0u) continue;                                        // the point keeps whatever it held (points_ops.cu:24)
        const int rem = (int)(i % per);
        const int y = rem / W, x = rem - y * W;
        const float dz = (float)d;
        float4 p;
        p.x = __fdiv_rn(__fmul_rn(dz, __fsub_rn((float)x, ppx)), f);
        p.y = __fdiv_rn(__fmul_rn(dz, __fsub_rn((float)y, ppy)), f);
        p.z = dz;
        p.w = 1.f;
        pts[i] = p;
    }
}

extern "C" int rdf_deproject_points(const uint16_t* depth_dev, int num_images, int dim_x, int dim_y, float ppx, float ppy, float focal,
                                    float* pts_dev, void* stream) {
    RDF_REQUIRE(depth_dev && pts_dev, "rdf_deproject_points: NULL argument");
    RDF_REQUIRE(num_images >= 0 && dim_x > 0 && dim_y > 0 && (int64_t)dim_x * dim_y < ((int64_t)1 << 31), "rdf_deproject_points: bad shape");
    RDF_REQUIRE((reinterpret_cast<uintptr_t>(pts_dev) & 15u) == 0, "rdf_deproject_points: pts_dev must be 16-byte aligned");
    const int64_t total = (int64_t)num_images * dim_x * dim_y;
    if (total == 0) return RDF_OK;
    rdf_deproject_points_kernel<<<rp_blocks(total, 256, 16), 256, 0, rdf_stream(stream)>>>(depth_dev, total, dim_x, dim_y, ppx, ppy, focal,
                                                                                          reinterpret_cast<float4*>(pts_dev));
    RDF_LAUNCH_CHECK("rdf_deproject_points_kernel");
    return RDF_OK;
}

// ---- transform ------------------------------------------------------------------------------------------------------------
struct rp_mat4 {
    float m[16];                                                      // row-major numpy float32[4][4]
};

__global__ void __launch_bounds__(256) rdf_transform_points_kernel(int64_t n, float4* __restrict__ pts, const rp_mat4 t) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        if (p.w != 1.f) continue;                                     // points_ops.cu:72
        float r[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float* m = t.m + 4 * c;                             // row c of M: (transpose(t) * p)[c] = sum_k M[c][k] p[k]
            const float a0 = __fmaf_rn(p.y, m[1], __fmul_rn(p.x, m[0]));
            const float a1 = __fmaf_rn(p.z, m[2], m[3]);              // m[3] * w with w == 1
            r[c] = __fadd_rn(a0, a1);
        }
        pts[i] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

extern "C" int rdf_transform_points(int num_pts, float* pts_dev, const float* mat_host, void* stream) {
    RDF_REQUIRE(pts_dev && mat_host && num_pts >= 0, "rdf_transform_points: bad argument");
    RDF_REQUIRE((reinterpret_cast<uintptr_t>(pts_dev) & 15u) == 0, "rdf_transform_points: pts_dev must be 16-byte aligned");
    if (num_pts == 0) return RDF_OK;
    rp_mat4 t;
    for (int i = 0; i < 16; i++) t.m[i] = mat_host[i];                // passed by value like the reference's glm::mat4 argument
    rdf_transform_points_kernel<<<rp_blocks(num_pts, 256, 16), 256, 0, rdf_stream(stream)>>>(num_pts, reinterpret_cast<float4*>(pts_dev), t);
    RDF_LAUNCH_CHECK("rdf_transform_points_kernel");
    return RDF_OK;
}

// ---- plane filter ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rdf_filter_points_by_plane_kernel(int64_t n, float thresh, float4* __restrict__ pts) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        if (p.w != 1.f) continue;                                     // calibrated_plane.cu:41
        if (p.z > -thresh) pts[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

extern "C" int rdf_filter_points_by_plane(int num_pts, float plane_z_threshold, float* pts_dev, void* stream) {
    RDF_REQUIRE(pts_dev && num_pts >= 0, "rdf_filter_points_by_plane: bad argument");
    RDF_REQUIRE((reinterpret_cast<uintptr_t>(pts_dev) & 15u) == 0, "rdf_filter_points_by_plane: pts_dev must be 16-byte aligned");
    if (num_pts == 0) return RDF_OK;
    rdf_filter_points_by_plane_kernel<<<rp_blocks(num_pts, 256, 16), 256, 0, rdf_stream(stream)>>>(num_pts, plane_z_threshold,
                                                                                                  reinterpret_cast<float4*>(pts_dev));
    RDF_LAUNCH_CHECK("rdf_filter_points_by_plane_kernel");
    return RDF_OK;
}

// ---- depth image fix-ups --------------------------------------------------------------------------------------------------
// MODE 0: remove_missing (w == 0 -> depth 0); MODE 1: setup_for_forest (depth 0 or w == 0 -> 65535); only the w lane is read
template <int MODE>
__global__ void __launch_bounds__(256) rdf_points_to_depth_kernel(int64_t n, const float4* __restrict__ pts, uint16_t* __restrict__ depth) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float w = __ldg(reinterpret_cast<const float*>(pts + i) + 3);
        if (MODE == 0) {
            if (w == 0.f) depth[i] = 0;                               // points_ops.cu:142-144
        } else {
            if (depth[i] == 0 || w == 0.f) depth[i] = (uint16_t)RDF_NO_PIXEL;   // points_ops.cu:161-163
        }
    }
}

extern "C" int rdf_remove_missing_points(int num_pts, const float* pts_dev, uint16_t* depth_dev, void* stream) {
    RDF_REQUIRE(pts_dev && depth_dev && num_pts >= 0, "rdf_remove_missing_points: bad argument");
    if (num_pts == 0) return RDF_OK;
    rdf_points_to_depth_kernel<0><<<rp_blocks(num_pts, 256, 16), 256, 0, rdf_stream(stream)>>>(num_pts, reinterpret_cast<const float4*>(pts_dev),
                                                                                              depth_dev);
    RDF_LAUNCH_CHECK("rdf_points_to_depth_kernel<0>");
    return RDF_OK;
}

extern "C" int rdf_setup_depth_for_forest(int num_pts, const float* pts_dev, uint16_t* depth_dev, void* stream) {
    RDF_REQUIRE(pts_dev && depth_dev && num_pts >= 0, "rdf_setup_depth_for_forest: bad argument");
    if (num_pts == 0) return RDF_OK;
    rdf_points_to_depth_kernel<1><<<rp_blocks(num_pts, 256, 16), 256, 0, rdf_stream(stream)>>>(num_pts, reinterpret_cast<const float4*>(pts_dev),
                                                                                              depth_dev);
    RDF_LAUNCH_CHECK("rdf_points_to_depth_kernel<1>");
    return RDF_OK;
}

__global__ void __launch_bounds__(256) rdf_zeros_to_no_pixel_kernel(int64_t n, uint16_t* __restrict__ depth) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (depth[i] == 0) depth[i] = (uint16_t)RDF_NO_PIXEL;        // points_ops.cu:124-126
}

extern "C" int rdf_zeros_to_no_pixel(int num_pixels, uint16_t* depth_dev, void* stream) {
    RDF_REQUIRE(depth_dev && num_pixels >= 0, "rdf_zeros_to_no_pixel: bad argument");
    if (num_pixels == 0) return RDF_OK;
    rdf_zeros_to_no_pixel_kernel<<<rp_blocks(num_pixels, 256, 16), 256, 0, rdf_stream(stream)>>>(num_pixels, depth_dev);
    RDF_LAUNCH_CHECK("rdf_zeros_to_no_pixel_kernel");
    return RDF_OK;
}

// ---- 1/2^level image, per-group stencil, coordinate scatter ---------------------------------------------------------------
__global__ void __launch_bounds__(256) rdf_shrink_image_kernel(const uint16_t* __restrict__ in, int W, int H, int level,
                                                               uint16_t* __restrict__ out) {
    const int w = W >> level, h = H >> level;                         // IMG_DIM_IN / f (points_ops.cu:382-386)
    const int n = w * h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / w, x = i - y * w;
        out[i] = __ldg(in + (size_t)(y << level) * W + (x << level)); // x_in < W and y_in < H always hold here
    }
}

extern "C" int rdf_shrink_image(const uint16_t* in_dev, int dim_x, int dim_y, int mipmap_level, uint16_t* out_dev, void* stream) {
    RDF_REQUIRE(in_dev && out_dev && dim_x > 0 && dim_y > 0 && mipmap_level >= 0 && mipmap_level < 16, "rdf_shrink_image: bad argument");
    RDF_REQUIRE((int64_t)dim_x * dim_y < ((int64_t)1 << 31), "rdf_shrink_image: image too large");
    const int n = (dim_x >> mipmap_level) * (dim_y >> mipmap_level);
    if (n == 0) return RDF_OK;
    rdf_shrink_image_kernel<<<rp_blocks(n, 256, 8), 256, 0, rdf_stream(stream)>>>(in_dev, dim_x, dim_y, mipmap_level, out_dev);
    RDF_LAUNCH_CHECK("rdf_shrink_image_kernel");
    return RDF_OK;
}

__global__ void __launch_bounds__(256) rdf_stencil_by_group_kernel(const uint16_t* __restrict__ groups, const uint16_t* __restrict__ depth,
                                                                   int W, int H, int level, int group, uint16_t* __restrict__ out) {
    const int gw = W >> level, gh = H >> level;                       // the group image is (IMG_DIM.y / f) x (IMG_DIM.x / f)
    const int n = W * H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        const int gy = y >> level, gx = x >> level;
        // outside the group image (dims not multiples of 2^level) Array2d::get returns its default 0 (cu_utils.hpp:103-107)
        const int g = (gy < gh && gx < gw) ? (int)__ldg(groups + (size_t)gy * gw + gx) : 0;
        if (g == group) out[i] = __ldg(depth + i);                    // points_ops.cu:458-462
    }
}

extern "C" int rdf_stencil_by_group(const uint16_t* groups_dev, const uint16_t* depth_dev, int dim_x, int dim_y, int mipmap_level,
                                    int group, uint16_t* out_dev, void* stream) {
    RDF_REQUIRE(groups_dev && depth_dev && out_dev && dim_x > 0 && dim_y > 0 && mipmap_level >= 0 && mipmap_level < 16,
                "rdf_stencil_by_group: bad argument");
    RDF_REQUIRE((int64_t)dim_x * dim_y < ((int64_t)1 << 31), "rdf_stencil_by_group: image too large");
    rdf_stencil_by_group_kernel<<<rp_blocks((int64_t)dim_x * dim_y, 256, 16), 256, 0, rdf_stream(stream)>>>(groups_dev, depth_dev, dim_x, dim_y,
                                                                                                           mipmap_level, group, out_dev);
    RDF_LAUNCH_CHECK("rdf_stencil_by_group_kernel");
    return RDF_OK;
}

__global__ void __launch_bounds__(256) rdf_scatter_groups_kernel(const int32_t* __restrict__ coords, int n, uint16_t* __restrict__ stencil,
                                                                 int rows, int cols) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = __ldg(coords + 3 * i), c = __ldg(coords + 3 * i + 1), g = __ldg(coords + 3 * i + 2);
        if ((unsigned)r < (unsigned)rows && (unsigned)c < (unsigned)cols)   // the reference's Array2d::set asserts the same bounds
            stencil[(size_t)r * cols + c] = (uint16_t)g;              // points_ops.cu:497-502
    }
}

extern "C" int rdf_scatter_groups(const int32_t* coords_dev, int num_coords, uint16_t* stencil_dev, int rows, int cols, void* stream) {
    RDF_REQUIRE(coords_dev && stencil_dev && num_coords >= 0 && rows > 0 && cols > 0, "rdf_scatter_groups: bad argument");
    if (num_coords == 0) return RDF_OK;
    rdf_scatter_groups_kernel<<<rp_blocks(num_coords, 256, 8), 256, 0, rdf_stream(stream)>>>(coords_dev, num_coords, stencil_dev, rows, cols);
    RDF_LAUNCH_CHECK("rdf_scatter_groups_kernel");
    return RDF_OK;
}
