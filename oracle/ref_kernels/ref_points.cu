// Test-only harness around the REFERENCE's own pre-/post-forest kernels (src/cuda/points_ops.cu, calibrated_plane.cu).
// TEST INFRASTRUCTURE - never linked into the product.  The sources are #included from where they lie under /root/reference
// (oracle/Makefile target `ref`); GLM, which upstream does not vendor, is satisfied by oracle/ref_kernels/glm_min (see there).
// Launch geometry = the reference host code's, cited per function.
#include <cuda_runtime.h>
#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <points_ops.cu>        // reference: src/cuda/points_ops.cu
#include <calibrated_plane.cu>  // reference: src/cuda/calibrated_plane.cu

#define REF_EXPORT extern "C" __attribute__((visibility("default")))
static int ref_check() { return cudaGetLastError() == cudaSuccess ? 0 : -2; }
static unsigned cdiv(int a, int b) { return (unsigned)((a + b - 1) / b); }   // src/util.py:20-27 make_grid

// src/3d_bz.py:159-170
REF_EXPORT int ref_deproject_points(int dim_x, int dim_y, float ppx, float ppy, float focal, uint16_t* depth, float* pts, void* stream) {
    dim3 block(1, 32, 32), grid(1, cdiv(dim_x, 32), cdiv(dim_y, 32));
    deproject_points<<<grid, block, 0, (cudaStream_t)stream>>>(int4{1, dim_x, dim_y, -1}, float2{ppx, ppy}, focal, depth, (float4*)pts);
    return ref_check();
}
// src/3d_bz.py:180-189; `plane` is the row-major numpy float32[4,4] passed by value as a glm::mat4
REF_EXPORT int ref_transform_points(int num_pts, float* pts, const float* plane_host, void* stream) {
    glm::mat4 t; memcpy(&t, plane_host, 64);
    transform_points<<<cdiv(num_pts, 1024), 1024, 0, (cudaStream_t)stream>>>(num_pts, (glm::vec4*)pts, t);
    return ref_check();
}
// src/3d_bz.py:191-196
REF_EXPORT int ref_filter_points_by_plane(int num_pts, float thresh, float* pts, void* stream) {
    filter_points_by_plane<<<cdiv(num_pts, 1024), 1024, 0, (cudaStream_t)stream>>>(num_pts, thresh, (glm::vec4*)pts);
    return ref_check();
}
// src/3d_bz.py:198-203
REF_EXPORT int ref_remove_missing(int num_pts, float* pts, uint16_t* depth, void* stream) {
    remove_missing_3d_points_from_depth_image<<<cdiv(num_pts, 1024), 1024, 0, (cudaStream_t)stream>>>(num_pts, (glm::vec4*)pts, depth);
    return ref_check();
}
// src/run_live.py:116-121, src/run_live_layered.py:117-122
REF_EXPORT int ref_setup_depth_image_for_forest(int num_pts, float* pts, uint16_t* depth, void* stream) {
    setup_depth_image_for_forest<<<num_pts / 1024 + 1, 1024, 0, (cudaStream_t)stream>>>(num_pts, (glm::vec4*)pts, depth);
    return ref_check();
}
// src/cuda/points_ops.py:62-98
REF_EXPORT int ref_gaussian_depth_filter(int dim_x, int dim_y, int k_size, float* k, uint16_t* d_in, uint16_t* d_out, void* stream) {
    dim3 block(32, 32, 1), grid(cdiv(dim_x, 32), cdiv(dim_y, 32), 1);
    gaussian_depth_filter<<<grid, block, 0, (cudaStream_t)stream>>>(int2{dim_x, dim_y}, k_size, k, d_in, d_out);
    return ref_check();
}
// src/3d_bz.py:213-220
REF_EXPORT int ref_shrink_image(int dim_x, int dim_y, int level, uint16_t* d_in, uint16_t* d_out, void* stream) {
    dim3 block(32, 32, 1), grid(cdiv(dim_x >> level, 32), cdiv(dim_y >> level, 32), 1);
    shrink_image<<<grid, block, 0, (cudaStream_t)stream>>>(int2{dim_x, dim_y}, level, d_in, d_out);
    return ref_check();
}
// src/3d_bz.py:252-259 (dims of the 1/2^level image)
REF_EXPORT int ref_grow_groups(int dim_x, int dim_y, uint16_t* g_in, uint16_t* g_out, void* stream) {
    dim3 block(32, 32, 1), grid(cdiv(dim_x, 32), cdiv(dim_y, 32), 1);
    grow_groups<<<grid, block, 0, (cudaStream_t)stream>>>(int2{dim_x, dim_y}, g_in, g_out);
    return ref_check();
}
// src/3d_bz.py:394-403
REF_EXPORT int ref_stencil_depth_image_by_group(int dim_x, int dim_y, int level, int group, uint16_t* g_in, uint16_t* d_in,
                                                uint16_t* d_out, void* stream) {
    dim3 block(32, 32, 1), grid(cdiv(dim_x, 32), cdiv(dim_y, 32), 1);
    stencil_depth_image_by_group<<<grid, block, 0, (cudaStream_t)stream>>>(int2{dim_x, dim_y}, level, group, g_in, d_in, d_out);
    return ref_check();
}
// src/3d_bz.py:405-411, 439-446
REF_EXPORT int ref_flip_x(int dim_x, int dim_y, uint16_t* in, uint16_t* out, void* stream) {
    dim3 block(32, 32, 1), grid(cdiv(dim_x, 32), cdiv(dim_y, 32), 1);
    flip_x<<<grid, block, 0, (cudaStream_t)stream>>>(int2{dim_x, dim_y}, in, out);
    return ref_check();
}
// src/3d_bz.py:415-419
REF_EXPORT int ref_convert_0s_to_maxuint(int num_pixels, uint16_t* depth, void* stream) {
    convert_0s_to_maxuint<<<cdiv(num_pixels, 1024), 1024, 0, (cudaStream_t)stream>>>(num_pixels, depth);
    return ref_check();
}
// src/3d_bz.py:448-456
REF_EXPORT int ref_make_rgba_from_labels(int dim_x, int dim_y, int num_colors, uint16_t* labels, uint8_t* colors, uint8_t* rgba,
                                         void* stream) {
    dim3 block(32, 32, 1), grid(dim_x / 32 + 1, dim_y / 32 + 1, 1);
    make_rgba_from_labels<<<grid, block, 0, (cudaStream_t)stream>>>(dim_x, dim_y, num_colors, labels, colors, rgba);
    return ref_check();
}
// src/3d_bz.py:266-274
REF_EXPORT int ref_make_depth_rgba(int dim_x, int dim_y, int d_min, int d_max, uint16_t* d, uint8_t* rgba, void* stream) {
    dim3 block(32, 32, 1), grid(cdiv(dim_x, 32), cdiv(dim_y, 32), 1);
    make_depth_rgba<<<grid, block, 0, (cudaStream_t)stream>>>(int2{dim_x, dim_y}, (uint16)d_min, (uint16)d_max, d, rgba);
    return ref_check();
}
