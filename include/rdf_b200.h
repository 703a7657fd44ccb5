/*
 * rdf_b200.h - C ABI of librdf_b200.so: the B200-native (sm_100a) implementation of the per-pixel randomized
 * decision forest hot path of carsonswope/3d-beats.
 *
 * The reference has no C ABI: its operator boundary is pycuda's
 *     module.get_function("<extern C kernel>")(np scalars, GPUArrays, grid=, block=, shared=)
 * (src/cuda/py_nvcc_utils.py:25-37, src/decision_tree.py:269-272,315-330).  Every entry point below replaces one
 * such kernel fetch + launch (cited per function).  All pointers marked `dev` are CUDA device pointers owned by
 * the caller; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *
 * Conventions
 *   - every function returns 0 on success and a negative rdf_status on failure; rdf_last_error() returns a
 *     thread-local, human-readable message for the last failure on the calling thread.
 *   - every launch is asynchronous on `stream`; nothing here allocates or synchronises after handle creation, so
 *     all eval / layered / mean-shift / train calls are CUDA-graph capturable.
 *   - images are C-contiguous: depth uint16[N,H,W]; label maps uint16[N,H/r,W/r] with r = labels_reduce.
 *   - 65535 is the "no pixel" sentinel in depth and label images (src/cuda/cu_utils.hpp:8).
 *   - outputs are NEVER written for skipped pixels (filter mismatch, centre depth 0 or 65535) - callers pre-fill,
 *     exactly as with the reference kernels (src/run_live.py:123, src/test_on_saved_model.py:55).
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails with RDF_ERR_CUDA.
 */
#ifndef RDF_B200_H
#define RDF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDF_B200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define RDF_API __attribute__((visibility("default")))
#else
#define RDF_API
#endif

typedef enum rdf_status {
    RDF_OK = 0,
    RDF_ERR_INVALID = -1,      /* bad argument (null pointer, non-positive size, unsupported shape) */
    RDF_ERR_CUDA = -2,         /* a CUDA runtime call or launch failed; message holds cudaGetErrorString */
    RDF_ERR_UNSUPPORTED = -3   /* valid request outside the compiled limits (see RDF_MAX_*) */
} rdf_status;

#define RDF_MAX_CLASSES 256    /* classes per forest, including class 0 = none */
#define RDF_MAX_LAYERS 8       /* layers in a stacked forest */
#define RDF_MAX_DEPTH 26       /* tree levels */

typedef struct rdf_forest rdf_forest_t; /* opaque: packed device copy of a forest */

RDF_API int rdf_version(void);
RDF_API const char* rdf_last_error(void);

/* ---- forest container ------------------------------------------------------------------------------------
 * Replaces DecisionForest.__init__/.load's `forest_cu` as the thing kernels read (src/decision_tree.py:146-168).
 * canon_dev: float32[T, 2^D-1, 7+2C] canonical layout (node = ux,uy,vx,vy,thresh,l_next,r_next,l_pdf[C],r_pdf[C];
 * src/cuda/tree_eval.cu:47, node addressing src/cuda/cu_utils.hpp:32-39).  The handle owns a packed shadow
 * (32-byte node headers + 16-byte aligned leaf pdf rows); the caller keeps ownership of canon_dev.
 * rdf_forest_update re-packs after the caller mutated canon_dev (e.g. forest_cu.set(...), src/train_model.py:126).
 * Create and update synchronise `stream` once (the pack reports back whether any node needs the exact-divide path, and the handle
 * keeps a host copy of the upper five levels of every tree, at most 8 KB, which rdf_eval_forest passes to its kernel as launch
 * parameters); they are set-up calls, not part of the per-frame path. */
RDF_API int rdf_forest_create(const float* canon_dev, int num_trees, int max_depth, int num_classes, void* stream,
                      rdf_forest_t** out);
RDF_API int rdf_forest_update(rdf_forest_t* forest, const float* canon_dev, void* stream);
RDF_API int rdf_forest_destroy(rdf_forest_t* forest);
RDF_API int rdf_forest_info(const rdf_forest_t* forest, int* num_trees, int* max_depth, int* num_classes,
                    size_t* packed_bytes);

/* ---- evaluation -----------------------------------------------------------------------------------------
 * rdf_eval_forest replaces kernel `evaluate_image_using_forest` (src/cuda/tree_eval.cu:24-137) and its host
 * launch DecisionTreeEvaluator.get_labels_forest (src/decision_tree.py:298-330).
 *   filter_dev   uint16[N,h,w] or NULL; a pixel is evaluated only if filter == filter_class (ignored when
 *                filter_dev is NULL or filter_class == -1; the reference passes a dummy pointer and -1).
 *   labels_dev   uint16[N,h,w], h = H/labels_reduce, w = W/labels_reduce; labels pixel (x,y) samples depth (x*r,y*r).
 *   probs_dev    optional float32[N,h,w,C]: mean over trees of the reached leaf pdfs (sum in tree order / T);
 *                NULL to skip.  Not part of the reference; written only for evaluated pixels.
 *   scale        uv_scale of compute_feature (src/cuda/decision_tree_common.hpp:8-28).
 * Label = first class with the strictly greatest summed pdf > 0, else 0 (src/cuda/tree_eval.cu:7-21); pdf sums are
 * accumulated in tree order 0..T-1. */
RDF_API int rdf_eval_forest(const rdf_forest_t* forest, const uint16_t* depth_dev, int num_images, int dim_x, int dim_y,
                    const uint16_t* filter_dev, int filter_class, uint16_t* labels_dev, float* probs_dev,
                    int labels_reduce, float scale, void* stream);

/* Texture-unit variant of rdf_eval_forest for batches (scale 1, no probability output, forests whose offsets are all inside
 * the fast-divide domain - otherwise RDF_ERR_UNSUPPORTED): the frames are first copied into a layered 2-D array holding
 * 65535 - d (rdf_depth_tex_upload, at most 2048 frames per array), then every depth probe is one integer-coordinate texel fetch
 * with border addressing (outside the image -> 65535, src/cuda/cu_utils.hpp:58-62,79-86).  depth_dev = the same frames in linear
 * memory (centre depth).  Handle creation / destruction allocate; upload and eval are asynchronous on `stream`. */
typedef struct rdf_depth_tex rdf_depth_tex_t;
RDF_API int rdf_depth_tex_create(int max_images, int dim_x, int dim_y, rdf_depth_tex_t** out);
RDF_API int rdf_depth_tex_destroy(rdf_depth_tex_t* tex);
RDF_API int rdf_depth_tex_upload(rdf_depth_tex_t* tex, const uint16_t* depth_dev, int num_images, void* stream);
RDF_API int rdf_eval_forest_tex(const rdf_forest_t* forest, const rdf_depth_tex_t* tex, const uint16_t* depth_dev, int num_images,
                        const uint16_t* filter_dev, int filter_class, uint16_t* labels_dev, int labels_reduce, void* stream);

/* Same contract as rdf_eval_forest, reading the canonical array float32[T,2^D-1,7+2C] directly (no handle, any
 * number of trees): the path for forests with more than 8 trees, which the packed fast path does not cover. */
RDF_API int rdf_eval_forest_canonical(const float* forest_dev, int num_trees, int max_depth, int num_classes,
                              const uint16_t* depth_dev, int num_images, int dim_x, int dim_y,
                              const uint16_t* filter_dev, int filter_class, uint16_t* labels_dev, float* probs_dev,
                              int labels_reduce, float scale, void* stream);

/* rdf_eval_tree replaces kernel `evaluate_image_using_tree` (src/cuda/tree_eval.cu:140-212) / get_labels
 * (src/decision_tree.py:277-294): one tree straight from the canonical layout float32[2^D-1,7+2C], scale 1,
 * labels_reduce 1; a pixel whose walk falls off the last level is not written. */
RDF_API int rdf_eval_tree(const float* tree_dev, int max_depth, int num_classes, const uint16_t* depth_dev, int num_images,
                  int dim_x, int dim_y, uint16_t* labels_dev, void* stream);

/* Same result as rdf_eval_tree through a packed one-tree handle (rdf_forest_create(tree_dev, 1, D, C, ...)): the fast path
 * DecisionTreeEvaluator.get_labels uses for the per-tree evaluation of train_model.py (src/train_model.py:104-105). */
RDF_API int rdf_eval_tree_packed(const rdf_forest_t* tree, const uint16_t* depth_dev, int num_images, int dim_x, int dim_y,
                         uint16_t* labels_dev, void* stream);

/* rdf_composite replaces kernel `make_composite_labels_image` (src/cuda/tree_eval.cu:214-248) and its wrapper
 * (src/decision_tree.py:333-347).  label_images_dev: DEVICE array of L device pointers (the reference's int64
 * pointer table, src/decision_tree.py:203-207); conditions_dev int32[n_cond,2]; composite_dev uint16[dim_y,dim_x]. */
RDF_API int rdf_composite(const uint16_t* const* label_images_dev, int num_label_images, int dim_x, int dim_y,
                  const int32_t* conditions_dev, uint16_t* composite_dev, void* stream);

/* rdf_layered_run replaces the whole of LayeredDecisionForest.run (src/decision_tree.py:233-264): 1+L fills,
 * L evaluate_image_using_forest launches and make_composite_labels_image, in ONE launch.  Layer i is gated by
 * layer filter_model[i]'s label == filter_class[i] (filter_model[i] < 0: ungated); the gate is read from registers,
 * never from memory.  Every pixel of every per-layer image and of the composite is written (65535 where the
 * reference's pre-fill would have survived), so no pre-fill is needed.
 *   forests, filter_model, filter_class, labels_per_layer: HOST arrays of length L (labels_per_layer holds device
 *   pointers to uint16[h,w]);  depth_dev uint16[H,W];  composite_dev uint16[h,w]. */
RDF_API int rdf_layered_run(const rdf_forest_t* const* forests, int num_layers, const int* filter_model,
                    const int* filter_class, const uint16_t* depth_dev, int dim_x, int dim_y,
                    uint16_t* const* labels_per_layer, const int32_t* conditions_dev, int n_cond,
                    uint16_t* composite_dev, int labels_reduce, float scale, void* stream);

/* rdf_layered_run_batch: rdf_layered_run over num_images images in one launch (depth_dev uint16[N,dim_y,dim_x], every
 * labels_per_layer[i] and composite_dev uint16[N,h,w]) - the two hands of the live product, which the reference evaluates one
 * after the other (src/3d_bz.py:281-285).  Bit n of composite_flip_x_mask writes the composite label of image n mirrored in x,
 * i.e. pixel (y,x) at (y, w-1-x): the left hand is evaluated on an x-mirrored depth image and its label image is mirrored back
 * before mean shift (labels_image_2.set + flip_x, src/3d_bz.py:439-446).  Per-layer label images stay unmirrored. */
RDF_API int rdf_layered_run_batch(const rdf_forest_t* const* forests, int num_layers, const int* filter_model,
                          const int* filter_class, const uint16_t* depth_dev, int num_images, int dim_x, int dim_y,
                          uint16_t* const* labels_per_layer, const int32_t* conditions_dev, int n_cond,
                          uint16_t* composite_dev, int labels_reduce, float scale, unsigned composite_flip_x_mask, void* stream);

/* rdf_upload_frame: the live frame's host-to-device copy (depth_image.cu().set(np), src/3d_bz.py:156-157) as a kernel that
 * reads PINNED, device-mapped host memory and writes device memory; both 16-byte aligned.  In a per-frame CUDA graph it chains
 * to rdf_layered_run by programmatic dependent launch. */
RDF_API int rdf_upload_frame(const void* host_pinned, void* dev, size_t bytes, void* stream);

/* ---- mean shift -----------------------------------------------------------------------------------------
 * rdf_mean_shift replaces MeanShift.run (src/cuda/mean_shift.py:19-59): per round 1 fill + kernel `run`
 * (src/cuda/mean_shift.cu:3-48) + 2 blocking D2H + host divide + H2D, all `rounds` rounds in ONE launch with no
 * host round trip.  labels_dev uint16[h,w]; variances_dev float32[K]; means_dev float64[K,2] = (x,y) per class,
 * NaN for classes without pixels.  Sums are fp64 and deterministic (fixed reduction order).
 * workspace_dev: caller-provided scratch of at least rdf_mean_shift_workspace_bytes(w,h,K) bytes. */
RDF_API int rdf_mean_shift_workspace_bytes(int dim_x, int dim_y, int num_labels, size_t* bytes);
RDF_API int rdf_mean_shift(const uint16_t* labels_dev, int dim_x, int dim_y, int num_labels, const float* variances_dev,
                   int rounds, double* means_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* rdf_mean_shift_batch: rdf_mean_shift over num_images label images in one launch (labels_dev uint16[N,h,w], means_dev
 * float64[N,K,2]; the two hands of the live product, src/3d_bz.py:458-462 called once per hand).  Same workspace size. */
RDF_API int rdf_mean_shift_batch(const uint16_t* labels_dev, int num_images, int dim_x, int dim_y, int num_labels,
                         const float* variances_dev, int rounds, double* means_dev, void* workspace_dev, size_t workspace_bytes,
                         void* stream);

/* ---- hand grouping (the step before the forest in the live product) -------------------------------
 * rdf_group_hands replaces the D2H copy + CppGrouping.make_groups (src/cpp_grouping/grouping.cpp:80-191, binding
 * src/cpp_grouping/cpp_grouping.pyx:15-26, call site src/3d_bz.py:222-231) + H2D copy + write_pixel_groups_to_stencil_image
 * (src/3d_bz.py:239-250) by one launch: 4-connected components of the non-zero pixels of img_dev uint16[dim_y,dim_x] (the
 * 1/8-resolution depth image; at most 16384 pixels), components with size/(dim_x*dim_y) <= pct_thresh dropped, the largest
 * component whose centroid x < dim_x/2 becomes group 1 ("right"), the largest other one group 2 ("left"), ties -> first in
 * raster order.  stencil_dev uint16[dim_y,dim_x] receives 1 / 2 / 0; g_info_dev float32[2][3] = (size, centroid x,
 * centroid y) per group (zeros for an empty group; the reference leaves the centroid unspecified there). */
RDF_API int rdf_group_hands(const uint16_t* img_dev, int dim_x, int dim_y, float pct_thresh, uint16_t* stencil_dev,
                    float* g_info_dev, void* stream);

/* ---- live-frame conditioning before the forest and fingertip read-out after it --------------
 * rdf_condition_depth replaces, in ONE launch, deproject_points + transform_points + filter_points_by_plane +
 * remove_missing_3d_points_from_depth_image + depth_image_2.set + gaussian_depth_filter + shrink_image
 * (src/3d_bz.py:159-220; src/cuda/points_ops.cu:5-36,66-75,131-146,327-373,375-404; src/cuda/calibrated_plane.cu:30-45):
 * a depth sample d > 0 at (x,y) is deprojected with (ppx, ppy, focal), moved into plane space by plane_dev (float32[4][4],
 * row-major as the numpy matrix CalibratedPlane.get_mat() returns) and zeroed when its plane-space z > -plane_z_threshold; the
 * k_size x k_size weights gauss_dev (float32, device; NULL = no filter, the reference's gauss_sigma <= 0.1) then smooth the
 * result with the reference's zero-aware rule (0 when the zero samples outweigh the others, else floor of the weighted mean of
 * the non-zero samples); depth_out_dev uint16[dim_y,dim_x] receives it and depth_mm_dev (nullable)
 * uint16[dim_y >> mipmap_level, dim_x >> mipmap_level] its point-sampled reduction.  fp32 operation order = the reference's
 * compiled kernels, results bit-identical.  depth_out_dev must not alias depth_in_dev (the raw frame stays intact for
 * rdf_fingertip_z). */
RDF_API int rdf_condition_depth(const uint16_t* depth_in_dev, int dim_x, int dim_y, float ppx, float ppy, float focal,
                        const float* plane_dev, float plane_z_threshold, const float* gauss_dev, int k_size, int mipmap_level,
                        uint16_t* depth_out_dev, uint16_t* depth_mm_dev, void* stream);

/* The reference's pre-forest point kernels one by one (csrc/rdf_points.cu), for callers that drive them individually with a
 * user-visible point image pts_dev float32[N*H*W][4] (16-byte aligned) - src/run_live.py:86-121, src/run_live_layered.py:87-122,
 * src/3d_bz.py:159-259,390-420.  Same results as the reference's kernels bit for bit (fp32 operation order of its compiled code);
 * the fused product path is rdf_condition_depth / rdf_stencil_hands below.
 *   rdf_deproject_points        deproject_points (src/cuda/points_ops.cu:5-36): d > 0 -> (d*(x-ppx)/f, d*(y-ppy)/f, d, 1); d == 0
 *                               leaves the point as it was.
 *   rdf_transform_points        transform_points (:66-75): p <- M p for points with w == 1; mat_host = 16 floats, row-major (numpy),
 *                               read at call time (passed to the kernel by value like the reference's glm::mat4 argument).
 *   rdf_filter_points_by_plane  filter_points_by_plane (src/cuda/calibrated_plane.cu:30-45): w == 1 and z > -threshold -> (0,0,0,0).
 *   rdf_remove_missing_points   remove_missing_3d_points_from_depth_image (points_ops.cu:131-146): w == 0 -> depth 0.
 *   rdf_setup_depth_for_forest  setup_depth_image_for_forest (:149-165): depth 0 or w == 0 -> depth 65535.
 *   rdf_zeros_to_no_pixel       convert_0s_to_maxuint (:118-127).
 *   rdf_shrink_image            shrink_image (:375-404): out[y][x] = in[y << level][x << level], out is (dim_y >> level) x (dim_x >> level).
 *   rdf_stencil_by_group        stencil_depth_image_by_group (:441-463): out = depth where groups[y >> level][x >> level] == group,
 *                               other pixels of out are left untouched.
 *   rdf_scatter_groups          write_pixel_groups_to_stencil_image (:486-503): stencil[coords[i][0]][coords[i][1]] = coords[i][2],
 *                               coords_dev int32[num_coords][3], stencil uint16[rows][cols]. */
RDF_API int rdf_deproject_points(const uint16_t* depth_dev, int num_images, int dim_x, int dim_y, float ppx, float ppy, float focal,
                         float* pts_dev, void* stream);
RDF_API int rdf_transform_points(int num_pts, float* pts_dev, const float* mat_host, void* stream);
RDF_API int rdf_filter_points_by_plane(int num_pts, float plane_z_threshold, float* pts_dev, void* stream);
RDF_API int rdf_remove_missing_points(int num_pts, const float* pts_dev, uint16_t* depth_dev, void* stream);
RDF_API int rdf_setup_depth_for_forest(int num_pts, const float* pts_dev, uint16_t* depth_dev, void* stream);
RDF_API int rdf_zeros_to_no_pixel(int num_pixels, uint16_t* depth_dev, void* stream);
RDF_API int rdf_shrink_image(const uint16_t* in_dev, int dim_x, int dim_y, int mipmap_level, uint16_t* out_dev, void* stream);
RDF_API int rdf_stencil_by_group(const uint16_t* groups_dev, const uint16_t* depth_dev, int dim_x, int dim_y, int mipmap_level, int group,
                         uint16_t* out_dev, void* stream);
RDF_API int rdf_scatter_groups(const int32_t* coords_dev, int num_coords, uint16_t* stencil_dev, int rows, int cols, void* stream);

/* rdf_grow_groups: grow_groups (src/cuda/points_ops.cu:407-438, call site src/3d_bz.py:252-259): a zero pixel takes the first
 * non-zero value among its left, right, upper, lower neighbour. */
RDF_API int rdf_grow_groups(const uint16_t* groups_in_dev, int dim_x, int dim_y, uint16_t* groups_out_dev, void* stream);

/* rdf_stencil_hands replaces, for all hands in ONE launch, run_per_hand_pipeline's pre-processing (src/3d_bz.py:390-420):
 * depth_image_group.fill(0) + stencil_depth_image_by_group + flip_x (or copy) + convert_0s_to_maxuint
 * (src/cuda/points_ops.cu:118-129,441-483).  groups_dev uint16[dim_y >> level, dim_x >> level] is the group image
 * (rdf_group_hands' stencil with grow != 0 - grow_groups is then applied on the fly - or an already grown image with grow = 0).
 * out_dev uint16[num_hands, dim_y, dim_x]: hand i keeps the samples of depth_dev whose group is group_ids[i], mirrored in x
 * when flip_x[i] != 0, everything else (and every 0) = 65535.  group_ids / flip_x are host arrays read at call time. */
RDF_API int rdf_stencil_hands(const uint16_t* depth_dev, int dim_x, int dim_y, const uint16_t* groups_dev, int mipmap_level, int grow,
                      int num_hands, const int* group_ids, const int* flip_x, uint16_t* out_dev, void* stream);

/* rdf_flip_x: flip_x (src/cuda/points_ops.cu:466-483). */
RDF_API int rdf_flip_x(const uint16_t* in_dev, int dim_x, int dim_y, uint16_t* out_dev, void* stream);

/* rdf_labels_to_rgba: make_rgba_from_labels (src/cuda/points_ops.cu:258-281, call site src/3d_bz.py:448-456); colors_dev
 * uint8[num_colors,4], rgba_dev uint8[dim_y,dim_x,4]; pixels labelled 0 / 65535 are left untouched.
 * rdf_depth_to_rgba: make_depth_rgba (src/cuda/points_ops.cu:283-325, call site src/3d_bz.py:266-274). */
RDF_API int rdf_labels_to_rgba(const uint16_t* labels_dev, int dim_x, int dim_y, const uint8_t* colors_dev, int num_colors,
                       uint8_t* rgba_dev, void* stream);
RDF_API int rdf_depth_to_rgba(const uint16_t* depth_dev, int dim_x, int dim_y, int d_min, int d_max, uint8_t* rgba_dev, void* stream);

/* rdf_fingertip_z replaces the per-fingertip host loop after mean shift (src/3d_bz.py:503-522), for num_images hands at once
 * (means_dev float64[num_images,num_labels,2], all looked up in the same raw frame): for fingertip i with class
 * f = fingertip_labels[i] (host array, 1-based label ids): (px,py) = int32(means[f-1]) * labels_reduce; outside the frame (or a
 * NaN centroid) -> z_out[i] = NaN (the reference's reset_positions()); else z = raw_depth[py,px], deprojected with
 * (ppx,ppy,fx,fy) in fp32 like rs2_deproject_pixel_to_point without distortion, z_out[i] = -(plane[2,:] . (pt,1)) in fp64.
 * z_out float64[num_images,num_fingertips] and means_copy_out (nullable, float64[num_images,num_labels,2], receives a copy of
 * means_dev) may be device memory or pinned host memory, so that this launch is the frame's only writer to the host. */
RDF_API int rdf_fingertip_z(const double* means_dev, int num_images, int num_labels, const int* fingertip_labels, int num_fingertips,
                    int labels_reduce, const uint16_t* raw_depth_dev, int dim_x, int dim_y, float ppx, float ppy, float fx, float fy,
                    const float* plane_dev, double* z_out, double* means_copy_out, void* stream);

/* rdf_mean_shift_fingertips = rdf_mean_shift_batch + rdf_fingertip_z in ONE launch where the class-parallel mean-shift kernel
 * applies (the thread that finishes a class' centroid reads its fingertip depth out at once), otherwise the two launches; same
 * results as the two calls.  raw_dim_x / raw_dim_y: size of the raw frame (labels image size x labels_reduce in the product). */
RDF_API int rdf_mean_shift_fingertips(const uint16_t* labels_dev, int num_images, int dim_x, int dim_y, int num_labels,
                              const float* variances_dev, int rounds, double* means_dev, void* workspace_dev, size_t workspace_bytes,
                              const int* fingertip_labels, int num_fingertips, int labels_reduce, const uint16_t* raw_depth_dev,
                              int raw_dim_x, int raw_dim_y, float ppx, float ppy, float fx, float fy, const float* plane_dev,
                              double* z_out, double* means_copy_out, void* stream);

/* ---- synthetic inputs (bench / tests; bit-exact twins of rdf_b200/synth.py) -------------------------------
 * kind: 0 dense-smooth, 1 dense-noise, 2 live-mask.  Frames first_frame .. first_frame+N-1. */
RDF_API int rdf_synth_depth(uint16_t* depth_dev, int kind, int num_images, int dim_x, int dim_y, uint32_t seed,
                    int first_frame, void* stream);
RDF_API int rdf_synth_forest(float* canon_dev, int num_trees, int max_depth, int num_classes, uint32_t seed, void* stream);

/* ---- self test ---------------------------------------------------------------------------------------------
 * The kernels replace the four IEEE divides of compute_feature by a correctly rounded reciprocal (once per pixel) plus
 * one correction step (csrc/rdf_common.cuh).  This entry point compares that sequence against div.rn.f32 bit for bit
 * over 65535 divisors x cases_per_divisor numerators (uniform-in-exponent, adversarial next-to-integer quotients,
 * reference-like offsets) and returns the number of differing quotients / differing floors (both must be 0). Synchronous. */
RDF_API int rdf_selftest_fastdiv(unsigned cases_per_divisor, uint32_t seed, unsigned long long* mismatches_host,
                         unsigned long long* floor_mismatches_host);

/* ---- training split search --------------------------------------------------------------------------------
 * Level-synchronous training of one tree (DecisionTreeTrainer.train, src/decision_tree.py:444-601).
 *
 * rdf_train_hist replaces kernel `evaluate_random_features` (src/cuda/tree_train.cu:4-64), generalised to NT sorted
 * thresholds per feature (the reference is NT = 1): one feature evaluation per (pixel, feature), then
 *     hist[slot][feature][bin][label] += 1,  bin = #{k : thresholds[feature][k] <= f}  in 0..NT
 * so the reference's left child for threshold k (f < t_k) is bins 0..k.
 *   nodes_by_pixel_dev int32[N,H,W] (-1 = inactive, else node index within the current level);
 *   node_slot_dev int32[2^level]: node index -> histogram slot, or -1 (node not in this block);
 *   offsets_dev float32[F,4] (ux,uy,vx,vy); thresholds_dev float32[F,NT];
 *   PRECONDITION (rdf_train_hist, rdf_train_hist_bucketed, rdf_train_hist_bucketed_p2p): every feature's NT thresholds are
 *   FINITE and sorted ASCENDING - the bin is found by bisection, and pick-best / advance-pixels read the same table.  The
 *   library does not check this on the device; the host mirror (DecisionTreeTrainer._next_proposals) sorts each row and
 *   raises on NaN / inf before uploading.
 *   hist_dev uint32[num_slots,F,NT+1,C], ACCUMULATED into (zero it first; multi-GPU: allreduce it afterwards).
 * rdf_train_pick_best replaces `pick_best_features` (src/cuda/tree_train.cu:99-236) incl. the Gini helpers (:66-97),
 *   reading hist_dev; candidate order is feature-major, threshold-minor; first strictly greatest gain wins.
 * rdf_train_next_active replaces `get_active_nodes_next_level` (:238-273) with a deterministic (ascending) order.
 * rdf_train_advance_pixels replaces `copy_pixel_groups` (:275-324). */
RDF_API int rdf_train_hist(const uint16_t* depth_dev, const uint16_t* labels_dev, const int32_t* nodes_by_pixel_dev,
                   int num_images, int dim_x, int dim_y, const int32_t* node_slot_dev, int num_slots,
                   const float* offsets_dev, const float* thresholds_dev, int num_features, int num_thresholds,
                   int num_classes, uint32_t* hist_dev, void* stream);
/* Bucketed form of rdf_train_hist (same histogram, bit for bit): rdf_train_bucket groups the active pixels by histogram
 * slot once per (level, node block) into a caller-provided workspace; rdf_train_hist_bucketed then builds the histogram of
 * any proposal block from it with a shared-memory histogram per node whatever the number of nodes. */
RDF_API int rdf_train_bucket_workspace_bytes(int64_t num_pixels, int num_slots, size_t* bytes);
RDF_API int rdf_train_bucket(const int32_t* nodes_by_pixel_dev, int64_t num_pixels, const int32_t* node_slot_dev, int num_slots,
                     void* workspace_dev, size_t workspace_bytes, void* stream);
RDF_API int rdf_train_hist_bucketed(const uint16_t* depth_dev, const uint16_t* labels_dev, int num_images, int dim_x, int dim_y,
                            const void* bucket_workspace_dev, int num_slots, const float* offsets_dev,
                            const float* thresholds_dev, int num_features, int num_thresholds, int num_classes,
                            uint32_t* hist_dev, void* stream);
/* Multi-GPU split search with the reduction fused into the histogram kernel (the path's only exchange step):
 * the F features of a proposal block are owned in contiguous slices of Fo = ceil(F / world) by the ranks; owner_hist_dev is
 * a DEVICE array of `world` peer-mapped pointers (NVLink), rank r's buffer being uint32[num_slots][Fo][NT+1][C], 8-byte aligned
 * (zeroed by its owner, all ranks synchronised before and after the launch; with an even class count two neighbouring counters
 * are flushed as one 64-bit reduction).  Every rank runs rdf_train_hist_bucketed_p2p on its own
 * pixels; counters are flushed as system-scope reductions straight into the owner's buffer, so after a barrier rank r holds
 * the FULL histogram of its feature slice - a reduce-scatter without a separate collective.  rdf_train_pick_candidates then
 * scores the local slice (hist_local_dev laid out [num_slots][feature_stride][NT+1][C], the first num_local_features of each
 * slot valid, feature_offset = r * Fo) and writes, per active node, the best local (gain, global candidate
 * index, child counts); after an all-gather of those small records rdf_train_pick_finalize picks the global winner
 * (greatest gain, ties -> smallest index: what one GPU scanning all features in order keeps) and writes the node record. */
RDF_API int rdf_train_hist_bucketed_p2p(const uint16_t* depth_dev, const uint16_t* labels_dev, int num_images, int dim_x, int dim_y,
                                const void* bucket_workspace_dev, int num_slots, const float* offsets_dev,
                                const float* thresholds_dev, int num_features, int num_thresholds, int num_classes,
                                uint32_t* const* owner_hist_dev, int world, void* stream);
RDF_API int rdf_train_pick_candidates(int num_active, const int32_t* active_nodes_dev, const int32_t* node_slot_dev,
                              const uint64_t* parent_counts_dev, const uint32_t* hist_local_dev, int num_slots,
                              int num_local_features, int feature_stride, int feature_offset, int num_thresholds,
                              int num_classes, float* cand_gain_dev, int32_t* cand_idx_dev, uint64_t* cand_counts_dev,
                              void* stream);
RDF_API int rdf_train_pick_finalize(int num_active, const int32_t* active_nodes_dev, const int32_t* node_slot_dev,
                            const uint64_t* parent_counts_dev, int world, const float* all_gain_dev,
                            const int32_t* all_idx_dev, const uint64_t* all_counts_dev, const float* offsets_dev,
                            const float* thresholds_dev, int num_thresholds, int num_classes, int level, int max_depth,
                            float* tree_dev, uint64_t* next_counts_dev, float* best_gain_dev, void* stream);
RDF_API int rdf_train_pick_best(int num_active, const int32_t* active_nodes_dev, const int32_t* node_slot_dev,
                        const uint64_t* parent_counts_dev, const uint32_t* hist_dev, int num_slots,
                        const float* offsets_dev, const float* thresholds_dev, int num_features, int num_thresholds,
                        int num_classes, int level, int max_depth, float* tree_dev, uint64_t* next_counts_dev,
                        float* best_gain_dev, void* stream);
RDF_API int rdf_train_next_active(const float* tree_dev, int level, int max_depth, int num_classes,
                          const int32_t* active_nodes_dev, int num_active, int32_t* next_active_dev,
                          int32_t* num_next_active_dev, void* stream);
RDF_API int rdf_train_advance_pixels(const uint16_t* depth_dev, int32_t* nodes_by_pixel_dev, int num_images, int dim_x,
                             int dim_y, const float* tree_dev, int level, int max_depth, int num_classes, void* stream);
/* root statistics (src/decision_tree.py:452-467): node_counts[0][label] and nodes_by_pixel = (label > 0 ? 0 : -1) */
RDF_API int rdf_train_init(const uint16_t* labels_dev, int64_t num_pixels, int num_classes, int32_t* nodes_by_pixel_dev,
                   uint64_t* root_counts_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RDF_B200_H */
