// Test-only harness around the REFERENCE's own kernels.  TEST INFRASTRUCTURE - never linked into the product.
//
// The three reference sources are #included from where they lie under /root/reference (found through -I, see
// oracle/Makefile target `ref`); nothing of them is copied into this repository.  The launchers below reproduce the
// launch geometry of the reference's host code (grid / block / dynamic shared memory), cited per function, so that
// both results and timing are those of the unmodified reference running on a B200.
#include <cuda_runtime.h>
#include <assert.h>   // upstream gets <cassert>/<cstdio> through GLM, which is un-vendored (stubbed empty)
#include <stdint.h>
#include <stdio.h>

#include <tree_eval.cu>     // reference: src/cuda/tree_eval.cu
#include <tree_train.cu>    // reference: src/cuda/tree_train.cu
#include <mean_shift.cu>    // reference: src/cuda/mean_shift.cu

#define REF_EXPORT extern "C" __attribute__((visibility("default")))
#define REF_MAX_THREADS_PER_BLOCK 1024   // src/decision_tree.py:16

static int ref_check() { return cudaGetLastError() == cudaSuccess ? 0 : -2; }

// DecisionTreeEvaluator.get_labels_forest, src/decision_tree.py:298-330
REF_EXPORT int ref_eval_forest(int num_trees, int num_images, int dim_x, int dim_y, int num_classes, int max_depth,
                               uint16_t* depth, int filter_class, uint16_t* filter, float* forest, uint16_t* labels,
                               int labels_reduce, float scale, void* stream) {
    const long long num_test_pixels = (long long)num_images * (dim_y / labels_reduce) * (dim_x / labels_reduce);
    const int BLOCK_DIM_X = REF_MAX_THREADS_PER_BLOCK / num_trees;
    dim3 grid((unsigned)(num_test_pixels / BLOCK_DIM_X) + 1, 1, 1), block(BLOCK_DIM_X, num_trees, 1);
    evaluate_image_using_forest<<<grid, block, (size_t)BLOCK_DIM_X * num_classes * 4, (cudaStream_t)stream>>>(
        num_trees, num_images, dim_x, dim_y, num_classes, max_depth, BLOCK_DIM_X, depth, filter ? filter_class : -1,
        filter ? filter : labels, forest, labels, labels_reduce, scale);
    return ref_check();
}

// DecisionTreeEvaluator.get_labels, src/decision_tree.py:277-294
REF_EXPORT int ref_eval_tree(int num_images, int dim_x, int dim_y, int num_classes, int max_depth, uint16_t* depth,
                             float* tree, uint16_t* labels, void* stream) {
    const long long npx = (long long)num_images * dim_y * dim_x;
    dim3 grid((unsigned)(npx / REF_MAX_THREADS_PER_BLOCK) + 1, 1, 1), block(REF_MAX_THREADS_PER_BLOCK, 1, 1);
    evaluate_image_using_tree<<<grid, block, 0, (cudaStream_t)stream>>>(num_images, dim_x, dim_y, num_classes, max_depth, depth,
                                                                        tree, labels);
    return ref_check();
}

// DecisionTreeEvaluator.make_composite_labels_image, src/decision_tree.py:333-347
REF_EXPORT int ref_composite(uint16_t** label_images, int num_label_images, int dim_x, int dim_y, int32_t* conditions,
                             uint16_t* composite, void* stream) {
    dim3 grid(dim_x / 32 + 1, dim_y / 32 + 1, 1), block(32, 32, 1);
    make_composite_labels_image<<<grid, block, 0, (cudaStream_t)stream>>>(label_images, num_label_images, dim_x, dim_y,
                                                                          (int2*)conditions, composite);
    return ref_check();
}

// one round of MeanShift.run's kernel, src/cuda/mean_shift.py:28-48
REF_EXPORT int ref_mean_shift_round(uint16_t* labels, int num_classes, int dim_x, int dim_y, float* variances, double* means,
                                    int iter_number, double* temp_sum, void* stream) {
    dim3 grid(dim_x / 32 + 1, dim_y / 32 + 1, 1), block(32, 32, 1);
    run<<<grid, block, 0, (cudaStream_t)stream>>>(labels, num_classes, dim_x, dim_y, variances, (double2*)means, iter_number,
                                                  temp_sum);
    return ref_check();
}

// DecisionTreeTrainer.train, src/decision_tree.py:512-534
REF_EXPORT int ref_evaluate_random_features(int num_images, int dim_x, int dim_y, int num_proposals, int num_classes,
                                            int max_depth, int max_next_nodes, int elig_min, int elig_max, uint16_t* labels,
                                            uint16_t* depth, float* proposals, int* nodes_by_pixel,
                                            unsigned long long* next_counts, void* stream) {
    const int bdx = REF_MAX_THREADS_PER_BLOCK / num_proposals;
    const long long npx = (long long)num_images * dim_x * dim_y;
    dim3 grid((unsigned)(npx / bdx) + 1, 1, 1), block(bdx, num_proposals, 1);
    evaluate_random_features<<<grid, block, 0, (cudaStream_t)stream>>>(num_images, dim_x, dim_y, num_proposals, num_classes,
                                                                       max_depth, max_next_nodes, elig_min, elig_max, labels,
                                                                       depth, proposals, nodes_by_pixel, next_counts);
    return ref_check();
}

// src/decision_tree.py:536-555
REF_EXPORT int ref_pick_best_features(int num_active, int num_proposals, int max_depth, int max_next_nodes, int elig_min,
                                      int elig_max, int num_classes, int level, int* active_nodes,
                                      unsigned long long* parent_counts, unsigned long long* child_counts_by_feature,
                                      float* proposals, float* tree, unsigned long long* child_counts, float* best_gain,
                                      void* stream) {
    dim3 grid(num_active / REF_MAX_THREADS_PER_BLOCK + 1, 1, 1), block(REF_MAX_THREADS_PER_BLOCK, 1, 1);
    pick_best_features<<<grid, block, 0, (cudaStream_t)stream>>>(num_active, num_proposals, max_depth, max_next_nodes, elig_min,
                                                                 elig_max, num_classes, level, active_nodes, parent_counts,
                                                                 child_counts_by_feature, proposals, tree, child_counts,
                                                                 best_gain);
    return ref_check();
}

// src/decision_tree.py:560-569
REF_EXPORT int ref_get_active_nodes_next_level(int level, int max_depth, int num_classes, float* tree, int* active_nodes,
                                               int num_active, int* next_active, int* num_next_active, void* stream) {
    get_active_nodes_next_level<<<dim3(num_active, 1, 1), dim3(1, 1, 1), 0, (cudaStream_t)stream>>>(
        level, max_depth, num_classes, tree, active_nodes, num_active, next_active, num_next_active);
    return ref_check();
}

// src/decision_tree.py:581-594
REF_EXPORT int ref_copy_pixel_groups(int num_images, int dim_x, int dim_y, int level, int max_depth, int num_classes,
                                     uint16_t* depth, int* nodes_by_pixel, float* tree, void* stream) {
    const long long npx = (long long)num_images * dim_x * dim_y;
    dim3 grid((unsigned)(npx / REF_MAX_THREADS_PER_BLOCK) + 1, 1, 1), block(REF_MAX_THREADS_PER_BLOCK, 1, 1);
    copy_pixel_groups<<<grid, block, 0, (cudaStream_t)stream>>>(num_images, dim_x, dim_y, level, max_depth, num_classes, depth,
                                                                nodes_by_pixel, tree);
    return ref_check();
}
