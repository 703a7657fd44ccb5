/* Plain-C consumer of include/rdf_b200.h: proves the boundary is a C ABI (no C++/torch types), compiled with gcc.
 * Without a CUDA device it exercises the error contract only; with one it evaluates a tiny hand-built forest and checks the
 * labels it must produce (the same case as tests/test_oracle_self.py::test_skip_and_probe_semantics).
 *   gcc -std=c99 -I include -I /usr/local/cuda/include tests/c/abi_smoke.c -o abi_smoke -L 3d-beats_b200/rdf_b200 -lrdf_b200 \
 *       -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,... */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rdf_b200.h"

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            printf("FAIL %s:%d %s (%s)\n", __FILE__, __LINE__, #cond, rdf_last_error()); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(void) {
    rdf_forest_t* h = NULL;
    CHECK(rdf_version() >= 100);
    CHECK(rdf_forest_create(NULL, 3, 16, 4, NULL, &h) == RDF_ERR_INVALID && h == NULL);
    CHECK(strstr(rdf_last_error(), "canon_dev") != NULL);
    CHECK(rdf_eval_tree(NULL, 4, 4, NULL, 1, 8, 8, NULL, NULL) == RDF_ERR_INVALID);
    size_t ws = 0;
    CHECK(rdf_mean_shift_workspace_bytes(424, 240, 11, &ws) == RDF_OK && ws > 0);
    CHECK(rdf_train_bucket_workspace_bytes(1000, 4, &ws) == RDF_OK && ws >= 4000);
    CHECK(rdf_condition_depth(NULL, 8, 8, 0.f, 0.f, 1.f, NULL, 40.f, NULL, 5, 3, NULL, NULL, NULL) == RDF_ERR_INVALID);
    CHECK(strstr(rdf_last_error(), "rdf_condition_depth") != NULL);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        printf("abi_smoke ok (no CUDA device: error contract only)\n");
        return 0;
    }
    /* one node, C = 3: u probes 2 px to the right (2000 / d with d = 1000), v probes the centre */
    enum { W = 8, H = 8, C = 3, E = 7 + 2 * C };
    float node[E] = {2000.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, /* left pdf */ 0.f, .5f, .25f, /* right pdf */ 0.f, .25f, .5f};
    uint16_t depth[W * H], labels[W * H];
    for (int i = 0; i < W * H; i++) { depth[i] = 1000; labels[i] = 9; }
    depth[0] = 0;
    depth[1] = 65535;
    float* node_d; uint16_t *depth_d, *labels_d;
    CHECK(cudaMalloc((void**)&node_d, sizeof(node)) == cudaSuccess);
    CHECK(cudaMalloc((void**)&depth_d, sizeof(depth)) == cudaSuccess);
    CHECK(cudaMalloc((void**)&labels_d, sizeof(labels)) == cudaSuccess);
    cudaMemcpy(node_d, node, sizeof(node), cudaMemcpyHostToDevice);
    cudaMemcpy(depth_d, depth, sizeof(depth), cudaMemcpyHostToDevice);
    cudaMemcpy(labels_d, labels, sizeof(labels), cudaMemcpyHostToDevice);
    CHECK(rdf_forest_create(node_d, 1, 1, C, NULL, &h) == RDF_OK && h != NULL);
    int t, d, c; size_t bytes;
    CHECK(rdf_forest_info(h, &t, &d, &c, &bytes) == RDF_OK && t == 1 && d == 1 && c == C && bytes > 0);
    CHECK(rdf_eval_forest(h, depth_d, 1, W, H, NULL, -1, labels_d, NULL, 1, 1.0f, NULL) == RDF_OK);
    CHECK(cudaMemcpy(labels, labels_d, sizeof(labels), cudaMemcpyDeviceToHost) == cudaSuccess);
    CHECK(labels[0] == 9 && labels[1] == 9);                       /* centre 0 / 65535: untouched */
    for (int y = 1; y < H; y++)
        for (int x = 0; x < W; x++) CHECK(labels[y * W + x] == (x < 6 ? 1 : 2));   /* in-image probe -> left, off-image -> right */
    /* same through the single-tree entry point */
    for (int i = 0; i < W * H; i++) labels[i] = 9;
    cudaMemcpy(labels_d, labels, sizeof(labels), cudaMemcpyHostToDevice);
    CHECK(rdf_eval_tree(node_d, 1, C, depth_d, 1, W, H, labels_d, NULL) == RDF_OK);
    CHECK(cudaMemcpy(labels, labels_d, sizeof(labels), cudaMemcpyDeviceToHost) == cudaSuccess);
    CHECK(labels[2 * W + 0] == 1 && labels[2 * W + 7] == 2 && labels[0] == 9);
    CHECK(rdf_forest_destroy(h) == RDF_OK);
    cudaFree(node_d); cudaFree(depth_d); cudaFree(labels_d);
    /* live-frame conditioning: identity plane, clip distance 500 (a sample survives when its z <= -500 ... here z = +d, so every
     * sample is clipped), then clip distance -2000 (z = d <= 2000 survives); no filter; 1/2 image; then the per-hand stencil */
    {
        enum { FW = 16, FH = 8 };
        uint16_t frame[FW * FH], cond[FW * FH], mm[(FW / 2) * (FH / 2)], groups[(FW / 2) * (FH / 2)], hands[2 * FW * FH];
        float plane[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        for (int i = 0; i < FW * FH; i++) frame[i] = (uint16_t)(i % 3 == 0 ? 0 : 1000 + 100 * (i % 16));
        uint16_t *frame_d, *cond_d, *mm_d, *groups_d, *hands_d; float* plane_d;
        CHECK(cudaMalloc((void**)&frame_d, sizeof(frame)) == cudaSuccess && cudaMalloc((void**)&cond_d, sizeof(cond)) == cudaSuccess);
        CHECK(cudaMalloc((void**)&mm_d, sizeof(mm)) == cudaSuccess && cudaMalloc((void**)&groups_d, sizeof(groups)) == cudaSuccess);
        CHECK(cudaMalloc((void**)&hands_d, sizeof(hands)) == cudaSuccess && cudaMalloc((void**)&plane_d, sizeof(plane)) == cudaSuccess);
        cudaMemcpy(frame_d, frame, sizeof(frame), cudaMemcpyHostToDevice);
        cudaMemcpy(plane_d, plane, sizeof(plane), cudaMemcpyHostToDevice);
        CHECK(rdf_condition_depth(frame_d, FW, FH, 8.f, 4.f, 400.f, plane_d, 500.f, NULL, 0, 1, cond_d, mm_d, NULL) == RDF_OK);
        CHECK(cudaMemcpy(cond, cond_d, sizeof(cond), cudaMemcpyDeviceToHost) == cudaSuccess);
        for (int i = 0; i < FW * FH; i++) CHECK(cond[i] == 0);                          /* z = d > -500: everything clipped */
        CHECK(rdf_condition_depth(frame_d, FW, FH, 8.f, 4.f, 400.f, plane_d, -2000.f, NULL, 0, 1, cond_d, mm_d, NULL) == RDF_OK);
        CHECK(cudaMemcpy(cond, cond_d, sizeof(cond), cudaMemcpyDeviceToHost) == cudaSuccess);
        CHECK(cudaMemcpy(mm, mm_d, sizeof(mm), cudaMemcpyDeviceToHost) == cudaSuccess);
        for (int i = 0; i < FW * FH; i++) CHECK(cond[i] == (frame[i] <= 2000 ? frame[i] : 0));   /* z = d > 2000 clipped */
        for (int y = 0; y < FH / 2; y++)
            for (int x = 0; x < FW / 2; x++) CHECK(mm[y * (FW / 2) + x] == cond[(2 * y) * FW + 2 * x]);
        for (int i = 0; i < (FW / 2) * (FH / 2); i++) groups[i] = (uint16_t)((i % (FW / 2)) < 4 ? 1 : 2);
        cudaMemcpy(groups_d, groups, sizeof(groups), cudaMemcpyHostToDevice);
        const int ids[2] = {1, 2}, flips[2] = {0, 1};
        CHECK(rdf_stencil_hands(cond_d, FW, FH, groups_d, 1, 0, 2, ids, flips, hands_d, NULL) == RDF_OK);
        CHECK(cudaMemcpy(hands, hands_d, sizeof(hands), cudaMemcpyDeviceToHost) == cudaSuccess);
        for (int y = 0; y < FH; y++)
            for (int x = 0; x < FW; x++) {
                const uint16_t v = cond[y * FW + x];
                const uint16_t right = (x < 8 && v) ? v : 65535, left = (x >= 8 && v) ? v : 65535;
                CHECK(hands[y * FW + x] == right);                                       /* group 1, as is */
                CHECK(hands[FW * FH + y * FW + (FW - 1 - x)] == left);                   /* group 2, mirrored */
            }
        cudaFree(frame_d); cudaFree(cond_d); cudaFree(mm_d); cudaFree(groups_d); cudaFree(hands_d); cudaFree(plane_d);
    }
    printf("abi_smoke ok (device path)\n");
    return 0;
}
