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
    printf("abi_smoke ok (device path)\n");
    return 0;
}
