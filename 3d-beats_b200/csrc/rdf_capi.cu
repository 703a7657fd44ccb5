// Error plumbing + forest container (pack canonical layout -> 32-byte node headers + aligned leaf pdf table).
#include <stdarg.h>
#include <string.h>

#include "rdf_common.cuh"

static thread_local char g_err[512] = "";

void rdf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int rdf_version(void) { return RDF_B200_VERSION; }
extern "C" const char* rdf_last_error(void) { return g_err; }

// canonical node = (ux,uy,vx,vy,thresh,l_next,r_next,l_pdf[C],r_pdf[C]) (src/cuda/tree_eval.cu:47).
// "child continues" is floor(flag) == -1 exactly as the kernels test it (__float2int_rd, tree_eval.cu:101-102).
__global__ void rdf_pack_kernel(const float* __restrict__ canon, rdf_node_hdr* __restrict__ hdr, float* __restrict__ pdf,
                                int64_t total_nodes, int C, int CP) {
    const int E = 7 + 2 * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_nodes; i += (int64_t)gridDim.x * blockDim.x) {
        const float* nd = canon + i * E;
        rdf_node_hdr h;
        h.a = make_float4(nd[0], nd[1], nd[2], nd[3]);
        h.thresh = nd[4];
        h.flags = (__float2int_rd(nd[5]) == -1 ? 1 : 0) | (__float2int_rd(nd[6]) == -1 ? 2 : 0);
        h.pad0 = 0;
        h.pad1 = 0;
        hdr[i] = h;
        float* p = pdf + i * 2 * CP;
        for (int c = 0; c < CP; c++) {
            p[c] = c < C ? nd[7 + c] : 0.f;
            p[CP + c] = c < C ? nd[7 + C + c] : 0.f;
        }
    }
}

static int rdf_pack(rdf_forest* f, const float* canon_dev, cudaStream_t stream) {
    const int64_t total = f->nodes_per_tree * f->T;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    rdf_pack_kernel<<<blocks, 256, 0, stream>>>(canon_dev, f->hdr, f->pdf, total, f->C, f->CP);
    RDF_LAUNCH_CHECK("rdf_pack_kernel");
    return RDF_OK;
}

extern "C" int rdf_forest_create(const float* canon_dev, int num_trees, int max_depth, int num_classes, void* stream,
                                 rdf_forest_t** out) {
    RDF_REQUIRE(out != nullptr, "rdf_forest_create: out is NULL");
    *out = nullptr;
    RDF_REQUIRE(canon_dev != nullptr, "rdf_forest_create: canon_dev is NULL");
    RDF_REQUIRE(num_trees >= 1, "rdf_forest_create: num_trees=%d", num_trees);
    RDF_REQUIRE(max_depth >= 1 && max_depth <= RDF_MAX_DEPTH, "rdf_forest_create: max_depth=%d outside 1..%d", max_depth,
                RDF_MAX_DEPTH);
    RDF_REQUIRE(num_classes >= 1 && num_classes <= RDF_MAX_CLASSES, "rdf_forest_create: num_classes=%d outside 1..%d",
                num_classes, RDF_MAX_CLASSES);
    rdf_forest* f = new rdf_forest();
    f->T = num_trees;
    f->D = max_depth;
    f->C = num_classes;
    f->CP = (num_classes + 3) & ~3;
    f->nodes_per_tree = ((int64_t)1 << max_depth) - 1;
    f->hdr = nullptr;
    f->pdf = nullptr;
    const size_t n = (size_t)f->nodes_per_tree * f->T;
    const size_t hdr_bytes = n * sizeof(rdf_node_hdr), pdf_bytes = n * 2 * f->CP * sizeof(float);
    f->packed_bytes = hdr_bytes + pdf_bytes;
    cudaError_t e = cudaGetDevice(&f->device);
    if (e == cudaSuccess) e = cudaMalloc(&f->hdr, hdr_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&f->pdf, pdf_bytes);
    if (e != cudaSuccess) {
        rdf_set_error("rdf_forest_create: allocating %zu packed bytes failed: %s", f->packed_bytes, cudaGetErrorString(e));
        if (f->hdr) cudaFree(f->hdr);
        delete f;
        return RDF_ERR_CUDA;
    }
    int rc = rdf_pack(f, canon_dev, rdf_stream(stream));
    if (rc != RDF_OK) {
        rdf_forest_destroy(f);
        return rc;
    }
    *out = f;
    return RDF_OK;
}

extern "C" int rdf_forest_update(rdf_forest_t* forest, const float* canon_dev, void* stream) {
    RDF_REQUIRE(forest != nullptr && canon_dev != nullptr, "rdf_forest_update: NULL argument");
    return rdf_pack(forest, canon_dev, rdf_stream(stream));
}

extern "C" int rdf_forest_destroy(rdf_forest_t* forest) {
    if (!forest) return RDF_OK;
    if (forest->hdr) cudaFree(forest->hdr);
    if (forest->pdf) cudaFree(forest->pdf);
    delete forest;
    return RDF_OK;
}

extern "C" int rdf_forest_info(const rdf_forest_t* forest, int* num_trees, int* max_depth, int* num_classes,
                               size_t* packed_bytes) {
    RDF_REQUIRE(forest != nullptr, "rdf_forest_info: forest is NULL");
    if (num_trees) *num_trees = forest->T;
    if (max_depth) *max_depth = forest->D;
    if (num_classes) *num_classes = forest->C;
    if (packed_bytes) *packed_bytes = forest->packed_bytes;
    return RDF_OK;
}
