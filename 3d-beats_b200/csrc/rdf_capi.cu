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
                                int64_t total_nodes, int64_t nodes_per_tree, int64_t rows_per_tree, int layout, int D, int C, int CP,
                                int* __restrict__ exact_flag) {
    const int E = 7 + 2 * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_nodes; i += (int64_t)gridDim.x * blockDim.x) {
        const float* nd = canon + i * E;
        const int64_t t = i / nodes_per_tree, row = i - t * nodes_per_tree;
        const int j = 63 - __clzll(row + 1);                             // level of the canonical row
        const int64_t g = row + 1 - ((int64_t)1 << j);                   // index within the level
        const bool last = j == D - 1;                                    // level D-1: no children
        int64_t self, left, right;
        if (layout == RDF_LAYOUT_BLOCKS) {
            self = t * rows_per_tree + rdf_blocks_row(j, g, D);
            left = last ? 0 : t * rows_per_tree + rdf_blocks_row(j + 1, 2 * g, D);
            right = last ? 0 : t * rows_per_tree + rdf_blocks_row(j + 1, 2 * g + 1, D);
        } else {
            self = i;
            left = t * nodes_per_tree + 2 * row + 1;
            right = left + 1;
        }
        rdf_node_hdr h;
        h.a = make_float4(nd[0], nd[1], nd[2], nd[3]);
        h.ithresh = rdf_int_thresh(nd[4]);
        h.left = __float2int_rd(nd[5]) == -1 ? (last ? RDF_NO_LEAF : (int)left) : ~(int)(2 * self);
        h.right = __float2int_rd(nd[6]) == -1 ? (last ? RDF_NO_LEAF : (int)right) : ~(int)(2 * self + 1);
        // offsets outside the domain of the reciprocal divide + magic-number floor (rdf_common.cuh) -> exact path
        h.flags = (rdf_fastfloor_domain(nd[0]) && rdf_fastfloor_domain(nd[1]) && rdf_fastfloor_domain(nd[2]) &&
                   rdf_fastfloor_domain(nd[3])) ? 0 : RDF_FLAG_EXACT_DIV;
        hdr[self] = h;
        // forest properties the host dispatches on (bit 0: some node needs the exact divide; bit 1: some walk can end above the
        // last level, i.e. the forest is not a complete tree of depth D)
        if (h.flags & RDF_FLAG_EXACT_DIV) atomicOr(exact_flag, 1);
        if (!last && (h.left < 0 || h.right < 0)) atomicOr(exact_flag, 2);
        float* p = pdf + self * 2 * CP;
        for (int c = 0; c < CP; c++) {
            p[c] = c < C ? nd[7 + c] : 0.f;
            p[CP + c] = c < C ? nd[7 + C + c] : 0.f;
        }
    }
}

static int rdf_pack(rdf_forest* f, const float* canon_dev, cudaStream_t stream) {
    const int64_t total = f->nodes_per_tree * f->T;
    int blocks = (int)((total + 255) / 256);
    if (blocks > rdf_sm_count() * 32) blocks = rdf_sm_count() * 32;
    RDF_CUDA(cudaMemsetAsync(f->exact_flag_dev, 0, sizeof(int), stream));
    rdf_pack_kernel<<<blocks, 256, 0, stream>>>(canon_dev, f->hdr, f->pdf, total, f->nodes_per_tree, f->rows_per_tree, f->layout, f->D, f->C,
                                                f->CP, f->exact_flag_dev);
    RDF_LAUNCH_CHECK("rdf_pack_kernel");
    // the flag is needed on the host to choose kernels: packing is handle creation / update, not the per-frame path
    int props = 0;
    RDF_CUDA(cudaMemcpyAsync(&props, f->exact_flag_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
    RDF_CUDA(cudaStreamSynchronize(stream));
    f->has_exact_nodes = props & 1;
    f->has_early_leaves = (props >> 1) & 1;
    // host copy of the upper levels for the eval kernel's launch parameters (rdf_eval.cu: rdf_eval_launch)
    f->top_levels = 0;
    if (f->T <= 8) {                                                  // RDF_FAST_MAX_TREES: the per-pixel kernels' limit
        const int KT = rdf_top_levels(f->T);
        const int KS = f->D < KT ? f->D : KT, M = (1 << KS) - 1, first_last = (1 << (KS - 1)) - 1;
        delete[] f->top_host;
        f->top_host = new rdf_node_hdr[(size_t)f->T * M];
        for (int t = 0; t < f->T; t++) {
            RDF_CUDA(cudaMemcpy(f->top_host + t * M, f->hdr + (size_t)t * f->rows_per_tree, sizeof(rdf_node_hdr) * M, cudaMemcpyDeviceToHost));
            const int shift = t * M - (int)(t * f->rows_per_tree);
            for (int row = 0; row < first_last; row++) {
                rdf_node_hdr& h = f->top_host[t * M + row];
                if (h.left >= 0) h.left += shift;
                if (h.right >= 0) h.right += shift;
            }
        }
        f->top_levels = KS;
    }
    return RDF_OK;
}

// RDF_PACK_LAYOUT=heap|blocks chooses the node order of handles created afterwards (read once; default: see rdf_pack_layout)
static int rdf_pack_layout() {
    const char* e = RDF_GETENV_ONCE("RDF_PACK_LAYOUT");
    if (e && !strcmp(e, "blocks")) return RDF_LAYOUT_BLOCKS;
    if (e && !strcmp(e, "heap")) return RDF_LAYOUT_HEAP;
    return RDF_LAYOUT_DEFAULT;
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
    RDF_REQUIRE((int64_t)num_trees * rdf_blocks_rows_per_tree(max_depth) < ((int64_t)1 << 30) &&
                    ((int64_t)num_trees << max_depth) < ((int64_t)1 << 30),
                "rdf_forest_create: %d trees of depth %d exceed the 2^30 node slots a packed forest can index", num_trees, max_depth);
    rdf_forest* f = new rdf_forest();
    f->T = num_trees;
    f->D = max_depth;
    f->C = num_classes;
    f->CP = (num_classes + 3) & ~3;
    f->nodes_per_tree = ((int64_t)1 << max_depth) - 1;
    f->layout = rdf_pack_layout();
    f->rows_per_tree = f->layout == RDF_LAYOUT_BLOCKS ? rdf_blocks_rows_per_tree(max_depth) : f->nodes_per_tree;
    f->hdr = nullptr;
    f->pdf = nullptr;
    f->top_host = nullptr;
    f->top_levels = 0;
    const size_t n = (size_t)f->rows_per_tree * f->T;
    const size_t hdr_bytes = n * sizeof(rdf_node_hdr), pdf_bytes = n * 2 * f->CP * sizeof(float);
    f->packed_bytes = hdr_bytes + pdf_bytes;
    cudaError_t e = cudaGetDevice(&f->device);
    if (e == cudaSuccess) e = cudaMalloc(&f->hdr, hdr_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&f->pdf, pdf_bytes);
    f->exact_flag_dev = nullptr;
    f->has_exact_nodes = 0;
    f->has_early_leaves = 1;
    if (e == cudaSuccess) e = cudaMalloc(&f->exact_flag_dev, sizeof(int));
    if (e != cudaSuccess) {
        rdf_set_error("rdf_forest_create: allocating %zu packed bytes failed: %s", f->packed_bytes, cudaGetErrorString(e));
        if (f->hdr) cudaFree(f->hdr);
        if (f->pdf) cudaFree(f->pdf);
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
    if (forest->exact_flag_dev) cudaFree(forest->exact_flag_dev);
    delete[] forest->top_host;
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

// ---- self test: reciprocal-based divide == div.rn.f32, bit for bit, on its whole domain ---------------------------------
__device__ __forceinline__ uint32_t st_mix(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256) rdf_selftest_fastdiv_kernel(unsigned cases_per_d, uint32_t seed,
                                                                   unsigned long long* __restrict__ mismatches,
                                                                   unsigned long long* __restrict__ floor_mismatches) {
    const unsigned long long total = 65535ull * cases_per_d;
    unsigned long long bad = 0, bad_floor = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned d = 1u + (unsigned)(i % 65535ull);
        const unsigned k = (unsigned)(i / 65535ull);
        const uint32_t h0 = st_mix(seed ^ (d * 0x9E3779B1u) ^ (k * 0x85EBCA77u));
        const uint32_t h1 = st_mix(h0 ^ 0xC2B2AE3Du);
        const float df = (float)d;
        float a;
        switch (k & 3u) {
            case 0: {   // log-uniform magnitude over the whole domain 2^-60 .. 2^60
                const uint32_t expo = 127u - 60u + (h0 >> 8) % 120u;
                a = __uint_as_float(((h0 & 1u) << 31) | (expo << 23) | (h1 & 0x7FFFFFu));
                break;
            }
            case 1: {   // adversarial: a within a few ulp of an integer multiple of d (quotient next to an integer)
                const int n = (int)(h0 % 2097153u) - 1048576;
                a = (float)n * df;
                a = __uint_as_float(__float_as_uint(a) + (h1 % 9u) - 4u);
                break;
            }
            case 2: {   // small quotients: |n| <= 512, the range of real probe offsets
                const int n = (int)(h0 % 1025u) - 512;
                a = (float)n * df;
                a = __uint_as_float(__float_as_uint(a) + (h1 % 17u) - 8u);
                break;
            }
            default: {  // reference-like offsets: magnitude e^U(0,14), any sign
                const float m = __expf(14.f * (float)(h0 >> 8) / 16777216.f);
                a = (h1 & 1u) ? -m : m;
                a = __uint_as_float(__float_as_uint(a) ^ ((h1 >> 1) & 0xFFu));
                break;
            }
        }
        if (!rdf_fastdiv_domain(a)) continue;
        const float exact = __fdiv_rn(a, df);
        const float fast = rdf_div_fast(a, df, __frcp_rn(df));
        if (__float_as_uint(exact) != __float_as_uint(fast)) bad++;
        if (__float2int_rd(exact) != __float2int_rd(fast)) bad_floor++;
    }
    if (bad) atomicAdd(mismatches, bad);
    if (bad_floor) atomicAdd(floor_mismatches, bad_floor);
}

extern "C" int rdf_selftest_fastdiv(unsigned cases_per_divisor, uint32_t seed, unsigned long long* mismatches_host,
                                    unsigned long long* floor_mismatches_host) {
    RDF_REQUIRE(mismatches_host && floor_mismatches_host && cases_per_divisor >= 1, "rdf_selftest_fastdiv: bad argument");
    unsigned long long* dev = nullptr;
    RDF_CUDA(cudaMalloc(&dev, 2 * sizeof(unsigned long long)));
    cudaMemset(dev, 0, 2 * sizeof(unsigned long long));
    rdf_selftest_fastdiv_kernel<<<rdf_sm_count() * 8, 256>>>(cases_per_divisor, seed, dev, dev + 1);
    unsigned long long host[2] = {~0ull, ~0ull};
    cudaError_t e = cudaMemcpy(host, dev, sizeof(host), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (e != cudaSuccess) {
        rdf_set_error("rdf_selftest_fastdiv: %s", cudaGetErrorString(e));
        return RDF_ERR_CUDA;
    }
    *mismatches_host = host[0];
    *floor_mismatches_host = host[1];
    return RDF_OK;
}
