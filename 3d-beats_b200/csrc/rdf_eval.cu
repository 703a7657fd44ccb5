// Forest / tree evaluation: replaces evaluate_image_using_forest and evaluate_image_using_tree
// (reference src/cuda/tree_eval.cu:24-137, :140-212).
#include "rdf_traverse.cuh"

// ---- fast path: packed forest, T <= 8, one thread per labels pixel ---------------------------------------------
// Block = 256 threads over a 32x8 tile of labels pixels; a warp covers WARP_W x (32/WARP_W) pixels so that the
// lanes of a warp are 2-D neighbours (neighbouring pixels share tree paths and probe cache lines).
struct rdf_eval_params {
    rdf_forest_view fv;
    const uint16_t* depth;
    const uint16_t* filter;
    uint16_t* labels;
    float* probs;
    int W, H, w, h, r;
    int tiles_x;
    int filter_class;
    int image0;
    int top_levels;           // upper tree levels that travel in `top` below (min(D, rdf_top_levels(T)))
    int tree_mode;            // evaluate_image_using_tree semantics: no write when the walk reaches no leaf (tree_eval.cu:174-210)
    float scale;
};

// The upper levels of every tree travel with the launch, as KERNEL PARAMETERS: the headers sit in the constant bank and a walk
// reads them with register-indexed constant loads (`LDC.64 R, c[0x0][R+imm]`), which do not pass through the L1 data pipe - the
// unit this kernel saturates (97 % of peak, profiles/r02_ncu_eval.md).  Against the previous form (every 256-pixel CTA staged the
// same levels into shared memory and read them with broadcast LDS.128 pairs) this removes 9 % of the kernel's L1 wavefronts, the
// 4 KB of staging loads per CTA, the shared memory and the block barrier: cfg3 7.60 -> 8.22 Gpx/s, cfg3 noise 3.10 -> 3.34, cfg5
// 2.53 -> 2.63 (profiles/r02_eval_const_top.md).  A constant load serves one address per pass, so lanes on different nodes are
// serialised: 5 levels (<= 16 nodes per tree on the last one) is where divergent frames stop gaining - with 6 / 7 levels the smooth
// workload reaches 8.43 / 8.63 Gpx/s but dense-noise frames fall to 2.05 / 0.80, so RDF_EVAL_TOP_LEVELS stays 5.
// The launch parameters may be as large as 32 764 bytes (CUDA 12.1+, sm_70+); rdf_top_levels(T) keeps T trees below 30 000.
static_assert(RDF_EVAL_TOP_LEVELS <= RDF_PACK_TOP_LEVELS, "the levels that travel with the launch must lie in the heap-ordered top of a packed tree");
template <int T>
struct __align__(32) rdf_eval_launch {
    rdf_eval_params p;
    rdf_node_hdr top[T * ((1 << rdf_top_levels(T)) - 1)];     // [t][row], child ids of the levels above the last in this indexing
};
template <int T>
struct rdf_top_ref {                                          // how the walk sees the array: a reference, so loads stay in param space
    const rdf_node_hdr (&top)[T * ((1 << rdf_top_levels(T)) - 1)];
};

// CTAs per SM the register allocation aims for: 4 (<= 64 registers) up to 5 interleaved trees; the state of 6..8 trees does not
// fit 64 registers (it spilled 64-128 B of stack and cfg5 lost a third of its speed), so those get 3 / 2 CTAs per SM.
// 5 / 6 CTAs per SM (the lean instances need only 38-48 registers) were measured again with the lean loops: no gain.
#define RDF_EVAL_MIN_BLOCKS(T) RDF_EVAL_MIN_BLOCKS_T(T)
// EXACT: 0 = no node of the forest needs the exact divide (the common case: the loop has no flag test and no __fdiv_rn path),
// 1 = per-node flags, 2 = always exact (scale outside the fast domain), 3 = as 0 and every tree is complete (no early leaves: no
// "walk ended" handling in the loop either)
template <int T, int WARP_W, bool SCALE1, int EXACT>
__global__ void __launch_bounds__(256, RDF_EVAL_MIN_BLOCKS(T)) rdf_eval_packed_kernel(const __grid_constant__ rdf_eval_launch<T> q) {
    const rdf_eval_params& p = q.p;
    constexpr bool FORCE_EXACT = EXACT == 2, NEVER_EXACT = EXACT == 0 || EXACT == 3, COMPLETE = EXACT == 3;
    constexpr int WARP_H = 32 / WARP_W;
    constexpr int WARPS_X = 32 / WARP_W;
    const rdf_top_ref<T> top{q.top};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x - tile_y * p.tiles_x;
    const int x = tile_x * 32 + (warp % WARPS_X) * WARP_W + (lane % WARP_W);
    const int y = tile_y * 8 + (warp / WARPS_X) * WARP_H + (lane / WARP_W);
    if (x >= p.w || y >= p.h) return;
    const int n = p.image0 + blockIdx.y;
    const size_t li = ((size_t)n * p.h + y) * p.w + x;
    if (p.filter_class != -1 && (int)__ldg(p.filter + li) != p.filter_class) return;   // tree_eval.cu:81-85
    const uint16_t* img = p.depth + (size_t)n * p.H * p.W;
    asm volatile("" : "+l"(img));                 // keep the image base in registers (do not rematerialise it per probe)
    const int X = x * p.r, Y = y * p.r;
    const unsigned d = __ldg(img + (size_t)Y * p.W + X);
    if (d == 0u || d == RDF_NO_PIXEL) return;                                            // tree_eval.cu:88-89
    int state[T];
    rdf_walk<T, SCALE1, FORCE_EXACT, NEVER_EXACT, COMPLETE, rdf_top_ref<T>>(p.fv, img, p.W, p.H, X, Y, d, p.scale, state, top, p.top_levels);
    if (T == 1 && p.tree_mode && state[0] == RDF_NO_LEAF) return;
    const int lab = rdf_vote<T>(p.fv, state, p.probs ? p.probs + li * p.fv.C : nullptr);
    p.labels[li] = (uint16_t)lab;
}

// ---- generic path: canonical layout, any T and C, trees walked one after another ----------------------------------
// Used for rdf_eval_tree (the trainer evaluates each freshly trained tree once, src/train_model.py:105) and for
// forests with more than RDF_FAST_MAX_TREES trees.  acc lives in shared memory as [C][blockDim] (conflict-free).
struct rdf_eval_canon_params {
    const float* forest;      // [T][2^D-1][7+2C]
    const uint16_t* depth;
    const uint16_t* filter;
    uint16_t* labels;
    float* probs;
    int T, D, C;
    int W, H, w, h, r;
    int filter_class;
    int tree_mode;            // 1: evaluate_image_using_tree semantics (no write when no leaf is reached)
    int64_t num_pixels;       // N*h*w
    float scale;
};

__global__ void __launch_bounds__(128) rdf_eval_canon_kernel(const rdf_eval_canon_params p) {
    extern __shared__ float acc_s[];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= p.num_pixels) return;
    const int64_t per = (int64_t)p.h * p.w;
    const int n = (int)(i / per);
    const int rem = (int)(i - n * per);
    const int y = rem / p.w, x = rem - y * p.w;
    if (p.filter_class != -1 && (int)__ldg(p.filter + i) != p.filter_class) return;
    const uint16_t* img = p.depth + (size_t)n * p.H * p.W;
    const int X = x * p.r, Y = y * p.r;
    const unsigned d = __ldg(img + (size_t)Y * p.W + X);
    if (d == 0u || d == RDF_NO_PIXEL) return;
    const float df = (float)d;
    const int E = 7 + 2 * p.C;
    const int64_t tree_stride = (((int64_t)1 << p.D) - 1) * E;
    float* acc = acc_s + threadIdx.x;
    for (int c = 0; c < p.C; c++) acc[c * blockDim.x] = 0.f;
    bool any_leaf = false;
    for (int t = 0; t < p.T; t++) {
        const float* tree = p.forest + t * tree_stride;
        int64_t row = 0;
        for (int j = 0; j < p.D; j++) {
            const float* nd = tree + row * E;
            const float f = rdf_feature<true>(img, p.W, p.H, X, Y, df, 0.f, __fmul_rn(p.scale, __ldg(nd + 0)),
                                              __fmul_rn(p.scale, __ldg(nd + 1)), __fmul_rn(p.scale, __ldg(nd + 2)),
                                              __fmul_rn(p.scale, __ldg(nd + 3)));
            const int side = (f < __ldg(nd + 4)) ? 0 : 1;
            if (__float2int_rd(__ldg(nd + 5 + side)) == -1) {
                row = 2 * row + 1 + side;
            } else {
                const float* pdf = nd + 7 + side * p.C;
                for (int c = 0; c < p.C; c++) acc[c * blockDim.x] = __fadd_rn(acc[c * blockDim.x], __ldg(pdf + c));
                any_leaf = true;
                break;
            }
        }
    }
    if (p.tree_mode && !any_leaf) return;                                                 // tree_eval.cu:174-210
    float best = 0.f;
    int lab = 0;
    for (int c = 0; c < p.C; c++) {
        const float v = acc[c * blockDim.x];
        if (v > best) { best = v; lab = c; }
        if (p.probs) p.probs[i * p.C + c] = __fdiv_rn(v, (float)p.T);
    }
    p.labels[i] = (uint16_t)lab;
}

static int rdf_launch_canon(const rdf_eval_canon_params& p, cudaStream_t stream) {
    const int threads = 128;
    const size_t smem = (size_t)p.C * threads * sizeof(float);
    RDF_ENSURE_DYN_SMEM(rdf_eval_canon_kernel, RDF_MAX_CLASSES * threads * sizeof(float));
    const int64_t blocks = (p.num_pixels + threads - 1) / threads;
    RDF_REQUIRE(blocks <= 0x7fffffffLL, "too many pixels for one launch: %lld", (long long)p.num_pixels);
    rdf_eval_canon_kernel<<<(unsigned)blocks, threads, smem, stream>>>(p);
    RDF_LAUNCH_CHECK("rdf_eval_canon_kernel");
    return RDF_OK;
}

// warp footprint: 16x2 patches by default (8x4 / 16x2 / 32x1 measured within 2 % of each other on cfg3: the probe gathers
// are scattered by path divergence, not by the patch shape), overridable for experiments: RDF_WARP_W in {8,16,32}
static int rdf_warp_w() {
    static int v = -1;
    if (v < 0) {
        const char* e = RDF_GETENV_ONCE("RDF_WARP_W");
        v = e ? atoi(e) : 16;
        if (v != 8 && v != 16 && v != 32) v = 16;
    }
    return v;
}

template <int T, int WARP_W>
static void rdf_launch_packed_tw(const rdf_eval_params& p, const rdf_forest* forest, dim3 grid, cudaStream_t stream) {
    const bool has_exact_nodes = forest->has_exact_nodes != 0, complete = forest->has_early_leaves == 0;
    rdf_eval_launch<T> q;
    q.p = p;
    memcpy(q.top, forest->top_host, sizeof(rdf_node_hdr) * T * ((1u << forest->top_levels) - 1u));
    // fast path (reciprocal divide + magic-number floor): |scale| in [2^-30, 1] and coordinates below 2^16
    if (!rdf_scale_fastfloor_ok(p.scale) || p.W > 65535 || p.H > 65535)
        rdf_eval_packed_kernel<T, WARP_W, false, 2><<<grid, 256, 0, stream>>>(q);
    else if (p.scale == 1.f) {
        if (has_exact_nodes) rdf_eval_packed_kernel<T, WARP_W, true, 1><<<grid, 256, 0, stream>>>(q);
        else if (complete) rdf_eval_packed_kernel<T, WARP_W, true, 3><<<grid, 256, 0, stream>>>(q);
        else rdf_eval_packed_kernel<T, WARP_W, true, 0><<<grid, 256, 0, stream>>>(q);
    } else {
        if (has_exact_nodes) rdf_eval_packed_kernel<T, WARP_W, false, 1><<<grid, 256, 0, stream>>>(q);
        else if (complete) rdf_eval_packed_kernel<T, WARP_W, false, 3><<<grid, 256, 0, stream>>>(q);
        else rdf_eval_packed_kernel<T, WARP_W, false, 0><<<grid, 256, 0, stream>>>(q);
    }
}

template <int T>
static void rdf_launch_packed_t(const rdf_eval_params& p, const rdf_forest* forest, dim3 grid, cudaStream_t stream) {
    switch (rdf_warp_w()) {
        case 32: rdf_launch_packed_tw<T, 32>(p, forest, grid, stream); break;
        case 16: rdf_launch_packed_tw<T, 16>(p, forest, grid, stream); break;
        default: rdf_launch_packed_tw<T, 8>(p, forest, grid, stream); break;
    }
}

static int rdf_eval_packed(const rdf_forest_t* forest, const uint16_t* depth_dev, int num_images, int dim_x, int dim_y,
                           const uint16_t* filter_dev, int filter_class, uint16_t* labels_dev, float* probs_dev,
                           int labels_reduce, float scale, int tree_mode, void* stream) {
    RDF_REQUIRE(forest && depth_dev && labels_dev, "rdf_eval_forest: NULL argument");
    RDF_REQUIRE(forest->device == rdf_current_device(), "rdf_eval_forest: the forest was packed on device %d, current device is %d",
                forest->device, rdf_current_device());
    RDF_REQUIRE(num_images >= 0 && dim_x > 0 && dim_y > 0 && labels_reduce >= 1, "rdf_eval_forest: bad shape N=%d W=%d H=%d r=%d",
                num_images, dim_x, dim_y, labels_reduce);
    RDF_REQUIRE((int64_t)dim_x * dim_y < ((int64_t)1 << 31), "rdf_eval_forest: image of %dx%d pixels is too large", dim_x, dim_y);
    const int w = dim_x / labels_reduce, h = dim_y / labels_reduce;
    if (num_images == 0 || w == 0 || h == 0) return RDF_OK;
    if (!filter_dev) filter_class = -1;
    cudaStream_t st = rdf_stream(stream);
    if (forest->T > RDF_FAST_MAX_TREES) {
        rdf_set_error("rdf_eval_forest: forests with more than %d trees need the canonical array (use rdf_eval_forest_canonical)",
                      RDF_FAST_MAX_TREES);
        return RDF_ERR_UNSUPPORTED;
    }
    rdf_eval_params p;
    p.fv = rdf_view(forest);
    p.depth = depth_dev;
    p.filter = filter_dev;
    p.labels = labels_dev;
    p.probs = probs_dev;
    p.W = dim_x; p.H = dim_y; p.w = w; p.h = h; p.r = labels_reduce;
    p.tiles_x = (w + 31) / 32;
    const int tiles_y = (h + 7) / 8;
    p.filter_class = filter_class;
    p.scale = scale;
    p.tree_mode = tree_mode;
    p.top_levels = forest->top_levels;
    for (int n0 = 0; n0 < num_images; n0 += 65535) {
        const int nb = num_images - n0 < 65535 ? num_images - n0 : 65535;
        p.image0 = n0;
        dim3 grid((unsigned)(p.tiles_x * tiles_y), (unsigned)nb);
        switch (forest->T) {
            case 1: rdf_launch_packed_t<1>(p, forest, grid, st); break;
            case 2: rdf_launch_packed_t<2>(p, forest, grid, st); break;
            case 3: rdf_launch_packed_t<3>(p, forest, grid, st); break;
            case 4: rdf_launch_packed_t<4>(p, forest, grid, st); break;
            case 5: rdf_launch_packed_t<5>(p, forest, grid, st); break;
            case 6: rdf_launch_packed_t<6>(p, forest, grid, st); break;
            case 7: rdf_launch_packed_t<7>(p, forest, grid, st); break;
            default: rdf_launch_packed_t<8>(p, forest, grid, st); break;
        }
        RDF_LAUNCH_CHECK("rdf_eval_packed_kernel");
    }
    return RDF_OK;
}

extern "C" int rdf_eval_forest(const rdf_forest_t* forest, const uint16_t* depth_dev, int num_images, int dim_x, int dim_y,
                               const uint16_t* filter_dev, int filter_class, uint16_t* labels_dev, float* probs_dev,
                               int labels_reduce, float scale, void* stream) {
    return rdf_eval_packed(forest, depth_dev, num_images, dim_x, dim_y, filter_dev, filter_class, labels_dev, probs_dev, labels_reduce,
                           scale, 0, stream);
}

// evaluate_image_using_tree over a packed one-tree handle (the fast path; rdf_eval_tree reads the canonical array)
extern "C" int rdf_eval_tree_packed(const rdf_forest_t* tree, const uint16_t* depth_dev, int num_images, int dim_x, int dim_y,
                                    uint16_t* labels_dev, void* stream) {
    RDF_REQUIRE(tree != nullptr && tree->T == 1, "rdf_eval_tree_packed: the handle must hold exactly one tree");
    return rdf_eval_packed(tree, depth_dev, num_images, dim_x, dim_y, nullptr, -1, labels_dev, nullptr, 1, 1.f, 1, stream);
}

// Same contract as rdf_eval_forest but reads the canonical array directly (no handle, any tree count).
extern "C" int rdf_eval_forest_canonical(const float* forest_dev, int num_trees, int max_depth, int num_classes,
                                         const uint16_t* depth_dev, int num_images, int dim_x, int dim_y,
                                         const uint16_t* filter_dev, int filter_class, uint16_t* labels_dev,
                                         float* probs_dev, int labels_reduce, float scale, void* stream) {
    RDF_REQUIRE(forest_dev && depth_dev && labels_dev, "rdf_eval_forest_canonical: NULL argument");
    RDF_REQUIRE(num_trees >= 1 && max_depth >= 1 && max_depth <= RDF_MAX_DEPTH && num_classes >= 1 && num_classes <= RDF_MAX_CLASSES,
                "rdf_eval_forest_canonical: bad forest shape T=%d D=%d C=%d", num_trees, max_depth, num_classes);
    RDF_REQUIRE(num_images >= 0 && dim_x > 0 && dim_y > 0 && labels_reduce >= 1, "rdf_eval_forest_canonical: bad image shape");
    rdf_eval_canon_params p;
    p.forest = forest_dev;
    p.depth = depth_dev;
    p.filter = filter_dev;
    p.labels = labels_dev;
    p.probs = probs_dev;
    p.T = num_trees; p.D = max_depth; p.C = num_classes;
    p.W = dim_x; p.H = dim_y; p.r = labels_reduce;
    p.w = dim_x / labels_reduce; p.h = dim_y / labels_reduce;
    p.filter_class = filter_dev ? filter_class : -1;
    p.tree_mode = 0;
    p.num_pixels = (int64_t)num_images * p.w * p.h;
    p.scale = scale;
    if (p.num_pixels == 0) return RDF_OK;
    return rdf_launch_canon(p, rdf_stream(stream));
}

extern "C" int rdf_eval_tree(const float* tree_dev, int max_depth, int num_classes, const uint16_t* depth_dev, int num_images,
                             int dim_x, int dim_y, uint16_t* labels_dev, void* stream) {
    RDF_REQUIRE(tree_dev && depth_dev && labels_dev, "rdf_eval_tree: NULL argument");
    RDF_REQUIRE(max_depth >= 1 && max_depth <= RDF_MAX_DEPTH && num_classes >= 1 && num_classes <= RDF_MAX_CLASSES,
                "rdf_eval_tree: bad tree shape D=%d C=%d", max_depth, num_classes);
    RDF_REQUIRE(num_images >= 0 && dim_x > 0 && dim_y > 0, "rdf_eval_tree: bad image shape");
    rdf_eval_canon_params p;
    p.forest = tree_dev;
    p.depth = depth_dev;
    p.filter = nullptr;
    p.labels = labels_dev;
    p.probs = nullptr;
    p.T = 1; p.D = max_depth; p.C = num_classes;
    p.W = dim_x; p.H = dim_y; p.w = dim_x; p.h = dim_y; p.r = 1;
    p.filter_class = -1;
    p.tree_mode = 1;
    p.num_pixels = (int64_t)num_images * dim_x * dim_y;
    p.scale = 1.f;
    if (p.num_pixels == 0) return RDF_OK;
    return rdf_launch_canon(p, rdf_stream(stream));
}
