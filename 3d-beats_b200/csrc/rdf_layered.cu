// Stacked ("layered") forests: replaces LayeredDecisionForest.run (reference src/decision_tree.py:233-264) and
// make_composite_labels_image (src/cuda/tree_eval.cu:214-248).
#include <stdlib.h>
#include <string.h>

#include "rdf_traverse.cuh"

// ---- stand-alone composite (API parity with DecisionTreeEvaluator.make_composite_labels_image) -----------------
__global__ void __launch_bounds__(256) rdf_composite_kernel(const uint16_t* const* __restrict__ label_images, int L, int w,
                                                            int h, const int2* __restrict__ cond,
                                                            uint16_t* __restrict__ composite) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    int off = 0;
    for (int k = 0; k < L; k++) {
        const unsigned l = __ldg(label_images[k] + i);
        if (l == 0u || l == RDF_NO_PIXEL) return;                  // composite left untouched (tree_eval.cu:235)
        const int2 tv = __ldg(cond + off + (int)l - 1);
        if (tv.x == 0) {
            composite[i] = (uint16_t)tv.y;
            return;
        }
        off = tv.y;
    }
    // the reference asserts here (tree_eval.cu:246-247); a malformed table simply leaves the pixel untouched
}

extern "C" int rdf_composite(const uint16_t* const* label_images_dev, int num_label_images, int dim_x, int dim_y,
                             const int32_t* conditions_dev, uint16_t* composite_dev, void* stream) {
    RDF_REQUIRE(label_images_dev && conditions_dev && composite_dev, "rdf_composite: NULL argument");
    RDF_REQUIRE(num_label_images >= 1 && dim_x > 0 && dim_y > 0, "rdf_composite: bad shape");
    const int n = dim_x * dim_y;
    rdf_composite_kernel<<<(n + 255) / 256, 256, 0, rdf_stream(stream)>>>(label_images_dev, num_label_images, dim_x, dim_y,
                                                                           reinterpret_cast<const int2*>(conditions_dev),
                                                                           composite_dev);
    RDF_LAUNCH_CHECK("rdf_composite_kernel");
    return RDF_OK;
}

// ---- fused layered run ---------------------------------------------------------------------------------------
// One thread per labels pixel evaluates every layer in turn; the gate of layer i (label of layer filter_model[i])
// is read from a register, the composite walk happens on the same registers, and every output pixel is written
// exactly once (65535 where the reference's pre-fill would survive).  No fills, no intermediate global reads.
struct rdf_layered_params {
    rdf_forest_view fv[RDF_MAX_LAYERS];
    int filter_model[RDF_MAX_LAYERS];
    int filter_class[RDF_MAX_LAYERS];
    uint16_t* layer_labels[RDF_MAX_LAYERS];
    const uint16_t* depth;
    const int2* cond;
    uint16_t* composite;
    int L, n_cond;
    int W, H, w, h, r;
    int tiles_x;
    float scale;
    unsigned flip_mask;        // bit n: composite of image n is written mirrored in x (the left hand of the live product,
                               // src/3d_bz.py:439-446)
    size_t depth_stride, label_stride;   // elements between consecutive images of a batch (blockIdx.y)
};

template <int WARP_W, bool SCALE1, bool FORCE_EXACT>
__global__ void __launch_bounds__(256) rdf_layered_kernel(const __grid_constant__ rdf_layered_params p) {
    constexpr int WARP_H = 32 / WARP_W;
    constexpr int WARPS_X = 32 / WARP_W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x - tile_y * p.tiles_x;
    const int x = tile_x * 32 + (warp % WARPS_X) * WARP_W + (lane % WARP_W);
    const int y = tile_y * 8 + (warp / WARPS_X) * WARP_H + (lane / WARP_W);
    if (x >= p.w || y >= p.h) return;
    const size_t lbase = (size_t)blockIdx.y * p.label_stride;
    const size_t li = lbase + (size_t)y * p.w + x;
    const uint16_t* __restrict__ depth = p.depth + (size_t)blockIdx.y * p.depth_stride;
    const int X = x * p.r, Y = y * p.r;
    const unsigned d = __ldg(depth + (size_t)Y * p.W + X);
    const bool valid = !(d == 0u || d == RDF_NO_PIXEL);

    // per-layer labels of this pixel, 16 bits each, packed so the layer loop can stay rolled (RDF_MAX_LAYERS == 8)
    unsigned long long lab_lo = ~0ull, lab_hi = ~0ull;              // all 65535 = the reference's pre-fill
    auto get_lab = [&](int i) -> unsigned {
        const unsigned long long wd = i < 4 ? lab_lo : lab_hi;
        return (unsigned)(wd >> (16 * (i & 3))) & 0xffffu;
    };
    auto set_lab = [&](int i, unsigned v) {
        const unsigned long long m = 0xffffull << (16 * (i & 3));
        const unsigned long long b = (unsigned long long)(v & 0xffffu) << (16 * (i & 3));
        if (i < 4) lab_lo = (lab_lo & ~m) | b; else lab_hi = (lab_hi & ~m) | b;
    };

#pragma unroll 1
    for (int i = 0; i < p.L; i++) {
        bool run = valid;
        const int fm = p.filter_model[i];
        // layers not yet evaluated still hold the 65535 pre-fill, as in the reference
        if (fm >= 0 && p.filter_class[i] != -1) run = run && ((int)get_lab(fm) == p.filter_class[i]);
        unsigned l = RDF_NO_PIXEL;
        if (run) {
            l = (unsigned)rdf_eval_pixel<SCALE1, FORCE_EXACT>(p.fv[i], depth, p.W, p.H, X, Y, d, p.scale, nullptr);
            set_lab(i, l);
        }
        p.layer_labels[i][li] = (uint16_t)l;
    }

    // composite walk (tree_eval.cu:232-244) on registers
    unsigned comp = RDF_NO_PIXEL;
    int off = 0;
#pragma unroll 1
    for (int i = 0; i < p.L; i++) {
        const unsigned l = get_lab(i);
        if (l == 0u || l == RDF_NO_PIXEL) break;
        const int idx = off + (int)l - 1;
        if (idx < 0 || idx >= p.n_cond) break;                     // malformed table: leave 65535
        const int2 tv = __ldg(p.cond + idx);
        if (tv.x == 0) {
            comp = (unsigned)tv.y & 0xffffu;
            break;
        }
        off = tv.y;
    }
    p.composite[((p.flip_mask >> blockIdx.y) & 1u) ? lbase + (size_t)y * p.w + (p.w - 1 - x) : li] = (uint16_t)comp;
}

// ---- latency path: one thread per (pixel, tree walk) ---------------------------------------------------------------
// The live product evaluates ONE frame with few valid pixels (a hand blob), so the fused kernel above is bound by the
// dependent-load chain of a single thread (L layers x D levels, two L2 round trips per level).  Here every tree of
// every layer is a separate walk running in its own warp (lanes = 32 neighbouring pixels, warp = one tree), layers
// are evaluated speculatively in parallel and gated afterwards, and each level prefetches BOTH child headers while
// the depth probes of the current node are in flight, so a level costs about one L2 round trip instead of two.
// Results are identical: gating only decides which labels are kept.
#define RL2_MAX_WALKS 32
#define RL2_MAX_SUB 4

struct rdf_layered2_params {
    rdf_layered_params base;
    int walk_layer[RL2_MAX_WALKS];
    int walk_tree[RL2_MAX_WALKS];
    int first_walk[RDF_MAX_LAYERS];
    int num_walks;
    int max_cp;                // largest padded class count of the layers (uniform trip count of the vote loop)
};

template <bool SCALE1, bool FORCE_EXACT>
__global__ void __launch_bounds__(1024) rdf_layered_walks_kernel(const __grid_constant__ rdf_layered2_params q) {
    const rdf_layered_params& p = q.base;
    // let a dependent grid (the mean shift that follows in the live pipeline) be scheduled now; it waits for this grid's
    // completion itself (griddepcontrol.wait) before reading the label images
    asm volatile("griddepcontrol.launch_dependents;");
    // ... and this grid may itself have been scheduled early behind the kernel that produces the depth frame
    // (rdf_upload_frame in the live pipeline): nothing of the frame is read before that kernel has completed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // blockDim = (32 lanes, walks, S sub-tiles): S patches of 8x4 pixels per CTA (S x walks <= 32 warps).  On a live frame nine tiles
    // out of ten hold no valid pixel and cost a CTA launch each (6.5 us for the 3180 empty CTAs of a 424x240 label image), but S > 1
    // measured slower (see the launch code): the default is S = 1.
    __shared__ unsigned short lab_s[RL2_MAX_SUB][RDF_MAX_LAYERS][32];
    const int lane = threadIdx.x, walk = threadIdx.y, sub = threadIdx.z;
    const int stile_y = blockIdx.x / p.tiles_x, stile_x = blockIdx.x - stile_y * p.tiles_x;
    const int sub_w = blockDim.z >= 2 ? 2 : 1;                               // sub-tiles side by side
    const int tile_x = stile_x * sub_w + (sub & (sub_w - 1)), tile_y = stile_y * (blockDim.z / sub_w) + sub / sub_w;
    const int x = tile_x * 8 + (lane & 7), y = tile_y * 4 + (lane >> 3);     // warp = 8x4 patch of labels pixels
    const bool inside = x < p.w && y < p.h;
    const int X = x * p.r, Y = y * p.r;
    const uint16_t* __restrict__ depth = p.depth + (size_t)blockIdx.y * p.depth_stride;
    unsigned d = RDF_NO_PIXEL;
    if (inside) d = __ldg(depth + (size_t)Y * p.W + X);
    // the root header of this warp's tree does not depend on the pixel: requested together with the centre depth (the same few
    // nodes for every CTA, so an empty tile pays an L1 hit for it)
    const int layer = q.walk_layer[walk], t = q.walk_tree[walk];
    const rdf_forest_view& fv = p.fv[layer];
    const rdf_hdr_regs root = rdf_load_hdr(fv.hdr, t * fv.nodes_per_tree);
    const bool valid = inside && d != 0u && d != RDF_NO_PIXEL;
    const size_t lbase = (size_t)blockIdx.y * p.label_stride;
    const size_t li = lbase + (size_t)y * p.w + x;
    const size_t lo = ((p.flip_mask >> blockIdx.y) & 1u) ? lbase + (size_t)y * p.w + (p.w - 1 - x) : li;   // composite label goes here
    if (__syncthreads_or(valid) == 0) {                                   // nothing to evaluate in this super-tile: pre-fill only
        if (walk == 0 && inside) {
            for (int i = 0; i < p.L; i++) p.layer_labels[i][li] = (uint16_t)RDF_NO_PIXEL;
            p.composite[lo] = (uint16_t)RDF_NO_PIXEL;
        }
        return;
    }
    const int wbase = sub * q.num_walks;                                     // this patch's rows of pdf_s
    int leaf = RDF_NO_LEAF;                                                  // ~leaf_id once the walk has ended
    if (valid) {
        // loop invariants pinned in registers: left to itself the compiler re-derives them every level from the parameter block
        // (a chain of dependent constant-bank loads indexed by threadIdx.y / blockIdx.y in front of every header and probe address)
        unsigned long long hdr_u = (unsigned long long)fv.hdr, depth_u = (unsigned long long)depth;
        int levels = fv.D, imgW = p.W, imgH = p.H;
        asm volatile("" : "+l"(hdr_u), "+l"(depth_u), "+r"(levels), "+r"(imgW), "+r"(imgH));
        const rdf_node_hdr* hdrp = reinterpret_cast<const rdf_node_hdr*>(hdr_u);
        const uint16_t* __restrict__ img = reinterpret_cast<const uint16_t*>(depth_u);
        const float df = (float)d;
        const float rcp = __frcp_rn(df);
        const float xm = (float)X + RDF_MAGIC_F, ym = (float)Y + RDF_MAGIC_F;
        rdf_hdr_regs h = root;
        for (int j = 0; j < levels; j++) {
            // both children (when they are nodes) are requested while the probes of this node are in flight
            // (a leaf side loads node 0 instead - never used, the walk ends there - which keeps the loads unpredicated and spares
            // the sixteen register moves of a default value)
            const rdf_hdr_regs hl = rdf_load_hdr(hdrp, max(h.b.y, 0));
            const rdf_hdr_regs hr = rdf_load_hdr(hdrp, max(h.b.z, 0));
            float sx = h.a.x, sy = h.a.y, sz = h.a.z, sw = h.a.w;
            if (!SCALE1) {
                sx = __fmul_rn(p.scale, sx); sy = __fmul_rn(p.scale, sy);
                sz = __fmul_rn(p.scale, sz); sw = __fmul_rn(p.scale, sw);
            }
            int f;
            if (FORCE_EXACT || (h.b.w & RDF_FLAG_EXACT_DIV)) f = rdf_feature_i<true>(img, imgW, imgH, X, Y, df, rcp, xm, ym, sx, sy, sz, sw);
            else f = rdf_feature_i<false>(img, imgW, imgH, X, Y, df, rcp, xm, ym, sx, sy, sz, sw);
            const bool go_left = f < h.b.x;
            const int next = go_left ? h.b.y : h.b.z;
            if (next < 0) {
                leaf = next;
                break;
            }
            h = go_left ? hl : hr;
        }
    }
    // ---- vote of each layer: leaf pdfs summed in tree order by the thread that owns the layer's first walk (speculative: gating is
    // applied below).  Every walk loads ITS leaf's pdf row, one 16-byte chunk per pass, and hands it over through shared memory;
    // the chunk of the next pass is requested before the barrier of this one, so all rows cost about one L2 round trip in total.
    // (One thread fetching T x CP/4 chunks in a loop paid a full round trip per chunk: 9 in a row for 3 trees x 11 classes.)
    extern __shared__ float4 pdf_s[];                                        // [2][blockDim.z * num_walks][32]
    const int nrow = blockDim.z * q.num_walks;
    const bool has_leaf = valid && leaf != RDF_NO_LEAF;
    const int myCP = fv.CP, myT = fv.T;
    const float* myrow = fv.pdf + (size_t)(unsigned)(~leaf) * myCP;          // dereferenced only when has_leaf
    const bool voter = walk == q.first_walk[layer] && valid;
    auto load_chunk = [&](int c) -> float4 {
        return (has_leaf && c < myCP) ? __ldg(reinterpret_cast<const float4*>(myrow + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 cur = load_chunk(0);
    float best = 0.f;
    int lab = 0;
    for (int c = 0, k = 0; c < q.max_cp; c += 4, k ^= 1) {                   // q.max_cp: block-uniform, every warp meets every barrier
        pdf_s[(k * nrow + wbase + walk) * 32 + lane] = cur;
        if (c + 4 < q.max_cp) cur = load_chunk(c + 4);
        __syncthreads();                                                     // double buffer: one barrier per pass is enough
        if (voter && c < myCP) {
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int tt = 0; tt < myT; tt++) {                               // a walk without leaf contributed zeros: x + 0 == x
                const float4 v = pdf_s[(k * nrow + wbase + walk + tt) * 32 + lane];
                sum.x = __fadd_rn(sum.x, v.x); sum.y = __fadd_rn(sum.y, v.y); sum.z = __fadd_rn(sum.z, v.z); sum.w = __fadd_rn(sum.w, v.w);
            }
            if (sum.x > best) { best = sum.x; lab = c; }
            if (sum.y > best) { best = sum.y; lab = c + 1; }
            if (sum.z > best) { best = sum.z; lab = c + 2; }
            if (sum.w > best) { best = sum.w; lab = c + 3; }
        }
    }
    if (voter) lab_s[sub][layer][lane] = (unsigned short)lab;
    __syncthreads();
    if (walk != 0 || !inside) return;
    // gating in layer order + composite walk (tree_eval.cu:232-244), identical to the fused kernel
    unsigned long long lab_lo = ~0ull, lab_hi = ~0ull;
    auto get_lab = [&](int i) -> unsigned {
        const unsigned long long wd = i < 4 ? lab_lo : lab_hi;
        return (unsigned)(wd >> (16 * (i & 3))) & 0xffffu;
    };
    auto set_lab = [&](int i, unsigned v) {
        const unsigned long long m = 0xffffull << (16 * (i & 3));
        const unsigned long long bb = (unsigned long long)(v & 0xffffu) << (16 * (i & 3));
        if (i < 4) lab_lo = (lab_lo & ~m) | bb; else lab_hi = (lab_hi & ~m) | bb;
    };
#pragma unroll 1
    for (int i = 0; i < p.L; i++) {
        bool run = valid;
        const int fm = p.filter_model[i];
        if (fm >= 0 && p.filter_class[i] != -1) run = run && ((int)get_lab(fm) == p.filter_class[i]);
        unsigned l = RDF_NO_PIXEL;
        if (run) {
            l = lab_s[sub][i][lane];
            set_lab(i, l);
        }
        p.layer_labels[i][li] = (uint16_t)l;
    }
    unsigned comp = RDF_NO_PIXEL;
    int off = 0;
#pragma unroll 1
    for (int i = 0; i < p.L; i++) {
        const unsigned l = get_lab(i);
        if (l == 0u || l == RDF_NO_PIXEL) break;
        const int idx = off + (int)l - 1;
        if (idx < 0 || idx >= p.n_cond) break;
        const int2 tv = __ldg(p.cond + idx);
        if (tv.x == 0) {
            comp = (unsigned)tv.y & 0xffffu;
            break;
        }
        off = tv.y;
    }
    p.composite[lo] = (uint16_t)comp;
}

static int rdf_layered_run_impl(const rdf_forest_t* const* forests, int num_layers, const int* filter_model,
                               const int* filter_class, const uint16_t* depth_dev, int dim_x, int dim_y,
                               uint16_t* const* labels_per_layer, const int32_t* conditions_dev, int n_cond,
                               uint16_t* composite_dev, int labels_reduce, float scale, int num_images, unsigned flip_mask,
                               void* stream) {
    RDF_REQUIRE(forests && filter_model && filter_class && depth_dev && labels_per_layer && conditions_dev && composite_dev,
                "rdf_layered_run: NULL argument");
    RDF_REQUIRE(num_images >= 1 && num_images <= 32, "rdf_layered_run: num_images=%d outside 1..32", num_images);
    RDF_REQUIRE(num_layers >= 1 && num_layers <= RDF_MAX_LAYERS, "rdf_layered_run: num_layers=%d outside 1..%d", num_layers,
                RDF_MAX_LAYERS);
    RDF_REQUIRE(dim_x > 0 && dim_y > 0 && labels_reduce >= 1 && n_cond >= 1, "rdf_layered_run: bad shape");
    rdf_layered_params p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < num_layers; i++) {
        RDF_REQUIRE(forests[i] != nullptr && labels_per_layer[i] != nullptr, "rdf_layered_run: layer %d is NULL", i);
        RDF_REQUIRE(forests[i]->device == rdf_current_device(), "rdf_layered_run: layer %d was packed on device %d, current device is %d",
                    i, forests[i]->device, rdf_current_device());
        if (forests[i]->T > RDF_FAST_MAX_TREES) {
            rdf_set_error("rdf_layered_run: layer %d has %d trees (max %d in the fused path)", i, forests[i]->T, RDF_FAST_MAX_TREES);
            return RDF_ERR_UNSUPPORTED;
        }
        RDF_REQUIRE(filter_model[i] < num_layers, "rdf_layered_run: filter_model[%d]=%d out of range", i, filter_model[i]);
        p.fv[i] = rdf_view(forests[i]);
        p.filter_model[i] = filter_model[i];
        p.filter_class[i] = filter_class[i];
        p.layer_labels[i] = labels_per_layer[i];
    }
    p.depth = depth_dev;
    p.cond = reinterpret_cast<const int2*>(conditions_dev);
    p.composite = composite_dev;
    p.L = num_layers;
    p.n_cond = n_cond;
    p.W = dim_x; p.H = dim_y; p.r = labels_reduce;
    p.w = dim_x / labels_reduce; p.h = dim_y / labels_reduce;
    if (p.w == 0 || p.h == 0) return RDF_OK;
    p.tiles_x = (p.w + 31) / 32;
    p.scale = scale;
    p.flip_mask = flip_mask;
    p.depth_stride = (size_t)dim_x * dim_y;
    p.label_stride = (size_t)p.w * p.h;
    const int tiles_y = (p.h + 7) / 8;
    RDF_REQUIRE((int64_t)dim_x * dim_y < ((int64_t)1 << 31), "rdf_layered_run: image of %dx%d pixels is too large", dim_x, dim_y);
    const bool fast = rdf_scale_fastfloor_ok(scale) && dim_x <= 65535 && dim_y <= 65535;   // see rdf_common.cuh
    // latency path: one warp per tree walk (see rdf_layered_walks_kernel)
    int num_walks = 0;
    for (int i = 0; i < num_layers; i++) num_walks += forests[i]->T;
    if (num_walks <= RL2_MAX_WALKS && !RDF_GETENV_ONCE("RDF_LAYERED_V1")) {
        rdf_layered2_params q;
        q.base = p;
        // 8x4 patches per CTA (S x walks <= 32 warps).  Measured (RDF_LAYERED_SUB = 1 / 2 / 4): product frame 98.0 / 98.8 / 104.3 us, cfg2
        // frame 69.6 / 70.3 / 70.6 us - fewer, fatter CTAs save launches of empty tiles but wait for their slowest patch: one patch it is
        int S = 1;
        if (const char* e = RDF_GETENV_ONCE("RDF_LAYERED_SUB")) { const int v = atoi(e); if ((v == 1 || v == 2 || v == 4) && v * num_walks <= 32) S = v; }
        const int sub_w = S >= 2 ? 2 : 1, sub_h = S / sub_w;
        q.base.tiles_x = (p.w + 8 * sub_w - 1) / (8 * sub_w);
        q.num_walks = num_walks;
        q.max_cp = 4;
        for (int i = 0; i < num_layers; i++) q.max_cp = q.max_cp > forests[i]->CP ? q.max_cp : forests[i]->CP;
        int wi = 0;
        for (int i = 0; i < num_layers; i++) {
            q.first_walk[i] = wi;
            for (int t = 0; t < forests[i]->T; t++, wi++) {
                q.walk_layer[wi] = i;
                q.walk_tree[wi] = t;
            }
        }
        const int nb = q.base.tiles_x * ((p.h + 4 * sub_h - 1) / (4 * sub_h));
        dim3 block(32, num_walks, S);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nb, num_images, 1);
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = (size_t)2 * S * num_walks * 32 * sizeof(float4);   // vote exchange, <= 32 KB
        cfg.stream = rdf_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see griddepcontrol.wait in the kernel
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = RDF_GETENV_ONCE("RDF_NO_PDL") ? 0 : 1;
        if (!fast)
            RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_layered_walks_kernel<false, true>, q));
        else if (scale == 1.f)
            RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_layered_walks_kernel<true, false>, q));
        else
            RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_layered_walks_kernel<false, false>, q));
        return RDF_OK;
    }
    const int nblk = p.tiles_x * tiles_y;
    if (!fast)
        rdf_layered_kernel<8, false, true><<<dim3(nblk, num_images, 1), 256, 0, rdf_stream(stream)>>>(p);
    else if (scale == 1.f)
        rdf_layered_kernel<8, true, false><<<dim3(nblk, num_images, 1), 256, 0, rdf_stream(stream)>>>(p);
    else
        rdf_layered_kernel<8, false, false><<<dim3(nblk, num_images, 1), 256, 0, rdf_stream(stream)>>>(p);
    RDF_LAUNCH_CHECK("rdf_layered_kernel");
    return RDF_OK;
}

extern "C" int rdf_layered_run(const rdf_forest_t* const* forests, int num_layers, const int* filter_model,
                               const int* filter_class, const uint16_t* depth_dev, int dim_x, int dim_y,
                               uint16_t* const* labels_per_layer, const int32_t* conditions_dev, int n_cond,
                               uint16_t* composite_dev, int labels_reduce, float scale, void* stream) {
    return rdf_layered_run_impl(forests, num_layers, filter_model, filter_class, depth_dev, dim_x, dim_y, labels_per_layer,
                                conditions_dev, n_cond, composite_dev, labels_reduce, scale, 1, 0u, stream);
}

extern "C" int rdf_layered_run_batch(const rdf_forest_t* const* forests, int num_layers, const int* filter_model,
                                     const int* filter_class, const uint16_t* depth_dev, int num_images, int dim_x, int dim_y,
                                     uint16_t* const* labels_per_layer, const int32_t* conditions_dev, int n_cond,
                                     uint16_t* composite_dev, int labels_reduce, float scale, unsigned composite_flip_x_mask,
                                     void* stream) {
    return rdf_layered_run_impl(forests, num_layers, filter_model, filter_class, depth_dev, dim_x, dim_y, labels_per_layer,
                                conditions_dev, n_cond, composite_dev, labels_reduce, scale, num_images, composite_flip_x_mask, stream);
}

// ---- frame upload as a kernel ------------------------------------------------------------------------------------------
// Replaces the host-to-device copy of the live frame (depth_image.cu().set(np), src/3d_bz.py:156-157, src/run_live.py:80) inside
// a per-frame CUDA graph: the SMs read the frame from pinned host memory (zero-copy over PCIe, one 16-byte load per thread,
// everything in flight at once) and write it to HBM.  As a kernel it can be chained to the layered forest with programmatic
// dependent launch, which a copy-engine node cannot (the copy -> kernel edge of the graph cost ~4 us).
__global__ void __launch_bounds__(256) rdf_upload_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16,
                                                         const unsigned char* __restrict__ src_tail, unsigned char* __restrict__ dst_tail,
                                                         int tail) {
    asm volatile("griddepcontrol.launch_dependents;");
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

extern "C" int rdf_upload_frame(const void* host_pinned, void* dev, size_t bytes, void* stream) {
    RDF_REQUIRE(host_pinned && dev, "rdf_upload_frame: NULL argument");
    RDF_REQUIRE(((uintptr_t)host_pinned & 15u) == 0 && ((uintptr_t)dev & 15u) == 0, "rdf_upload_frame: buffers must be 16-byte aligned");
    if (bytes == 0) return RDF_OK;
    const size_t n16 = bytes / 16;
    const int tail = (int)(bytes - n16 * 16);
    size_t blocks = (n16 + 255) / 256;
    if (blocks > rdf_sm_count() * 8) blocks = rdf_sm_count() * 8;
    if (blocks < 1) blocks = 1;
    rdf_upload_kernel<<<(unsigned)blocks, 256, 0, rdf_stream(stream)>>>(
        reinterpret_cast<const uint4*>(host_pinned), reinterpret_cast<uint4*>(dev), n16,
        reinterpret_cast<const unsigned char*>(host_pinned) + n16 * 16, reinterpret_cast<unsigned char*>(dev) + n16 * 16, tail);
    RDF_LAUNCH_CHECK("rdf_upload_kernel");
    return RDF_OK;
}
