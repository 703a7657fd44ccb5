// Per-pixel forest traversal over the packed layout, shared by rdf_eval.cu and rdf_layered.cu.
//
// One thread owns one labels pixel and walks all T trees of the forest *interleaved*: every level issues the T
// node-header loads back to back, then the 2T depth probes, then the T compares.  The T walks are independent
// dependency chains, so a single thread keeps T header loads / 2T probes in flight (memory-level parallelism
// without needing T times the warps), and the summation order of leaf pdfs stays tree 0..T-1 (SURVEY note N1).
#pragma once
#include "rdf_common.cuh"

#define RDF_FAST_MAX_TREES 8

struct rdf_forest_view {
    const rdf_node_hdr* hdr;
    const float* pdf;
    int64_t nodes_per_tree;
    int T, D, C, CP;
};

static inline rdf_forest_view rdf_view(const rdf_forest* f) {
    rdf_forest_view v;
    v.hdr = f->hdr;
    v.pdf = f->pdf;
    v.nodes_per_tree = f->nodes_per_tree;
    v.T = f->T;
    v.D = f->D;
    v.C = f->C;
    v.CP = f->CP;
    return v;
}

// Walk T trees from the root.  leaf[t] = 2*row + side of the reached leaf, or -1 if the walk fell off level D-1
// with a "continue" flag (adds nothing, src/cuda/tree_eval.cu:95-128).
// SCALE1: scale == 1.0f, so scale*u == u exactly and the multiplies are dropped.  FORCE_EXACT: always use __fdiv_rn
// (scale outside the fast-divide domain); otherwise nodes flagged RDF_FLAG_EXACT_DIV take the exact path per level.
template <int T, bool SCALE1, bool FORCE_EXACT>
__device__ __forceinline__ void rdf_walk(const rdf_forest_view& fv, const uint16_t* __restrict__ img, int W, int H, int X,
                                         int Y, float df, float scale, int (&leaf)[T]) {
    int row[T];
#pragma unroll
    for (int t = 0; t < T; t++) {
        row[t] = 0;
        leaf[t] = -1;
    }
    const float rcp = __frcp_rn(df);                                 // RN(1/d), once per pixel
    unsigned alive = (1u << T) - 1u;
    for (int j = 0; j < fv.D && alive; j++) {
        float4 a[T];
        float th[T];
        int fl[T];
        int any_flags = 0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            // dead walks re-read their last node: harmless, keeps the loop branch-free
            const float4* p = reinterpret_cast<const float4*>(fv.hdr + (int64_t)t * fv.nodes_per_tree + row[t]);
            a[t] = __ldg(p);
            const float2 b = __ldg(reinterpret_cast<const float2*>(p + 1));
            th[t] = b.x;
            fl[t] = __float_as_int(b.y);
            any_flags |= fl[t];
            if (!SCALE1) {
                a[t].x = __fmul_rn(scale, a[t].x);
                a[t].y = __fmul_rn(scale, a[t].y);
                a[t].z = __fmul_rn(scale, a[t].z);
                a[t].w = __fmul_rn(scale, a[t].w);
            }
        }
        float f[T];
        if (FORCE_EXACT || (any_flags & RDF_FLAG_EXACT_DIV)) {
#pragma unroll
            for (int t = 0; t < T; t++) f[t] = rdf_feature<true>(img, W, H, X, Y, df, rcp, a[t].x, a[t].y, a[t].z, a[t].w);
        } else {
#pragma unroll
            for (int t = 0; t < T; t++) f[t] = rdf_feature<false>(img, W, H, X, Y, df, rcp, a[t].x, a[t].y, a[t].z, a[t].w);
        }
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int side = (f[t] < th[t]) ? 0 : 1;                // NaN threshold -> right (tree_eval.cu:106)
            const bool is_alive = (alive >> t) & 1u;
            const bool cont = (fl[t] >> side) & 1;
            if (is_alive) {
                if (cont) {
                    row[t] = 2 * row[t] + 1 + side;                  // (2^(j+1)-1) + 2g + side
                } else {
                    leaf[t] = 2 * row[t] + side;
                    alive &= ~(1u << t);
                }
            }
        }
    }
}

// get_best_pdf_chance over the tree-ordered sum of the reached leaf pdfs (src/cuda/tree_eval.cu:7-21,125).
// probs (optional): receives sum / T per class.
template <int T>
__device__ __forceinline__ int rdf_vote(const rdf_forest_view& fv, const int (&leaf)[T], float* __restrict__ probs) {
    float best = 0.f;
    int lab = 0;
    const float inv_t = (float)fv.T;
    for (int c = 0; c < fv.CP; c += 4) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < T; t++) {
            if (leaf[t] >= 0) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(
                    fv.pdf + ((int64_t)t * fv.nodes_per_tree * 2 + leaf[t]) * fv.CP + c));
                s.x = __fadd_rn(s.x, v.x);
                s.y = __fadd_rn(s.y, v.y);
                s.z = __fadd_rn(s.z, v.z);
                s.w = __fadd_rn(s.w, v.w);
            }
        }
        if (s.x > best) { best = s.x; lab = c; }
        if (s.y > best) { best = s.y; lab = c + 1; }
        if (s.z > best) { best = s.z; lab = c + 2; }
        if (s.w > best) { best = s.w; lab = c + 3; }
        if (probs) {
            if (c + 0 < fv.C) probs[c + 0] = __fdiv_rn(s.x, inv_t);
            if (c + 1 < fv.C) probs[c + 1] = __fdiv_rn(s.y, inv_t);
            if (c + 2 < fv.C) probs[c + 2] = __fdiv_rn(s.z, inv_t);
            if (c + 3 < fv.C) probs[c + 3] = __fdiv_rn(s.w, inv_t);
        }
    }
    return lab;
}

// Evaluate one forest at one pixel (T <= RDF_FAST_MAX_TREES, dispatched on the runtime tree count).
template <bool SCALE1, bool FORCE_EXACT>
__device__ __forceinline__ int rdf_eval_pixel(const rdf_forest_view& fv, const uint16_t* __restrict__ img, int W, int H,
                                              int X, int Y, float df, float scale, float* __restrict__ probs) {
#define RDF_CASE(TT)                                                             \
    case TT: {                                                                   \
        int leaf[TT];                                                            \
        rdf_walk<TT, SCALE1, FORCE_EXACT>(fv, img, W, H, X, Y, df, scale, leaf); \
        return rdf_vote<TT>(fv, leaf, probs);                                    \
    }
    switch (fv.T) {
        RDF_CASE(1)
        RDF_CASE(2)
        RDF_CASE(3)
        RDF_CASE(4)
        RDF_CASE(5)
        RDF_CASE(6)
        RDF_CASE(7)
        RDF_CASE(8)
    }
#undef RDF_CASE
    return 0;
}
