// Per-pixel forest traversal over the packed layout, shared by rdf_eval.cu and rdf_layered.cu.
//
// One thread owns one labels pixel and walks all T trees of the forest *interleaved*: every level issues the T
// node-header loads back to back, then the 2T depth probes, then the T compares.  The T walks are independent
// dependency chains, so a single thread keeps T header loads / 2T probes in flight (memory-level parallelism
// without needing T times the warps), and the summation order of leaf pdfs stays tree 0..T-1 (SURVEY note N1).
//
// Instruction budget per node-step (the kernel is issue-bound, profiles/r01_ncu_eval_v1.md): explicit child ids (one
// select per level), integer thresholds (no int->float conversions), magic-number floor (no float->int conversions),
// 32-bit image indices (no 64-bit address arithmetic per probe).
#pragma once
#include "rdf_common.cuh"
#include <type_traits>

#define RDF_FAST_MAX_TREES 8

struct rdf_forest_view {
    const rdf_node_hdr* hdr;
    const float* pdf;
    int nodes_per_tree;
    int T, D, C, CP;
};

static inline rdf_forest_view rdf_view(const rdf_forest* f) {
    rdf_forest_view v;
    v.hdr = f->hdr;
    v.pdf = f->pdf;
    v.nodes_per_tree = (int)f->rows_per_tree;      // slots per tree in the packed arrays: roots sit at t * nodes_per_tree
    v.T = f->T;
    v.D = f->D;
    v.C = f->C;
    v.CP = f->CP;
    return v;
}

// One node-step of one walk: header (two 128-bit loads) -> feature -> next node id (>= 0) or ~leaf_id / RDF_NO_LEAF (< 0).
struct rdf_hdr_regs {
    float4 a;
    int4 b;          // ithresh, left, right, flags
};

// one 256-bit load per header (LDG.E.256, new with sm_100): half the L1 tag lookups of two 128-bit loads
__device__ __forceinline__ rdf_hdr_regs rdf_load_hdr(const rdf_node_hdr* __restrict__ hdr, int node) {
    rdf_hdr_regs h;
    asm("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %8, 32, %9;\n\t"
        "ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [a];\n\t}"
        : "=f"(h.a.x), "=f"(h.a.y), "=f"(h.a.z), "=f"(h.a.w), "=r"(h.b.x), "=r"(h.b.y), "=r"(h.b.z), "=r"(h.b.w)
        : "r"(node), "l"(hdr));
    return h;
}

// Levels [j0, j1) of T interleaved walks.  SMEM = true: headers come from a shared-memory copy of the upper levels (index =
// t * stride + row, child ids already rewritten to that indexing, see rdf_stage_upper_levels); else from the packed forest.
// state[t] >= 0: current node; < 0: ended (~leaf_id or RDF_NO_LEAF).
// NEVER_EXACT: the forest holds no node flagged RDF_FLAG_EXACT_DIV (known on the host after packing) and the scale is in the fast
// domain, so the loop carries neither the flag test nor the __fdiv_rn path.
// COMPLETE: every tree of the forest is a complete tree (no leaf above level D-1, known on the host after packing): no walk ends
// early, so the loop needs neither the "all walks ended" exit nor the per-tree "ended" clamp and select.
template <int T, bool SCALE1, bool FORCE_EXACT, bool SMEM, bool NEVER_EXACT = false, bool COMPLETE = false, typename HDR = const rdf_node_hdr*>
__device__ __forceinline__ void rdf_walk_levels(HDR hdr, int tree_stride, int j0, int j1,
                                                const uint16_t* __restrict__ img, int W, int H, int X, int Y, float df, float rcp,
                                                float xm, float ym, float scale, int (&state)[T]) {
    for (int j = j0; j < j1; j++) {
        if (!COMPLETE) {
            int all = state[0];
#pragma unroll
            for (int t = 1; t < T; t++) all &= state[t];
            if (all < 0) break;                                      // every walk of this pixel has ended
        }
        rdf_hdr_regs h[T];
        int any_flags = 0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            // ended walks re-read their tree's root (cached): harmless, keeps the loop branch-free
            const int node = COMPLETE ? state[t] : max(state[t], t * tree_stride);
            if constexpr (SMEM) {
                if constexpr (std::is_pointer<HDR>::value) {
                    const float4* sp = reinterpret_cast<const float4*>(hdr + node);
                    h[t].a = sp[0];
                    h[t].b = *reinterpret_cast<const int4*>(sp + 1);
                } else {                                             // kernel-parameter array: constant-bank loads with a register index
                    h[t].a = hdr.top[node].a;
                    h[t].b = make_int4(hdr.top[node].ithresh, hdr.top[node].left, hdr.top[node].right, hdr.top[node].flags);
                }
            } else {
                h[t] = rdf_load_hdr(hdr, node);
            }
            if (!NEVER_EXACT) any_flags |= h[t].b.w;
            if (!SCALE1) {
                h[t].a.x = __fmul_rn(scale, h[t].a.x);
                h[t].a.y = __fmul_rn(scale, h[t].a.y);
                h[t].a.z = __fmul_rn(scale, h[t].a.z);
                h[t].a.w = __fmul_rn(scale, h[t].a.w);
            }
        }
        int f[T];
        if (!NEVER_EXACT && (FORCE_EXACT || (any_flags & RDF_FLAG_EXACT_DIV))) {
#pragma unroll
            for (int t = 0; t < T; t++)
                f[t] = rdf_feature_i<true>(img, W, H, X, Y, df, rcp, xm, ym, h[t].a.x, h[t].a.y, h[t].a.z, h[t].a.w);
        } else {
#pragma unroll
            for (int t = 0; t < T; t++)
                f[t] = rdf_feature_i<false>(img, W, H, X, Y, df, rcp, xm, ym, h[t].a.x, h[t].a.y, h[t].a.z, h[t].a.w);
        }
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int next = (f[t] < h[t].b.x) ? h[t].b.y : h[t].b.z;   // NaN threshold -> INT_MIN -> right (tree_eval.cu:106)
            state[t] = (COMPLETE || state[t] >= 0) ? next : state[t];
        }
    }
}

// Upper levels of every tree in shared memory (north_star: "upper tree levels are packed into shared memory"): on those
// levels most lanes of a warp sit on the same node, and a broadcast pair of 128-bit shared loads costs 2 L1 data-pipe
// wavefronts where the 256-bit global load costs 8 whatever the addresses (profiles/r01_micro_l1.md).
// Copies levels 0 .. KS-1 of the T trees to hdr_s[t * M + row], M = 2^KS - 1, rewriting child ids that stay inside the copy
// (levels < KS-1) to the shared indexing; children of level KS-1 keep their global ids, leaves stay negative.
// Must be called by all threads of the block, followed by __syncthreads().
__device__ __forceinline__ void rdf_stage_upper_levels(const rdf_forest_view& fv, int KS, rdf_node_hdr* __restrict__ hdr_s) {
    const int M = (1 << KS) - 1;
    const int first_last = (1 << (KS - 1)) - 1;                      // rows >= this are on level KS-1
    for (int i = threadIdx.x; i < fv.T * M; i += blockDim.x) {
        const int t = i / M, row = i - t * M;
        rdf_hdr_regs h = rdf_load_hdr(fv.hdr, t * fv.nodes_per_tree + row);
        if (row < first_last) {
            const int shift = t * M - t * fv.nodes_per_tree;
            if (h.b.y >= 0) h.b.y += shift;
            if (h.b.z >= 0) h.b.z += shift;
        }
        float4* sp = reinterpret_cast<float4*>(hdr_s + i);
        sp[0] = h.a;
        *reinterpret_cast<int4*>(sp + 1) = h.b;
    }
}

// Walk T trees from their roots.  On return state[t] < 0: ~leaf_id of the reached leaf, or RDF_NO_LEAF if the walk fell
// off level D-1 with a "continue" flag (adds nothing, src/cuda/tree_eval.cu:95-128).
// SCALE1: scale == 1.0f, so scale*u == u exactly and the multiplies are dropped.  FORCE_EXACT: always use __fdiv_rn
// (scale outside the fast domain); otherwise nodes flagged RDF_FLAG_EXACT_DIV take the exact path per level.
// hdr_s / KS: optional copy of levels 0 .. KS-1, either a shared-memory pointer (rdf_stage_upper_levels) or a reference wrapper around a kernel-parameter array (rdf_eval.cu: rdf_top_ref, constant-bank loads); KS = 0: everything from global memory.
template <int T, bool SCALE1, bool FORCE_EXACT, bool NEVER_EXACT = false, bool COMPLETE = false, typename TOP = const rdf_node_hdr*>
__device__ __forceinline__ void rdf_walk(const rdf_forest_view& fv, const uint16_t* __restrict__ img, int W, int H, int X,
                                         int Y, unsigned d, float scale, int (&state)[T], TOP hdr_s = nullptr,
                                         int KS = 0) {
    const float df = (float)d;
    const float rcp = __frcp_rn(df);                                 // RN(1/d), once per pixel
    const float xm = (float)X + RDF_MAGIC_F, ym = (float)Y + RDF_MAGIC_F;   // exact: X, Y < 2^16
    if (KS > 0) {
        const int M = (1 << KS) - 1;
#pragma unroll
        for (int t = 0; t < T; t++) state[t] = t * M;
        rdf_walk_levels<T, SCALE1, FORCE_EXACT, true, NEVER_EXACT, COMPLETE, TOP>(hdr_s, M, 0, KS, img, W, H, X, Y, df, rcp, xm, ym, scale, state);
    } else {
#pragma unroll
        for (int t = 0; t < T; t++) state[t] = t * fv.nodes_per_tree;
    }
    rdf_walk_levels<T, SCALE1, FORCE_EXACT, false, NEVER_EXACT, COMPLETE>(fv.hdr, fv.nodes_per_tree, KS, fv.D, img, W, H, X, Y, df, rcp, xm, ym, scale, state);
#pragma unroll
    for (int t = 0; t < T; t++)
        if (state[t] >= 0) state[t] = RDF_NO_LEAF;                   // D levels done and still on a node (cannot happen
                                                                     // with a packed forest, kept as a guard)
}

// get_best_pdf_chance over the tree-ordered sum of the reached leaf pdfs (src/cuda/tree_eval.cu:7-21,125).
// probs (optional): receives sum / T per class.
template <int T>
__device__ __forceinline__ int rdf_vote(const rdf_forest_view& fv, const int (&state)[T], float* __restrict__ probs) {
    float best = 0.f;
    int lab = 0;
    const float inv_t = (float)fv.T;
    for (int c = 0; c < fv.CP; c += 4) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < T; t++) {
            if (state[t] != RDF_NO_LEAF) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(fv.pdf + (size_t)(unsigned)(~state[t]) * fv.CP + c));
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
                                              int X, int Y, unsigned d, float scale, float* __restrict__ probs) {
#define RDF_CASE(TT)                                                            \
    case TT: {                                                                  \
        int st[TT];                                                             \
        rdf_walk<TT, SCALE1, FORCE_EXACT>(fv, img, W, H, X, Y, d, scale, st);   \
        return rdf_vote<TT>(fv, st, probs);                                     \
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
