/*
 * Plain-C CPU restatement of the 3d-beats RDF hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library (oracle/_build/librdf_oracle.so).  The product never links or calls it.
 *
 * Parity pinning: see the header of oracle/numpy_oracle.py - the reference has no golden vectors for this
 * path; both oracles are pinned against the reference's own kernels compiled unchanged for sm_100a
 * (oracle/ref_kernels) and against tests/golden/*.npz produced by those kernels on a B200.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: every fp32 op below must round exactly once, like the PTX of the reference
 * (mul.f32, div.rn.f32, cvt.rmi.s32.f32, sub.f32); fused ops appear only where the reference's compiled
 * code has them (gini, via fmaf).
 *
 * Citations are relative to the reference repository root.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_UINT16 65535 /* src/cuda/cu_utils.hpp:8 */

/* __float2int_rd == cvt.rmi.s32.f32: floor, saturate, NaN -> 0 */
static inline int32_t float2int_rd(float x) {
    if (x != x) return 0;
    float f = floorf(x);
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

/* Array3d<uint16>::get with default 65535 (src/cuda/cu_utils.hpp:58-62,79-86): bounds are per image */
static inline uint16_t probe(const uint16_t* img, int H, int W, int32_t yy, int32_t xx) {
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) return MAX_UINT16;
    return img[(size_t)yy * W + xx];
}

/* compute_feature (src/cuda/decision_tree_common.hpp:8-28) */
static inline float compute_feature(const uint16_t* img, int H, int W, int X, int Y,
                                    float ux, float uy, float vx, float vy, float scale) {
    const uint16_t d = img[(size_t)Y * W + X];
    if (d == 0) return 0.f;
    const float df = (float)d;
    volatile float sux = scale * ux, suy = scale * uy, svx = scale * vx, svy = scale * vy; /* one rounding each */
    const int32_t oux = float2int_rd(sux / df), ouy = float2int_rd(suy / df);
    const int32_t ovx = float2int_rd(svx / df), ovy = float2int_rd(svy / df);
    const float pu = (float)probe(img, H, W, wrap_add(Y, ouy), wrap_add(X, oux));
    const float pv = (float)probe(img, H, W, wrap_add(Y, ovy), wrap_add(X, ovx));
    return pu - pv;
}

/* Traverse one tree (src/cuda/tree_eval.cu:95-128, node addressing src/cuda/cu_utils.hpp:32-39).
 * Returns pointer to the reached leaf pdf, or NULL when the walk falls off level D-1 with a -1 flag. */
static inline const float* traverse(const float* tree, int D, int C, const uint16_t* img, int H, int W,
                                    int X, int Y, float scale) {
    const int E = 7 + 2 * C;
    int64_t g = 0;
    for (int j = 0; j < D; j++) {
        const float* nd = tree + (((int64_t)1 << j) - 1 + g) * E;
        const float f = compute_feature(img, H, W, X, Y, nd[0], nd[1], nd[2], nd[3], scale);
        if (f < nd[4]) {
            if (float2int_rd(nd[5]) == -1) g = 2 * g; else return nd + 7;
        } else {
            if (float2int_rd(nd[6]) == -1) g = 2 * g + 1; else return nd + 7 + C;
        }
    }
    return NULL;
}

/* get_best_pdf_chance (src/cuda/tree_eval.cu:7-21) */
static inline int best_pdf_chance(const float* pdf, int C) {
    float best = 0.f; int lab = 0;
    for (int c = 0; c < C; c++) if (pdf[c] > best) { best = pdf[c]; lab = c; }
    return lab;
}

/* evaluate_image_using_forest (src/cuda/tree_eval.cu:24-137) + host convention src/decision_tree.py:298-330.
 * filter may be NULL (== filter_class -1).  probs may be NULL; else float[N,h,w,C] receives sum/T.
 * Skipped pixels are not written.  Tree order 0..T-1 fixes the fp32 accumulation order (SURVEY note N1). */
int oracle_eval_forest(const float* forest, int T, int D, int C, const uint16_t* depth, int N, int H, int W,
                       const uint16_t* filter, int filter_class, uint16_t* labels, float* probs,
                       int labels_reduce, float scale, int nthreads) {
    if (C > 256 || C < 1 || labels_reduce < 1) return -1;
    const int h = H / labels_reduce, w = W / labels_reduce;
    const int64_t tree_stride = (((int64_t)1 << D) - 1) * (7 + 2 * C);
    if (filter == NULL) filter_class = -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const int64_t rows = (int64_t)N * h;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < rows; r++) {
        const int n = (int)(r / h), y = (int)(r % h);
        const uint16_t* img = depth + (size_t)n * H * W;
        for (int x = 0; x < w; x++) {
            const size_t li = ((size_t)n * h + y) * w + x;
            if (filter_class != -1 && (int)filter[li] != filter_class) continue;
            const int Y = y * labels_reduce, X = x * labels_reduce;
            const uint16_t d = img[(size_t)Y * W + X];
            if (d == 0 || d == MAX_UINT16) continue;
            float acc[256];
            for (int c = 0; c < C; c++) acc[c] = 0.f;
            for (int t = 0; t < T; t++) {
                const float* pdf = traverse(forest + t * tree_stride, D, C, img, H, W, X, Y, scale);
                if (pdf) for (int c = 0; c < C; c++) acc[c] = acc[c] + pdf[c];
            }
            labels[li] = (uint16_t)best_pdf_chance(acc, C);
            if (probs) for (int c = 0; c < C; c++) probs[li * C + c] = acc[c] / (float)T;
        }
    }
    return 0;
}

/* evaluate_image_using_tree (src/cuda/tree_eval.cu:140-212) */
int oracle_eval_tree(const float* tree, int D, int C, const uint16_t* depth, int N, int H, int W,
                     uint16_t* labels, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const int64_t rows = (int64_t)N * H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < rows; r++) {
        const int n = (int)(r / H), y = (int)(r % H);
        const uint16_t* img = depth + (size_t)n * H * W;
        for (int x = 0; x < W; x++) {
            const uint16_t d = img[(size_t)y * W + x];
            if (d == 0 || d == MAX_UINT16) continue;
            const float* pdf = traverse(tree, D, C, img, H, W, x, y, 1.f);
            if (pdf) labels[((size_t)n * H + y) * W + x] = (uint16_t)best_pdf_chance(pdf, C);
        }
    }
    return 0;
}

/* make_composite_labels_image (src/cuda/tree_eval.cu:214-248) */
int oracle_composite(const uint16_t* const* label_images, int L, int w, int h, const int32_t* conditions,
                     uint16_t* composite) {
    for (int64_t i = 0; i < (int64_t)w * h; i++) {
        int off = 0;
        for (int k = 0; k < L; k++) {
            const uint16_t l = label_images[k][i];
            if (l == 0 || l == MAX_UINT16) break;
            const int32_t* tv = conditions + 2 * (off + l - 1);
            if (tv[0] == 0) { composite[i] = (uint16_t)tv[1]; break; }
            off = tv[1];
        }
    }
    return 0;
}

/* MeanShift.run (src/cuda/mean_shift.py:19-59) + kernel run (src/cuda/mean_shift.cu:3-48); means = double[K,2] (x,y) */
int oracle_mean_shift(const uint16_t* labels, int w, int h, int K, const float* variances, int rounds, double* means) {
    double* S = (double*)malloc(sizeof(double) * 3 * K);
    for (int k = 0; k < 2 * K; k++) means[k] = 0.0;
    for (int it = 0; it < rounds; it++) {
        for (int k = 0; k < 3 * K; k++) S[k] = 0.0;
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
            const uint16_t l = labels[(size_t)y * w + x];
            if (l == 0 || l == MAX_UINT16 || l > K) continue;
            const int k = l - 1;
            if (it == 0) { S[3 * k] += (double)x; S[3 * k + 1] += (double)y; S[3 * k + 2] += 1.0; }
            else {
                const double dx = (double)x - means[2 * k], dy = (double)y - means[2 * k + 1];
                volatile float v2f = variances[k] * variances[k];       /* fp32 product (mean_shift.cu:41) */
                const double p = exp(-(dx * dx + dy * dy) / (2 * (double)v2f));
                S[3 * k] += dx * p; S[3 * k + 1] += dy * p; S[3 * k + 2] += p;
            }
        }
        for (int k = 0; k < K; k++) { means[2 * k] += S[3 * k] / S[3 * k + 2]; means[2 * k + 1] += S[3 * k + 1] / S[3 * k + 2]; }
    }
    free(S);
    return 0;
}

/* Generalised split histogram (SURVEY 8d cfg 4), restating evaluate_random_features (src/cuda/tree_train.cu:4-64):
 * hist[slot][feature][bin][label] += 1 with bin = #{k : thresholds[feature][k] <= f}.  node_slot maps node id -> slot or -1.
 * hist is uint32[num_slots, F, NT+1, C] and must be zeroed by the caller. */
int oracle_train_hist(const uint16_t* depth, const uint16_t* labels, const int32_t* nodes_by_pixel, int N, int H, int W,
                      const int32_t* node_slot, const float* offsets, const float* thresholds, int F, int NT, int C,
                      int f_begin, int f_end, uint32_t* hist, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const int NB = NT + 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = f_begin; j < f_end; j++) {
        const float* o = offsets + 4 * j;
        const float* th = thresholds + (size_t)NT * j;
        for (int n = 0; n < N; n++) {
            const uint16_t* img = depth + (size_t)n * H * W;
            for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
                const size_t pi = ((size_t)n * H + y) * W + x;
                const int32_t g = nodes_by_pixel[pi];
                if (g < 0) continue;
                const int32_t slot = node_slot[g];
                if (slot < 0) continue;
                const float f = compute_feature(img, H, W, x, y, o[0], o[1], o[2], o[3], 1.f);
                int b = 0;
                while (b < NT && th[b] <= f) b++;
                hist[(((size_t)slot * F + j) * NB + b) * C + labels[pi]] += 1;
            }
        }
    }
    return 0;
}

/* gini_impurity / gini_gain (src/cuda/tree_train.cu:72-89) in the operation order nvcc 12.9 emits for sm_100a
 * (verified in PTX): p = fma(c_i/s, c_i/s, p); left_term = (l/p)*gini(l) [mul]; rem = fma(r/p, gini(r), left_term);
 * gain = gini(parent) - rem. */
static float gini_impurity(const uint64_t* c, int C) {
    uint64_t s = 0; for (int i = 0; i < C; i++) s += c[i];
    const float sf = (float)s; float p = 0.f;
    for (int i = 0; i < C; i++) { const float pi = (float)c[i] / sf; p = fmaf(pi, pi, p); }
    return 1.f - p;
}
float oracle_gini_gain(const uint64_t* parent, const uint64_t* left, const uint64_t* right, int C) {
    uint64_t ps = 0, ls = 0, rs = 0;
    for (int i = 0; i < C; i++) { ps += parent[i]; ls += left[i]; rs += right[i]; }
    if (!ls || !rs) return 0.f;                                            /* :158-160 */
    const float pf = (float)ps;
    volatile float lt = ((float)ls / pf) * gini_impurity(left, C);
    const float rem = fmaf((float)rs / pf, gini_impurity(right, C), lt);
    return gini_impurity(parent, C) - rem;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
