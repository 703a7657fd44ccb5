// Mean-shift fingertip centroids: replaces MeanShift.run (reference src/cuda/mean_shift.py:19-59) and its kernel
// `run` (src/cuda/mean_shift.cu:3-48).
//
// The reference runs, per round, one fill + one kernel doing three global fp64 atomics per labelled pixel onto K*3
// addresses + two blocking D2H copies + a host divide + one H2D copy.  Here all rounds run in ONE launch of a single
// thread-block cluster:
//   1. each CTA counting-sorts the labelled pixels of its slice of the image by class into a compact coordinate
//      list (ballot/match based, deterministic order, no atomics);
//   2. per round, warps reduce fixed-size items of one class each with shuffles (fp64, fixed order), CTAs exchange
//      their K*3 partial sums through distributed shared memory, and after one cluster barrier every CTA computes
//      the new means redundantly - no global atomics, no grid relaunch, no host round trip.
// Results are bitwise reproducible run to run for a given launch shape.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "rdf_common.cuh"
#include "rdf_fingertip.cuh"

namespace cg = cooperative_groups;

#define MS_THREADS 512
#define MS_WARPS (MS_THREADS / 32)
#define MS_MAX_ITEMS 1024
#define MS_MAX_CLUSTER 8

struct rdf_ms_params {
    const uint16_t* labels;
    const float* variances;
    double* means_out;
    uint32_t* entries;      // workspace: uint32[h*w], x | y << 16
    int w, h, K, rounds;
    int chunk;              // pixels per CTA (multiple of 32 * MS_WARPS)
    int item;               // entries per work item (multiple of 32)
};

__device__ __forceinline__ double ms_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(MS_THREADS) rdf_mean_shift_kernel(const rdf_ms_params p) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int NC = (int)cluster.num_blocks();
    const int K = p.K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(16) unsigned char ms_smem[];
    double* all_partial = reinterpret_cast<double*>(ms_smem);                 // [2][NC][K][3]
    double* partial = all_partial + 2 * (size_t)NC * K * 3;                   // [MS_MAX_ITEMS][3]
    double* means_s = partial + MS_MAX_ITEMS * 3;                             // [K][2]
    double* v2_s = means_s + 2 * K;                                           // [K]
    int* cnt = reinterpret_cast<int*>(v2_s + K);                              // [MS_WARPS][K]  (counts, then cursors)
    int* seg_start = cnt + MS_WARPS * K;                                      // [K+1]
    int* item_start = seg_start + (K + 1);                                    // [K+1]

    const int npx = p.w * p.h;
    const int p0 = rank * p.chunk;
    const int p1 = min(npx, p0 + p.chunk);
    const int warp_chunk = p.chunk / MS_WARPS;                                // multiple of 32
    const int wp0 = p0 + warp * warp_chunk;
    const int wp1 = min(p1, wp0 + warp_chunk);

    for (int i = tid; i < MS_WARPS * K; i += MS_THREADS) cnt[i] = 0;
    for (int k = tid; k < K; k += MS_THREADS) {
        means_s[2 * k] = 0.0;
        means_s[2 * k + 1] = 0.0;
        const float v = p.variances[k];
        v2_s[k] = (double)__fmul_rn(v, v);                                    // fp32 product, widened (mean_shift.cu:41)
    }
    __syncthreads();

    // ---- pass A: per-warp class counts --------------------------------------------------------------------
    for (int base = wp0; base < wp1; base += 32) {
        const int px = base + lane;
        int key = -1;
        if (px < wp1) {
            const unsigned l = __ldg(p.labels + px);
            if (l != 0u && l != RDF_NO_PIXEL && (int)l <= K) key = (int)l - 1;  // mean_shift.cu:23
        }
        const unsigned grp = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == __ffs(grp) - 1) cnt[warp * K + key] += __popc(grp);
        __syncwarp();
    }
    __syncthreads();

    // ---- class segment offsets and per-warp cursors (class-major, warp-minor => raster order within a class) ----
    for (int k = tid; k < K; k += MS_THREADS) {
        int s = 0;
        for (int wv = 0; wv < MS_WARPS; wv++) s += cnt[wv * K + k];
        seg_start[k + 1] = s;                                                 // lengths for now
    }
    __syncthreads();
    if (tid == 0) {
        seg_start[0] = 0;
        item_start[0] = 0;
        for (int k = 0; k < K; k++) {
            const int len = seg_start[k + 1];
            seg_start[k + 1] = seg_start[k] + len;
            item_start[k + 1] = item_start[k] + (len + p.item - 1) / p.item;
        }
    }
    __syncthreads();
    for (int k = tid; k < K; k += MS_THREADS) {
        int s = seg_start[k];
        for (int wv = 0; wv < MS_WARPS; wv++) {
            const int c = cnt[wv * K + k];
            cnt[wv * K + k] = s;
            s += c;
        }
    }
    __syncthreads();

    // ---- pass B: scatter coordinates into the class-sorted list ----------------------------------------------
    uint32_t* my_entries = p.entries + p0;
    for (int base = wp0; base < wp1; base += 32) {
        const int px = base + lane;
        int key = -1;
        if (px < wp1) {
            const unsigned l = __ldg(p.labels + px);
            if (l != 0u && l != RDF_NO_PIXEL && (int)l <= K) key = (int)l - 1;
        }
        const unsigned grp = __match_any_sync(0xffffffffu, key);
        int pos = 0;
        if (key >= 0) pos = cnt[warp * K + key] + __popc(grp & ((1u << lane) - 1u));
        __syncwarp();
        if (key >= 0) {
            const int y = px / p.w, x = px - y * p.w;
            my_entries[pos] = (uint32_t)x | ((uint32_t)y << 16);
            if (lane == __ffs(grp) - 1) cnt[warp * K + key] += __popc(grp);
        }
        __syncwarp();
    }
    __syncthreads();   // entries written by this CTA are read back by this CTA only

    // ---- rounds ----------------------------------------------------------------------------------------------
    const int n_items = item_start[K];
    for (int it = 0; it < p.rounds; it++) {
        for (int item = warp; item < n_items; item += MS_WARPS) {
            // class of this item: last k with item_start[k] <= item
            int lo = 0, hi = K - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (item_start[mid] <= item) lo = mid; else hi = mid - 1;
            }
            const int k = lo;
            const int e0 = seg_start[k] + (item - item_start[k]) * p.item;
            const int e1 = min(seg_start[k + 1], e0 + p.item);
            const double mx = means_s[2 * k], my = means_s[2 * k + 1];
            const double two_v2 = 2.0 * v2_s[k];
            double sx = 0.0, sy = 0.0, sp = 0.0;
            for (int e = e0 + lane; e < e1; e += 32) {
                const uint32_t c = my_entries[e];
                const double cx = (double)(c & 0xffffu), cy = (double)(c >> 16);
                if (it == 0) {                                               // mean_shift.cu:31-34
                    sx += cx;
                    sy += cy;
                    sp += 1.0;
                } else {                                                     // mean_shift.cu:36-46
                    const double dx = cx - mx, dy = cy - my;
                    const double pr = exp(-(dx * dx + dy * dy) / two_v2);
                    sx += dx * pr;
                    sy += dy * pr;
                    sp += pr;
                }
            }
            sx = ms_warp_sum(sx);
            sy = ms_warp_sum(sy);
            sp = ms_warp_sum(sp);
            if (lane == 0) {
                partial[item * 3 + 0] = sx;
                partial[item * 3 + 1] = sy;
                partial[item * 3 + 2] = sp;
            }
        }
        __syncthreads();
        // CTA partial per (class, component), items in order; broadcast to every CTA of the cluster through DSMEM
        double* buf = all_partial + (size_t)(it & 1) * NC * K * 3;
        for (int i = tid; i < 3 * K; i += MS_THREADS) {
            const int k = i / 3, comp = i - 3 * k;
            double s = 0.0;
            for (int item = item_start[k]; item < item_start[k + 1]; item++) s += partial[item * 3 + comp];
            for (int r = 0; r < NC; r++) {
                double* remote = cluster.map_shared_rank(buf, r);
                remote[((size_t)rank * K + k) * 3 + comp] = s;
            }
        }
        cluster.sync();
        for (int k = tid; k < K; k += MS_THREADS) {
            double sx = 0.0, sy = 0.0, sp = 0.0;
            for (int r = 0; r < NC; r++) {
                sx += buf[((size_t)r * K + k) * 3 + 0];
                sy += buf[((size_t)r * K + k) * 3 + 1];
                sp += buf[((size_t)r * K + k) * 3 + 2];
            }
            means_s[2 * k] += sx / sp;                                       // mean_shift.py:53-55 (0/0 -> NaN)
            means_s[2 * k + 1] += sy / sp;
        }
        __syncthreads();
    }
    if (rank == 0)
        for (int i = tid; i < 2 * K; i += MS_THREADS) p.means_out[i] = means_s[i];
    cluster.sync();   // no CTA may exit while peers can still address its shared memory
}


__device__ __forceinline__ unsigned long long ms_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MS_TRACE(slot)                                                        \
    do {                                                                      \
        if (p.trace && rank == 0 && tid == 0) p.trace[slot] = ms_now();       \
    } while (0)

// exp(x) for x <= 0 (or NaN), straight-line: the libdevice exp is ~60 dependent fp64 instructions with branches, and the
// rounds of the latency path are one long dependency chain per lane.  Cody-Waite reduction x = k ln2 + r, |r| <= ln2/2,
// degree-13 Taylor polynomial in Estrin form (truncation 4e-18, ~11 dependent operations), 2^k applied in two factors so that
// results below 2^-1022 underflow gradually like exp() does.  Differs from exp() by a few ulp at most (centroids are compared
// at 1e-5); NaN stays NaN (a zero variance with a pixel on the mean gives 0/0, as in the reference).
__device__ __forceinline__ double ms_exp_nonpos(double x) {
    x = x < -746.0 ? -746.0 : x;                                            // exp(-746) == 0 in fp64; keeps NaN
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);    // 1.5 * 2^52: low word of t = rint(x log2 e)
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = fma(1.0, r, 1.0);
    const double a1 = fma(1.0 / 6.0, r, 0.5);
    const double a2 = fma(1.0 / 120.0, r, 1.0 / 24.0);
    const double a3 = fma(1.0 / 5040.0, r, 1.0 / 720.0);
    const double a4 = fma(1.0 / 362880.0, r, 1.0 / 40320.0);
    const double a5 = fma(1.0 / 39916800.0, r, 1.0 / 3628800.0);
    const double a6 = fma(1.0 / 6227020800.0, r, 1.0 / 479001600.0);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double d0 = fma(b1, r4, b0), d1 = fma(a6, r4, b2);
    const double pl = fma(d1, r8, d0);
    const int k1 = k >> 1, k2 = k - k1;
    return (pl * __hiloint2double((k1 + 1023) << 20, 0)) * __hiloint2double((k2 + 1023) << 20, 0);
}

// =================================================================================================================
// v3: class-parallel latency path (THE DEFAULT; rdf_mean_shift_kernel above is the one documented fallback, taken only for
// images above MS3_MAX_R x MS3_CAP = 407 552 pixels or a label pointer that is not 16-byte aligned).  The classes never
// interact (each has its own mean), so instead of spreading the PIXELS over one cluster and meeting at a cluster barrier every
// round (the retired pixel-parallel shared-memory version: ~2.4 us per round, mostly synchronisation), every CLASS
// gets its own small cluster of R CTAs (R = 1 for images up to 50 880 pixels, 2 for the product's 424 x 240 label image,
// 8 for 848 x 480).  Each CTA filters its share of the label image for its class (128-bit loads, the image is read once per
// class out of L2), compacts the coordinates into shared memory (one 32-bit block scan), and then iterates: every thread owns
// fixed entries, the CTA reduces with shuffles, and only when R > 1 the R partial sums cross distributed shared memory.
// Results are bitwise reproducible (fixed entry -> thread mapping, fixed reduction trees).
// =================================================================================================================
#define MS3_THREADS 1024
#define MS3_WARPS 32
#define MS3_GROUPS 7                      // 8-pixel groups per thread: up to 57 344 pixels per CTA
#define MS3_CAP 50944                     // pixels per CTA (= entries it may have to hold): 8 CTAs cover 848 x 480
#define MS3_MAX_R 8

struct rdf_ms3_params {
    const uint16_t* labels;
    const float* variances;
    double* means_out;
    int w, h, K, rounds, R;
    unsigned long long* trace;   // optional: %globaltimer stamps of class 0 / rank 0 (workspace head), phase by phase
    // batch: cluster c serves class c % K of image c / K (labels + image * w*h, means_out + image * 2K)
    int with_fingertips;         // fused read-out (rdf_mean_shift_fingertips): ft below is valid
    rf_fingertip_spec ft;
};
#define MS3_TRACE(slot)                                                              \
    do {                                                                             \
        if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[slot] = ms_now();        \
    } while (0)

__global__ void __launch_bounds__(MS3_THREADS, 1) rdf_mean_shift_v3_kernel(const rdf_ms3_params p) {
    cg::cluster_group cluster = cg::this_cluster();
    const int R = p.R;
    const int rank = R > 1 ? (int)cluster.block_rank() : 0;
    const int ck = blockIdx.x / R;
    const int img = ck / p.K;                                                // image of this cluster (batched launch)
    const int k = ck - img * p.K;                                            // class of this cluster (label k + 1)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(16) unsigned char ms_smem[];
    double* xpart = reinterpret_cast<double*>(ms_smem);                       // [2][MS3_MAX_R][3] partial sums of the cluster's CTAs
    double* wpart = xpart + 2 * MS3_MAX_R * 3;                                // [2][MS3_WARPS][3] warp partial sums, by round parity
    int* wtot = reinterpret_cast<int*>(wpart + 2 * MS3_WARPS * 3);            // [MS3_WARPS]
    uint32_t* entries = reinterpret_cast<uint32_t*>(wtot + MS3_WARPS);        // [<= MS3_CAP]

    asm volatile("griddepcontrol.wait;" ::: "memory");                        // programmatic dependent launch (see v2)
    const int npx = p.w * p.h;
    const uint16_t* __restrict__ labels = p.labels + (size_t)img * npx;
    const unsigned want = (unsigned)k + 1u;
    MS3_TRACE(0);

    // ---- my 8-pixel groups (dealt round-robin to the R CTAs), all loads issued first ----
    // groups a thread can own at all at this image size and cluster size (block-uniform): the rest of the unrolled code is skipped
    const int gmax = ((npx + 7) / 8 + MS3_THREADS * R - 1) / (MS3_THREADS * R);
    uint4 px[MS3_GROUPS];
#pragma unroll
    for (int g = 0; g < MS3_GROUPS; g++) {
        const int q = ((g * MS3_THREADS + tid) * R + rank) * 8;
        px[g] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (g >= gmax) continue;
        if (q + 8 <= npx) {
            px[g] = __ldg(reinterpret_cast<const uint4*>(labels + q));
        } else if (q < npx) {
            unsigned short tmp[8];
#pragma unroll
            for (int j = 0; j < 8; j++) tmp[j] = q + j < npx ? __ldg(labels + q + j) : (unsigned short)0xffff;
            px[g] = make_uint4(tmp[0] | (tmp[1] << 16), tmp[2] | (tmp[3] << 16), tmp[4] | (tmp[5] << 16), tmp[6] | (tmp[7] << 16));
        }
    }
    // match mask of my pixels: bit (8 g + j) set when pixel j of group g carries this class' label (two labels per SIMD compare)
    const unsigned want2 = want | (want << 16);
    unsigned long long mask = 0ull;
#pragma unroll
    for (int g = 0; g < MS3_GROUPS; g++) {
        if (g >= gmax) continue;
        const unsigned w0 = __vcmpeq2(px[g].x, want2), w1 = __vcmpeq2(px[g].y, want2), w2 = __vcmpeq2(px[g].z, want2),
                       w3 = __vcmpeq2(px[g].w, want2);                   // 0xffff per matching half-word
        const unsigned m8 = (w0 & 1u) | ((w0 >> 15) & 2u) | ((w1 & 1u) << 2) | ((w1 >> 13) & 8u) | ((w2 & 1u) << 4) | ((w2 >> 11) & 32u) |
                            ((w3 & 1u) << 6) | ((w3 >> 9) & 128u);
        mask |= (unsigned long long)m8 << (8 * g);
    }
    const int cnt = __popcll(mask);
    MS3_TRACE(1);
    // ---- block exclusive scan -> positions (thread-major, then pixel order: deterministic) ----
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = wtot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        wtot[lane] = t;
    }
    __syncthreads();
    const int n = wtot[MS3_WARPS - 1];                                        // entries of this CTA
    MS3_TRACE(2);
    {
        int pos = (warp ? wtot[warp - 1] : 0) + incl - cnt;
        // only the matching pixels are visited; one division per 8-pixel group (its first pixel), not per pixel
#pragma unroll
        for (int g = 0; g < MS3_GROUPS; g++) {
            if (g >= gmax) continue;
            unsigned m8 = (unsigned)(mask >> (8 * g)) & 0xffu;
            if (!m8) continue;
            const int q0 = ((g * MS3_THREADS + tid) * R + rank) * 8;
            const int y0 = q0 / p.w, x0 = q0 - y0 * p.w;
            for (; m8; m8 &= m8 - 1) {
                int x = x0 + __ffs(m8) - 1, y = y0;
                while (x >= p.w) {                                           // the group wraps into the next row(s)
                    x -= p.w;
                    y++;
                }
                entries[pos++] = (uint32_t)x | ((uint32_t)y << 16);
            }
        }
    }
    const float var = p.variances[k];
    const double nis = -1.0 / (2.0 * (double)__fmul_rn(var, var));            // sigma^2 = fp32 product, widened (mean_shift.cu:41)
    __syncthreads();

    MS3_TRACE(3);
    // ---- rounds ----
    // Barriers per round: one block barrier (warp partials visible) and, for R > 1, the cluster barrier.  The new mean is then
    // computed redundantly by every thread that owns entries (nobody else needs it), from buffers that alternate by round
    // parity, so no broadcast barrier is needed.
    const int nw = (min(n, MS3_THREADS) + 31) >> 5;                           // warps that own entries
    const bool owner = tid < nw * 32;
    double mx = 0.0, my = 0.0;
    for (int it = 0; it < p.rounds; it++) {
        if (it < 10) MS3_TRACE(4 + it);
        double* wp = wpart + (size_t)(it & 1) * MS3_WARPS * 3;
        if (owner) {
            double sx = 0.0, sy = 0.0, sp = 0.0;
            for (int e = tid; e < n; e += 2 * MS3_THREADS) {                  // two entries per trip: independent exp chains
                const int e2 = e + MS3_THREADS;
                const bool v2 = e2 < n;
                const uint32_t ca = entries[e], cb = v2 ? entries[e2] : 0u;
                const double cxa = (double)(ca & 0xffffu), cya = (double)(ca >> 16);
                const double cxb = (double)(cb & 0xffffu), cyb = (double)(cb >> 16);
                if (it == 0) {                                               // mean_shift.cu:31-34
                    const double wb = v2 ? 1.0 : 0.0;
                    sx += cxa + cxb * wb;
                    sy += cya + cyb * wb;
                    sp += 1.0 + wb;
                } else {                                                     // mean_shift.cu:36-46
                    const double dxa = cxa - mx, dya = cya - my, dxb = cxb - mx, dyb = cyb - my;
                    const double pa = ms_exp_nonpos((dxa * dxa + dya * dya) * nis);
                    double pb = ms_exp_nonpos((dxb * dxb + dyb * dyb) * nis);
                    pb = v2 ? pb : 0.0;
                    sx += dxa * pa + dxb * pb;
                    sy += dya * pa + dyb * pb;
                    sp += pa + pb;
                }
            }
            if (it == 1) MS3_TRACE(20);
            sx = ms_warp_sum(sx);
            sy = ms_warp_sum(sy);
            sp = ms_warp_sum(sp);
            if (it == 1) MS3_TRACE(21);
            if (lane == 0) {
                wp[warp * 3 + 0] = sx;
                wp[warp * 3 + 1] = sy;
                wp[warp * 3 + 2] = sp;
            }
        }
        __syncthreads();
        double a = 0.0, b = 0.0, c = 0.0;
        if (R > 1) {
            double* buf = xpart + (size_t)(it & 1) * MS3_MAX_R * 3;
            if (tid < 3) {                                                   // CTA total of one component, warps in order
                double t = 0.0;
                for (int wv = 0; wv < nw; wv++) t += wp[wv * 3 + tid];
                for (int r = 0; r < R; r++) cluster.map_shared_rank(buf, r)[rank * 3 + tid] = t;
            }
            if (it == 1) MS3_TRACE(22);
            cluster.sync();
            if (it == 1) MS3_TRACE(23);
            if (owner) {
#pragma unroll
                for (int r = 0; r < MS3_MAX_R; r++) {                        // unrolled: all loads in flight at once, sums in rank order
                    if (r < R) {
                        a += buf[r * 3 + 0];
                        b += buf[r * 3 + 1];
                        c += buf[r * 3 + 2];
                    }
                }
            }
        } else if (owner) {
            for (int wv = 0; wv < nw; wv++) {
                a += wp[wv * 3 + 0];
                b += wp[wv * 3 + 1];
                c += wp[wv * 3 + 2];
            }
        }
        if (owner || tid == 0) {
            if (R > 1 && !owner) {                                           // tid 0 of a CTA without entries still reports the mean
                const double* buf = xpart + (size_t)(it & 1) * MS3_MAX_R * 3;
                for (int r = 0; r < R; r++) {
                    a += buf[r * 3 + 0];
                    b += buf[r * 3 + 1];
                    c += buf[r * 3 + 2];
                }
            }
            const double inv = 1.0 / c;                                      // mean_shift.py:53-55 (0/0 -> NaN): one divide for x and y
            mx += a * inv;
            my += b * inv;
        }
        if (it == 1) MS3_TRACE(24);
    }
    MS3_TRACE(14);
    if (rank == 0 && tid == 0) {
        p.means_out[2 * ck] = mx;
        p.means_out[2 * ck + 1] = my;
        if (p.with_fingertips) {                                             // src/3d_bz.py:503-522 for the fingertips of this class
            if (p.ft.means_copy) {
                p.ft.means_copy[2 * ck] = mx;
                p.ft.means_copy[2 * ck + 1] = my;
            }
            for (int j = 0; j < p.ft.n; j++)
                if (p.ft.idx[j] == k + 1) p.ft.z_out[(size_t)img * p.ft.n + j] = rf_fingertip_eval(p.ft, mx, my);
        }
    }
    if (R > 1) cluster.sync();   // no CTA may exit while peers can still address its shared memory
}

static size_t ms3_smem_bytes(int chunk) {
    return sizeof(double) * (2 * MS3_MAX_R * 3 + 2 * MS3_WARPS * 3) + sizeof(int) * MS3_WARPS + sizeof(uint32_t) * (size_t)chunk;
}


static size_t ms_smem_bytes(int K, int NC) {
    size_t b = 0;
    b += sizeof(double) * 2 * (size_t)NC * K * 3;
    b += sizeof(double) * MS_MAX_ITEMS * 3;
    b += sizeof(double) * 3 * (size_t)K;
    b += sizeof(int) * ((size_t)MS_WARPS * K + 2 * (K + 1));
    return b;
}

extern "C" int rdf_mean_shift_workspace_bytes(int dim_x, int dim_y, int num_labels, size_t* bytes) {
    RDF_REQUIRE(bytes != nullptr && dim_x > 0 && dim_y > 0 && num_labels >= 0, "rdf_mean_shift_workspace_bytes: bad argument");
    // one uint32 per pixel, padded so every CTA slice (multiple of 512 pixels) stays in bounds
    *bytes = sizeof(uint32_t) * ((size_t)dim_x * dim_y + (size_t)MS_MAX_CLUSTER * 32 * MS_WARPS);
    return RDF_OK;
}

// defined in rdf_frame.cu
int rdf_fingertip_fill_spec(rf_fingertip_spec* ft, const char* who, const int* fingertip_labels, int num_fingertips, int labels_reduce,
                            const uint16_t* raw_depth_dev, int dim_x, int dim_y, float ppx, float ppy, float fx, float fy,
                            const float* plane_dev, double* z_out, double* means_copy_out);

// returns RDF_OK and sets *fused_done when the fused read-out ran inside the launch
static int rdf_mean_shift_impl(const uint16_t* labels_dev, int num_images, int dim_x, int dim_y, int num_labels, const float* variances_dev,
                               int rounds, double* means_dev, void* workspace_dev, size_t workspace_bytes, void* stream,
                               const rf_fingertip_spec* ft = nullptr, bool* fused_done = nullptr) {
    RDF_REQUIRE(labels_dev && variances_dev && means_dev && workspace_dev, "rdf_mean_shift: NULL argument");
    RDF_REQUIRE(num_images >= 1 && num_images <= 64, "rdf_mean_shift: num_images=%d outside 1..64", num_images);
    RDF_REQUIRE(dim_x > 0 && dim_y > 0 && dim_x <= 65535 && dim_y <= 65535, "rdf_mean_shift: bad image shape %dx%d", dim_x, dim_y);
    RDF_REQUIRE(num_labels >= 1 && num_labels <= RDF_MAX_CLASSES, "rdf_mean_shift: num_labels=%d outside 1..%d", num_labels,
                RDF_MAX_CLASSES);
    RDF_REQUIRE(rounds >= 0, "rdf_mean_shift: rounds=%d", rounds);
    size_t need = 0;
    rdf_mean_shift_workspace_bytes(dim_x, dim_y, num_labels, &need);
    RDF_REQUIRE(workspace_bytes >= need, "rdf_mean_shift: workspace %zu < %zu bytes", workspace_bytes, need);
    RDF_REQUIRE((int64_t)dim_x * dim_y < (1LL << 30), "rdf_mean_shift: image too large");

    const int npx = dim_x * dim_y;
    const bool v3_ok = npx <= MS3_MAX_R * MS3_CAP && (reinterpret_cast<uintptr_t>(labels_dev) & 15u) == 0 &&
                       !RDF_GETENV_ONCE("RDF_MS_V1");            // RDF_MS_V1: force the fallback (tests / experiments)
    if (num_images > 1 && !(v3_ok && (npx & 7) == 0)) {
        // no batched form of the other paths: one launch per image (same results, the workspace is reused in stream order)
        for (int n = 0; n < num_images; n++) {
            const int rc = rdf_mean_shift_impl(labels_dev + (size_t)n * npx, 1, dim_x, dim_y, num_labels, variances_dev, rounds,
                                               means_dev + (size_t)n * num_labels * 2, workspace_dev, workspace_bytes, stream, nullptr,
                                               nullptr);
            if (rc != RDF_OK) return rc;
        }
        return RDF_OK;
    }
    // class-parallel latency path (v3 above): one small cluster per class (and per image)
    if (v3_ok) {
        // CTAs per class: enough that a CTA scans at most ~12 800 pixels (52 KB of shared memory: such CTAs can be scheduled
        // beside the still-running layered kernel under programmatic dependent launch, and the scan of the label image is
        // spread over more SMs).  Measured on cfg2 (424 x 240 labels): R = 2 / 4 / 8 -> 28.7 / 26.6 / 22.6-24.6 us per launch,
        // e2e frame latency 74.7 / 72.7 / 68.2-69.5 us.  RDF_MS3_R overrides the minimum (experiments).
        int R = (npx + MS3_CAP - 1) / MS3_CAP;
        {
            static int r_min = -1;
            if (r_min < 0) {
                const char* e = RDF_GETENV_ONCE("RDF_MS3_R");
                r_min = e ? atoi(e) : 0;
                if (r_min < 0 || r_min > MS3_MAX_R) r_min = 0;
            }
            int want = r_min ? r_min : (npx + 12799) / 12800;
            if (want > MS3_MAX_R) want = MS3_MAX_R;
            if (R < want) R = want;
        }
        if (R == 3) R = 4;                                            // cluster sizes: 1, 2, 4, 8
        if (R > 4 && R < 8) R = 8;
        // a CTA of this kernel owns an SM: keep all clusters of a batch co-resident when the image capacity allows it (two hands
        // x 11 classes x 8 CTAs would queue behind each other on 148 SMs; measured 114 -> 106 us per product frame with R = 4)
        {
            const int r_cap = (npx + MS3_CAP - 1) / MS3_CAP;
            const long long clusters = (long long)num_images * num_labels;
            if (clusters * R > rdf_sm_count()) {                                   // any cluster size works (measured 3 / 4 / 5 / 6 CTAs: 100.8 / 99.1 /
                int fit = (int)(rdf_sm_count() / clusters);                        // 98.9 / 98.2 us per product frame)
                if (fit < r_cap) fit = r_cap;
                if (fit < 1) fit = 1;
                if (fit < R) R = fit;
            }
            const char* e = RDF_GETENV_ONCE("RDF_MS3_RFINAL");                   // experiments: any cluster size 1..8 that holds the image
            if (e && atoi(e) >= r_cap && atoi(e) >= 1 && atoi(e) <= MS3_MAX_R) R = atoi(e);
        }
        rdf_ms3_params q;
        q.labels = labels_dev; q.variances = variances_dev; q.means_out = means_dev;
        q.w = dim_x; q.h = dim_y; q.K = num_labels; q.rounds = rounds; q.R = R;
        q.trace = RDF_GETENV_ONCE("RDF_MS_TRACE") ? reinterpret_cast<unsigned long long*>(workspace_dev) : nullptr;
        q.with_fingertips = 0;
        if (ft) {
            q.with_fingertips = 1;
            q.ft = *ft;
            if (fused_done) *fused_done = true;
        }
        const int ngroups = (npx + 7) / 8;
        const int chunk = ((ngroups + R - 1) / R) * 8;                // pixels (= upper bound of entries) per CTA
        const size_t smem3 = ms3_smem_bytes(chunk);
        RDF_ENSURE_DYN_SMEM(rdf_mean_shift_v3_kernel, smem3);
        cudaLaunchConfig_t cfg3 = {};
        cfg3.gridDim = dim3(num_images * num_labels * R, 1, 1);
        cfg3.blockDim = dim3(MS3_THREADS, 1, 1);
        cfg3.dynamicSmemBytes = smem3;
        cfg3.stream = rdf_stream(stream);
        cudaLaunchAttribute attr3[2];
        attr3[0].id = cudaLaunchAttributeClusterDimension;
        attr3[0].val.clusterDim.x = R;
        attr3[0].val.clusterDim.y = 1;
        attr3[0].val.clusterDim.z = 1;
        attr3[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr3[1].val.programmaticStreamSerializationAllowed = 1;
        cfg3.attrs = attr3;
        cfg3.numAttrs = RDF_GETENV_ONCE("RDF_NO_PDL") ? 1 : 2;
        RDF_CUDA(cudaLaunchKernelEx(&cfg3, rdf_mean_shift_v3_kernel, q));
        return RDF_OK;
    }
    const int gran = 32 * MS_WARPS;
    int NC = (npx + 8191) / 8192;                  // >= 8192 pixels per CTA before adding CTAs
    if (NC > MS_MAX_CLUSTER) NC = MS_MAX_CLUSTER;
    if (NC < 1) NC = 1;
    rdf_ms_params p;
    p.labels = labels_dev;
    p.variances = variances_dev;
    p.means_out = means_dev;
    p.entries = reinterpret_cast<uint32_t*>(workspace_dev);
    p.w = dim_x; p.h = dim_y; p.K = num_labels; p.rounds = rounds;
    p.chunk = (((npx + NC - 1) / NC) + gran - 1) / gran * gran;
    int item = (p.chunk + (MS_MAX_ITEMS - RDF_MAX_CLASSES) - 1) / (MS_MAX_ITEMS - RDF_MAX_CLASSES);
    item = (item + 31) / 32 * 32;
    if (item < 256) item = 256;
    p.item = item;

    const size_t smem = ms_smem_bytes(num_labels, NC);
    RDF_ENSURE_DYN_SMEM(rdf_mean_shift_kernel, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NC, 1, 1);
    cfg.blockDim = dim3(MS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = rdf_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_mean_shift_kernel, p));
    return RDF_OK;
}

extern "C" int rdf_mean_shift(const uint16_t* labels_dev, int dim_x, int dim_y, int num_labels, const float* variances_dev,
                              int rounds, double* means_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    return rdf_mean_shift_impl(labels_dev, 1, dim_x, dim_y, num_labels, variances_dev, rounds, means_dev, workspace_dev, workspace_bytes,
                               stream);
}

extern "C" int rdf_mean_shift_batch(const uint16_t* labels_dev, int num_images, int dim_x, int dim_y, int num_labels,
                                    const float* variances_dev, int rounds, double* means_dev, void* workspace_dev,
                                    size_t workspace_bytes, void* stream) {
    return rdf_mean_shift_impl(labels_dev, num_images, dim_x, dim_y, num_labels, variances_dev, rounds, means_dev, workspace_dev,
                               workspace_bytes, stream);
}

extern "C" int rdf_mean_shift_fingertips(const uint16_t* labels_dev, int num_images, int dim_x, int dim_y, int num_labels,
                                         const float* variances_dev, int rounds, double* means_dev, void* workspace_dev,
                                         size_t workspace_bytes, const int* fingertip_labels, int num_fingertips, int labels_reduce,
                                         const uint16_t* raw_depth_dev, int raw_dim_x, int raw_dim_y, float ppx, float ppy, float fx,
                                         float fy, const float* plane_dev, double* z_out, double* means_copy_out, void* stream) {
    rf_fingertip_spec ft;
    int rc = rdf_fingertip_fill_spec(&ft, "rdf_mean_shift_fingertips", fingertip_labels, num_fingertips, labels_reduce, raw_depth_dev,
                                     raw_dim_x, raw_dim_y, ppx, ppy, fx, fy, plane_dev, z_out, means_copy_out);
    if (rc != RDF_OK) return rc;
    // a fingertip id outside 1..num_labels has no class cluster to write it: handled by the separate read-out
    bool all_inside = true;
    for (int i = 0; i < num_fingertips; i++) all_inside = all_inside && fingertip_labels[i] >= 1 && fingertip_labels[i] <= num_labels;
    bool fused = false;
    rc = rdf_mean_shift_impl(labels_dev, num_images, dim_x, dim_y, num_labels, variances_dev, rounds, means_dev, workspace_dev,
                             workspace_bytes, stream, all_inside ? &ft : nullptr, &fused);
    if (rc != RDF_OK || fused) return rc;
    return rdf_fingertip_z(means_dev, num_images, num_labels, fingertip_labels, num_fingertips, labels_reduce, raw_depth_dev, raw_dim_x,
                           raw_dim_y, ppx, ppy, fx, fy, plane_dev, z_out, means_copy_out, stream);
}
