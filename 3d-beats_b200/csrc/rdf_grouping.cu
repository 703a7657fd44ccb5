// Hand grouping on the device: replaces the reference's D2H copy -> C++ BFS flood fill -> H2D copy -> scatter kernel
// (src/3d_bz.py:222-250, src/cpp_grouping/grouping.cpp:80-191, write_pixel_groups_to_stencil_image) by ONE launch on the
// 1/8-resolution depth image (106 x 60 pixels for the product): 4-connected components of the non-zero pixels with a
// shared-memory union-find whose roots are the smallest raster index of each component, component sizes and x-sums by
// shared-memory atomics, then the reference's selection rule - components above the size threshold, centroid x < w/2 ->
// "right" candidate else "left", per side the largest component, ties -> the one met first in raster order (= smallest root).
// Output is the stencil image the reference builds from the coordinate list (1 = right group, 2 = left group, 0 elsewhere) and
// g_info[2][3] = (size, centroid x, centroid y).  SURVEY 8(f) rank 2: the only native C++ component of the reference and the
// only mandatory host round trip of the live frame.
#include "rdf_common.cuh"

#include <stdlib.h>

#define GR_THREADS 1024
#define GR_MAX_PIXELS 16384

// find with path halving: chains stay short although roots are chosen by smallest index, not by rank
__device__ __forceinline__ int gr_find(volatile int* parent, int i) {
    int p = parent[i];
    while (p != i) {
        const int gp = parent[p];
        if (gp != p) parent[i] = gp;        // benign race: any ancestor is a valid parent
        i = p;
        p = gp;
    }
    return i;
}

// read-only find for the passes in which other threads must not see their results undone: a path-halving write (an ancestor that
// is not the root) landing after a thread had recorded its root showed up as a few pixels missing from the stencil on tall narrow
// images, where chains are long (tools/stress_grouping.py)
__device__ __forceinline__ int gr_find_ro(const volatile int* parent, int i) {
    int p = parent[i];
    while (p != i) {
        i = p;
        p = parent[i];
    }
    return i;
}

__device__ __forceinline__ void gr_unite(int* parent, int a, int b) {
    bool done;
    do {
        a = gr_find(parent, a);
        b = gr_find(parent, b);
        if (a < b) {
            const int old = atomicMin(&parent[b], a);
            done = old == b;
            b = old;
        } else if (b < a) {
            const int old = atomicMin(&parent[a], b);
            done = old == a;
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// ---- row bit masks: bit x & 31 of word x >> 5 of a row is set when pixel x is non-zero; bits beyond the row are clear -----------
// first pixel of the horizontal run that contains foreground pixel x
__device__ __forceinline__ int gr_run_start(const unsigned* rm, int x) {
    int wj = x >> 5;
    unsigned z = ~rm[wj] & ((1u << (x & 31)) - 1u);              // background pixels to my left in this word
    for (;;) {
        if (z) return wj * 32 + 32 - __clz(z);                   // one past the last of them
        if (--wj < 0) return 0;
        z = ~rm[wj];
    }
}

// one past the last pixel of the run that contains foreground pixel x
__device__ __forceinline__ int gr_run_end(const unsigned* rm, int wpr, int w, int x) {
    int wj = x >> 5;
    const int bi = x & 31;
    unsigned z = bi == 31 ? 0u : (~rm[wj] & ~((2u << bi) - 1u)); // background pixels to my right in this word
    for (;;) {
        if (z) {
            const int e = wj * 32 + __ffs(z) - 1;
            return e < w ? e : w;
        }
        if (++wj >= wpr) return w;
        z = ~rm[wj];
    }
}

// first pixels of the runs that begin in word wi of a row
__device__ __forceinline__ unsigned gr_starts(const unsigned* rm, int wi) {
    const unsigned m = rm[wi];
    return m & ~((m << 1) | (wi > 0 ? rm[wi - 1] >> 31 : 0u));
}

// Everything after the masks works on horizontal RUNS (a few hundred on a product frame) instead of pixels (6360): one thread per
// mask word (row, 32 columns) enumerates the runs that begin in its word.  The first version did every pass per pixel and was bound
// by the issue rate of the one SM it runs on (58 k warp instructions, 17 us).  Union-find entries exist only at the first pixel of a
// run (indexed by pixel, so the smallest root is the component's first pixel in raster order - the reference's tie rule).
//   0. masks (warp ballots over the image);
//   1. per run: parent = itself, cnt = length, sumx = sum of its x;
//   2. per run: unions with the runs of the next row it touches (one per overlap segment);
//   3. per run: statistics added to the root (path-halving finds shorten every chain);
//   5. per root: the reference's selection (size threshold, side by centroid x, largest wins, ties -> first in raster order);
//   6. per run: label (1 right / 2 left / 0) into cnt, y-sums of the selected components;
//   7. per pixel: stencil = label of its run; g_info.
#define GR_TRACE(slot)                                                    \
    do {                                                                  \
        if (trace && threadIdx.x == 0) trace[slot] = clock64();           \
    } while (0)

__global__ void __launch_bounds__(GR_THREADS) rdf_group_hands_kernel(const uint16_t* __restrict__ img, int w, int h, float pct_thresh,
                                                                     uint16_t* __restrict__ stencil, float* __restrict__ g_info,
                                                                     long long* __restrict__ trace /* debug: RDF_GR_TRACE */) {
    GR_TRACE(0);
    // the kernel that consumes the stencil may be scheduled now (it waits for this grid itself); this grid may have been
    // scheduled early behind the kernel that produces img
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    GR_TRACE(15);
    extern __shared__ int gr_smem[];
    const int N = w * h;
    const int wpr = (w + 31) >> 5;      // mask words per row
    const int nwords = h * wpr;
    int* parent = gr_smem;              // [N]  valid at run starts only
    int* cnt = parent + N;              // [N]  run length at run starts, component size at roots; from pass 6 on: the run's label
    int* sumx = cnt + N;                // [N]  run x-sum at run starts, component x-sum at roots
    unsigned* rowmask = reinterpret_cast<unsigned*>(sumx + N);   // [h * wpr]
    __shared__ unsigned long long best[2];      // per side: size << 32 | ~root
    __shared__ int sumy[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // 0. masks (grouping.cpp:107-108: non-zero = foreground).  Eight words per warp and trip with all loads issued before the
    // first ballot: the image comes out of L2 / HBM, one latency per trip instead of one per word
    for (int t0 = warp; t0 < nwords; t0 += 8 * (GR_THREADS / 32)) {
        unsigned v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int t = t0 + j * (GR_THREADS / 32);
            const int y = t / wpr, x = (t - y * wpr) * 32 + lane;
            v[j] = (t < nwords && x < w) ? (unsigned)__ldg(img + y * w + x) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int t = t0 + j * (GR_THREADS / 32);
            const unsigned m = __ballot_sync(0xffffffffu, v[j] != 0u);
            if (lane == 0 && t < nwords) rowmask[t] = m;
        }
    }
    if (tid < 2) {
        best[tid] = 0ull;
        sumy[tid] = 0;
    }
    __syncthreads();
    GR_TRACE(1);
    for (int t = tid; t < nwords; t += GR_THREADS) {               // 1. runs
        const int y = t / wpr, wi = t - y * wpr;
        const unsigned* rm = rowmask + y * wpr;
        for (unsigned st = gr_starts(rm, wi); st; st &= st - 1) {
            const int x = wi * 32 + __ffs(st) - 1;
            const int len = gr_run_end(rm, wpr, w, x) - x;
            const int i = y * w + x;
            parent[i] = i;
            cnt[i] = len;
            sumx[i] = (2 * x + len - 1) * len / 2;                 // x + ... + (x + len - 1)
        }
    }
    __syncthreads();
    GR_TRACE(2);
    for (int t = tid; t < nwords - wpr; t += GR_THREADS) {         // 2. 4-connectivity (grouping.cpp:82-87): runs of row y with row y + 1
        const int y = t / wpr, wi = t - y * wpr;
        const unsigned* rm = rowmask + y * wpr;
        const unsigned* rb = rm + wpr;
        for (unsigned st = gr_starts(rm, wi); st; st &= st - 1) {
            const int x = wi * 32 + __ffs(st) - 1;
            const int end = gr_run_end(rm, wpr, w, x);
            for (int wj = wi; wj * 32 < end; wj++) {
                const int lo = max(x - wj * 32, 0), hi = min(end - wj * 32, 32);            // my columns in word wj: bits [lo, hi)
                const unsigned rng = (hi == 32 ? 0xffffffffu : (1u << hi) - 1u) & ~((1u << lo) - 1u);
                const unsigned below = rb[wj];
                const unsigned bprev = (below << 1) | (wj > 0 ? rb[wj - 1] >> 31 : 0u);
                // one union per overlap segment: a pixel below me whose left neighbour is background, or the one under my first pixel
                unsigned cand = below & rng & (~bprev | (wj == wi ? 1u << (x & 31) : 0u));
                for (; cand; cand &= cand - 1) {
                    const int xb = wj * 32 + __ffs(cand) - 1;
                    gr_unite(parent, y * w + x, (y + 1) * w + gr_run_start(rb, xb));
                }
            }
        }
    }
    __syncthreads();
    GR_TRACE(3);
    for (int t = tid; t < nwords; t += GR_THREADS) {               // 3. run statistics -> root (roots keep their own run in place)
        const int y = t / wpr, wi = t - y * wpr;
        for (unsigned st = gr_starts(rowmask + y * wpr, wi); st; st &= st - 1) {
            const int i = y * w + wi * 32 + __ffs(st) - 1;
            const int r = gr_find(parent, i);
            if (r != i) {
                atomicAdd(&cnt[r], cnt[i]);
                atomicAdd(&sumx[r], sumx[i]);
            }
        }
    }
    __syncthreads();
    GR_TRACE(4);
    for (int t = tid; t < nwords; t += GR_THREADS) {               // 5. selection
        const int y = t / wpr, wi = t - y * wpr;
        for (unsigned st = gr_starts(rowmask + y * wpr, wi); st; st &= st - 1) {
            const int i = y * w + wi * 32 + __ffs(st) - 1;
            if (parent[i] != i) continue;                          // roots only, one per component
            const int n = cnt[i];
            if (__fdiv_rn((float)n, (float)N) <= pct_thresh) continue;            // grouping.cpp:137
            const float cx = __fdiv_rn((float)sumx[i], (float)n);                 // grouping.cpp:148
            const int side = cx < __fdiv_rn((float)w, 2.f) ? 0 : 1;              // grouping.cpp:150
            // strictly largest wins, so among equal sizes the component met first in raster order = smallest root (grouping.cpp:151,157)
            atomicMax(&best[side], ((unsigned long long)(unsigned)n << 32) | (unsigned long long)(0xffffffffu - (unsigned)i));
        }
    }
    __syncthreads();
    GR_TRACE(5);
    int sel[2], seln[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        seln[s] = (int)(best[s] >> 32);
        sel[s] = seln[s] ? (int)(0xffffffffu - (unsigned)(best[s] & 0xffffffffull)) : -2;
    }
    for (int t = tid; t < nwords; t += GR_THREADS) {               // 6. labels of the runs (into cnt, which nobody needs any more) + y-sums
        const int y = t / wpr, wi = t - y * wpr;
        const unsigned* rm = rowmask + y * wpr;
        for (unsigned st = gr_starts(rm, wi); st; st &= st - 1) {
            const int x = wi * 32 + __ffs(st) - 1;
            const int i = y * w + x;
            const int r = gr_find_ro(parent, i);
            const int lab = r == sel[0] ? 1 : r == sel[1] ? 2 : 0;
            if (lab) atomicAdd(&sumy[lab - 1], y * (gr_run_end(rm, wpr, w, x) - x));
            cnt[i] = lab;
        }
    }
    __syncthreads();
    GR_TRACE(6);
    for (int t = warp; t < nwords; t += GR_THREADS / 32) {         // 7. stencil (src/3d_bz.py:243-250): a warp per mask word
        const int y = t / wpr, wi = t - y * wpr;
        const unsigned* rm = rowmask + y * wpr;
        const int x = wi * 32 + lane;
        unsigned short v = 0;
        if ((rm[wi] >> lane) & 1u) v = (unsigned short)cnt[y * w + gr_run_start(rm, x)];
        if (x < w) stencil[y * w + x] = v;
    }
    GR_TRACE(7);
    if (tid < 2) {
        const int s = tid;
        float n = 0.f, cx = 0.f, cy = 0.f;
        if (seln[s]) {
            n = (float)seln[s];
            cx = __fdiv_rn((float)sumx[sel[s]], n);
            cy = __fdiv_rn((float)sumy[s], n);                                // grouping.cpp:147
        }
        g_info[3 * s + 0] = n;
        g_info[3 * s + 1] = cx;
        g_info[3 * s + 2] = cy;
    }
    GR_TRACE(8);
}

extern "C" int rdf_group_hands(const uint16_t* img_dev, int dim_x, int dim_y, float pct_thresh, uint16_t* stencil_dev,
                               float* g_info_dev, void* stream) {
    RDF_REQUIRE(img_dev && stencil_dev && g_info_dev, "rdf_group_hands: NULL argument");
    RDF_REQUIRE(dim_x > 0 && dim_y > 0, "rdf_group_hands: bad shape %dx%d", dim_x, dim_y);
    const int64_t n = (int64_t)dim_x * dim_y;
    if (n > GR_MAX_PIXELS) {
        rdf_set_error("rdf_group_hands: %dx%d exceeds the %d pixels of the shared-memory union-find (the product runs it on the "
                      "1/8-resolution image)", dim_x, dim_y, GR_MAX_PIXELS);
        return RDF_ERR_UNSUPPORTED;
    }
    const size_t smem = sizeof(int) * (3 * (size_t)n + (size_t)dim_y * ((dim_x + 31) / 32));
    if (smem > 227 * 1024) {                                       // only degenerate shapes (16384 rows of one pixel)
        rdf_set_error("rdf_group_hands: %dx%d needs %zu bytes of shared memory", dim_x, dim_y, smem);
        return RDF_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024) RDF_ENSURE_DYN_SMEM(rdf_group_hands_kernel, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, 1, 1);
    cfg.blockDim = dim3(GR_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = rdf_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = RDF_GETENV_ONCE("RDF_NO_PDL") ? 0 : 1;
    static long long* trace_dev = nullptr;                         // debug aid: RDF_GR_TRACE=1 prints clock64 deltas per phase (synchronises)
    const bool tracing = RDF_GETENV_ONCE("RDF_GR_TRACE") != nullptr;
    if (tracing && !trace_dev) RDF_CUDA(cudaMalloc(&trace_dev, 16 * sizeof(long long)));
    long long* trace_arg = tracing ? trace_dev : nullptr;
    RDF_CUDA(cudaLaunchKernelEx(&cfg, rdf_group_hands_kernel, img_dev, dim_x, dim_y, pct_thresh, stencil_dev, g_info_dev, trace_arg));
    if (tracing) {
        long long t[16];
        RDF_CUDA(cudaMemcpy(t, trace_dev, sizeof(t), cudaMemcpyDeviceToHost));
        fprintf(stderr, "rdf_group_hands phases (cycles): wait %lld |", t[15] - t[0]);
        for (int i = 1; i <= 8; i++) fprintf(stderr, " %lld", t[i] - (i == 1 ? t[15] : t[i - 1]));
        fprintf(stderr, "\n");
    }
    return RDF_OK;
}
