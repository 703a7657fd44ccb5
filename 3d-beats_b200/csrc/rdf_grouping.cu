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

// read-only find for the flatten pass: there every thread overwrites its OWN entry with the root, so a concurrent path-halving
// write of another thread (an ancestor that is not the root) landing afterwards would undo it - seen as a few pixels missing from
// the stencil on tall narrow images, where chains are long (tools/stress_grouping.py)
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

// Phases (one CTA):
//   1. a bit mask of the non-zero pixels per row (warp ballots), then one thread per PIXEL: the start of its horizontal run from
//      the mask (count-leading-zeros over at most a few words); every pixel of a run points at the run's first pixel, which holds
//      the run's length and x-sum (so later statistics cost one atomic per run, not per pixel: a 1500-pixel hand blob would
//      otherwise serialise 1500 shared-memory atomics on one address);
//   2. vertical unions between runs of neighbouring rows (only where an overlap segment starts);
//   3. run statistics are added to their component's root (path-halving finds shorten every chain);
//   5. selection over roots; 6. y-sums of the two selected components (warp-reduced, one atomic per warp); 7. stencil + g_info.
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
    int* parent = gr_smem;              // [N]  -1 = background
    int* cnt = parent + N;              // [N]  run length at run starts, then component size at roots
    int* sumx = cnt + N;                // [N]  run x-sum at run starts, then component x-sum at roots
    unsigned* rowmask = reinterpret_cast<unsigned*>(sumx + N);   // [h * wpr]  bit x & 31 of word x >> 5: pixel x of the row is non-zero
    __shared__ unsigned long long best[2];      // per side: size << 32 | ~root
    __shared__ int sumy[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // 1a. masks (grouping.cpp:107-108: non-zero = foreground).  Eight words per warp and trip with all loads issued before the
    // first ballot: the image comes out of L2 / HBM, one latency per trip instead of one per word
    for (int t0 = warp; t0 < h * wpr; t0 += 8 * (GR_THREADS / 32)) {
        unsigned v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int t = t0 + j * (GR_THREADS / 32);
            const int y = t / wpr, x = (t - y * wpr) * 32 + lane;
            v[j] = (t < h * wpr && x < w) ? (unsigned)__ldg(img + y * w + x) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int t = t0 + j * (GR_THREADS / 32);
            const unsigned m = __ballot_sync(0xffffffffu, v[j] != 0u);
            if (lane == 0 && t < h * wpr) rowmask[t] = m;
        }
    }
    if (tid < 2) {
        best[tid] = 0ull;
        sumy[tid] = 0;
    }
    __syncthreads();
    GR_TRACE(1);
    for (int i = tid; i < N; i += GR_THREADS) {                    // 1b. runs
        const int y = i / w, x = i - y * w;
        const unsigned* rm = rowmask + y * wpr;
        const int wi = x >> 5, bi = x & 31;
        int par = -1, len = 0, sx = 0;
        if ((rm[wi] >> bi) & 1u) {
            // start of my run: one past the last background pixel to my left
            int start = 0;
            unsigned z = ~rm[wi] & ((1u << bi) - 1u);
            for (int wj = wi;;) {
                if (z) {
                    start = wj * 32 + 32 - __clz(z);
                    break;
                }
                if (--wj < 0) break;
                z = ~rm[wj];
            }
            par = y * w + start;
            if (start == x) {                                      // first pixel of the run: its end = first background pixel to my right
                int end = w;
                unsigned z2 = ~rm[wi] & ~((2u << bi) - 1u);        // bits above bi (2u << 31 wraps to 0: mask becomes all ones -> none)
                if (bi == 31) z2 = 0u;
                for (int wj = wi;;) {
                    if (z2) {
                        end = wj * 32 + __ffs(z2) - 1;
                        break;
                    }
                    if (++wj >= wpr) break;
                    z2 = ~rm[wj];
                }
                if (end > w) end = w;                              // mask bits beyond the row are background
                len = end - start;
                sx = (start + end - 1) * len / 2;                  // start + ... + (end - 1)
            }
        }
        parent[i] = par;
        cnt[i] = len;
        sumx[i] = sx;
    }
    __syncthreads();
    GR_TRACE(2);
    for (int i = tid; i < N - w; i += GR_THREADS) {                // 2. 4-connectivity (grouping.cpp:82-87): down edges between runs
        if (parent[i] < 0 || parent[i + w] < 0) continue;
        const int x = i % w;
        if (x > 0 && parent[i - 1] >= 0 && parent[i + w - 1] >= 0) continue;   // same pair of runs as the pixel to the left
        gr_unite(parent, i, i + w);
    }
    __syncthreads();
    GR_TRACE(3);
    for (int i = tid; i < N; i += GR_THREADS) {                    // 3. run statistics -> root (roots keep their own run in place)
        const int len = cnt[i];
        if (len == 0) continue;
        const int r = gr_find(parent, i);
        if (r != i) {
            atomicAdd(&cnt[r], len);
            atomicAdd(&sumx[r], sumx[i]);
        }
    }
    __syncthreads();
    GR_TRACE(4);
    // (4. no flatten pass: the stencil pass below walks to the root itself; after the path-halving finds of pass 3 that is a hop or
    //  two, and nothing writes `parent` any more, so no pass can undo another's result - see gr_find_ro)
    for (int i = tid; i < N; i += GR_THREADS) {                    // 5. selection
        if (parent[i] != i) continue;                              // roots only, one per component
        const int n = cnt[i];
        if (__fdiv_rn((float)n, (float)N) <= pct_thresh) continue;            // grouping.cpp:137
        const float cx = __fdiv_rn((float)sumx[i], (float)n);                 // grouping.cpp:148
        const int side = cx < __fdiv_rn((float)w, 2.f) ? 0 : 1;              // grouping.cpp:150
        // strictly largest wins, so among equal sizes the component met first in raster order = smallest root (grouping.cpp:151,157)
        atomicMax(&best[side], ((unsigned long long)(unsigned)n << 32) | (unsigned long long)(0xffffffffu - (unsigned)i));
    }
    __syncthreads();
    GR_TRACE(6);
    int sel[2], seln[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        seln[s] = (int)(best[s] >> 32);
        sel[s] = seln[s] ? (int)(0xffffffffu - (unsigned)(best[s] & 0xffffffffull)) : -2;
    }
    for (int i0 = 0; i0 < N; i0 += GR_THREADS) {                   // 6. y-sums + 7. stencil (src/3d_bz.py:243-250)
        const int i = i0 + tid;
        const int r = (i < N && parent[i] >= 0) ? gr_find_ro(parent, i) : -1;
        const int y = i / w;
        const int c0 = __reduce_add_sync(0xffffffffu, r == sel[0] ? y : 0);
        const int c1 = __reduce_add_sync(0xffffffffu, r == sel[1] ? y : 0);
        if (lane == 0) {
            if (c0) atomicAdd(&sumy[0], c0);
            if (c1) atomicAdd(&sumy[1], c1);
        }
        if (i < N) stencil[i] = (unsigned short)(r < 0 ? 0 : r == sel[0] ? 1 : r == sel[1] ? 2 : 0);
    }
    __syncthreads();
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
    cfg.numAttrs = getenv("RDF_NO_PDL") ? 0 : 1;
    static long long* trace_dev = nullptr;                         // debug aid: RDF_GR_TRACE=1 prints clock64 deltas per phase (synchronises)
    const bool tracing = getenv("RDF_GR_TRACE") != nullptr;
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
