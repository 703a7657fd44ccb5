// Training split search: replaces evaluate_random_features, pick_best_features, get_active_nodes_next_level and
// copy_pixel_groups (reference src/cuda/tree_train.cu:4-64, :66-236, :238-273, :275-324) plus the host-side root
// statistics of DecisionTreeTrainer.train (src/decision_tree.py:452-467).
#include <string.h>

#include "rdf_common.cuh"

// ---------------------------------------------------------------------------------------------------------------
// root statistics
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rdf_train_init_kernel(const uint16_t* __restrict__ labels, int64_t n, int C,
                                                             int32_t* __restrict__ nodes_by_pixel,
                                                             unsigned long long* __restrict__ root_counts) {
    extern __shared__ unsigned int cnt_s[];
    for (int c = threadIdx.x; c < C; c += blockDim.x) cnt_s[c] = 0u;
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned l = __ldg(labels + i);
        // labels outside 1..C-1 cannot be counted (the reference would index node_counts out of bounds); keep them inactive
        const bool active = l > 0u && l < (unsigned)C;
        nodes_by_pixel[i] = active ? 0 : -1;                       // decision_tree.py:462-463
        if (active) atomicAdd(&cnt_s[l], 1u);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        if (cnt_s[c]) atomicAdd(&root_counts[c], (unsigned long long)cnt_s[c]);   // decision_tree.py:457-460
}

extern "C" int rdf_train_init(const uint16_t* labels_dev, int64_t num_pixels, int num_classes, int32_t* nodes_by_pixel_dev,
                              uint64_t* root_counts_dev, void* stream) {
    RDF_REQUIRE(labels_dev && nodes_by_pixel_dev && root_counts_dev, "rdf_train_init: NULL argument");
    RDF_REQUIRE(num_pixels >= 0 && num_classes >= 1 && num_classes <= RDF_MAX_CLASSES, "rdf_train_init: bad argument");
    cudaStream_t st = rdf_stream(stream);
    RDF_CUDA(cudaMemsetAsync(root_counts_dev, 0, sizeof(uint64_t) * num_classes, st));
    if (num_pixels == 0) return RDF_OK;
    rdf_train_init_kernel<<<rdf_sm_count() * 8, 256, sizeof(unsigned) * num_classes, st>>>(
        labels_dev, num_pixels, num_classes, nodes_by_pixel_dev, reinterpret_cast<unsigned long long*>(root_counts_dev));
    RDF_LAUNCH_CHECK("rdf_train_init_kernel");
    return RDF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// split histograms
// ---------------------------------------------------------------------------------------------------------------
// Work decomposition: blockIdx.x = tile of TH_PIX consecutive pixels, blockIdx.y = chunk of FC features.
// Lanes of a warp are consecutive pixels evaluating the SAME feature, so both depth probes of a warp land on
// neighbouring addresses (coalesced gathers), and the feature's offsets / thresholds are shared-memory broadcasts.
// Histogram updates go to a shared-memory privatised copy when all slots x FC features fit (shallow levels, where
// every pixel of the GPU hits the same few nodes), otherwise straight to global memory (deep levels, low contention).
#define TH_THREADS 512
#define TH_PIX 4096          // pixels per block
#define TH_MAX_FC 128

struct rdf_hist_params {
    const uint16_t* depth;
    const uint16_t* labels;
    const int32_t* nodes_by_pixel;
    const int32_t* node_slot;
    const float* offsets;        // [F][4]
    const float* thresholds;     // [F][NT]
    uint32_t* hist;              // [S][F][NB][C]
    int64_t num_pixels;
    int W, H, S, F, NT, NB, C, FC;
    int privatise;               // 1: shared-memory histogram [S][FC][NB][C]
};

__global__ void __launch_bounds__(TH_THREADS) rdf_train_hist_kernel(const rdf_hist_params p) {
    extern __shared__ __align__(16) unsigned char th_smem[];
    float* off_s = reinterpret_cast<float*>(th_smem);                         // [FC][4]
    float* thr_s = off_s + 4 * p.FC;                                          // [FC][NT]
    uint32_t* hist_s = reinterpret_cast<uint32_t*>(thr_s + (size_t)p.FC * p.NT);   // [S][FC][NB][C] if privatise

    const int f0 = blockIdx.y * p.FC;
    const int nf = min(p.FC, p.F - f0);
    for (int i = threadIdx.x; i < nf * 4; i += TH_THREADS) off_s[i] = __ldg(p.offsets + (size_t)f0 * 4 + i);
    __shared__ unsigned char exact_s[TH_MAX_FC];      // feature needs the exact divide (offsets outside the fast domain)
    for (int j = threadIdx.x; j < nf; j += TH_THREADS) {
        const float* o = p.offsets + (size_t)(f0 + j) * 4;
        exact_s[j] = !(rdf_fastdiv_domain(o[0]) && rdf_fastdiv_domain(o[1]) && rdf_fastdiv_domain(o[2]) && rdf_fastdiv_domain(o[3]));
    }
    for (int i = threadIdx.x; i < nf * p.NT; i += TH_THREADS) thr_s[i] = __ldg(p.thresholds + (size_t)f0 * p.NT + i);
    const int per_slot = p.FC * p.NB * p.C;
    if (p.privatise)
        for (int i = threadIdx.x; i < p.S * per_slot; i += TH_THREADS) hist_s[i] = 0u;
    __syncthreads();

    const int64_t tile0 = (int64_t)blockIdx.x * TH_PIX;
    const int per_img = p.W * p.H;
    for (int k = threadIdx.x; k < TH_PIX; k += TH_THREADS) {
        const int64_t i = tile0 + k;
        if (i >= p.num_pixels) break;
        const int g = __ldg(p.nodes_by_pixel + i);
        if (g < 0) continue;                                                  // tree_train.cu:36-37
        const int slot = __ldg(p.node_slot + g);
        if (slot < 0) continue;                                               // node not in this block (tree_train.cu:42)
        const int n = (int)(i / per_img);
        const int rem = (int)(i - (int64_t)n * per_img);
        const int Y = rem / p.W, X = rem - Y * p.W;
        const uint16_t* img = p.depth + (size_t)n * per_img;
        const unsigned d = __ldg(img + rem);
        const unsigned label = __ldg(p.labels + i);
        if (label >= (unsigned)p.C) continue;
        const float df = (float)d;
        const float rcp = __frcp_rn(df);
        uint32_t* dst = p.privatise ? hist_s + (size_t)slot * per_slot
                                    : p.hist + ((size_t)slot * p.F + f0) * p.NB * p.C;
        for (int j = 0; j < nf; j++) {
            const float4 o = *reinterpret_cast<const float4*>(off_s + 4 * j);
            // compute_feature with scale 1 (tree_train.cu:58); d == 0 -> 0.f (decision_tree_common.hpp:12)
            float f = 0.f;
            if (d != 0u)
                f = exact_s[j] ? rdf_feature<true>(img, p.W, p.H, X, Y, df, rcp, o.x, o.y, o.z, o.w)
                               : rdf_feature<false>(img, p.W, p.H, X, Y, df, rcp, o.x, o.y, o.z, o.w);
            // bin = #{k : t_k <= f}: branch-free binary search over the ascending thresholds
            const float* th = thr_s + j * p.NT;
            int lo = 0, len = p.NT;
            while (len > 0) {
                const int half = len >> 1;
                const bool go = th[lo + half] <= f;
                lo = go ? lo + half + 1 : lo;
                len = go ? len - half - 1 : half;
            }
            atomicAdd(dst + ((size_t)j * p.NB + lo) * p.C + label, 1u);
        }
    }
    if (p.privatise) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.S * per_slot; i += TH_THREADS) {
            const uint32_t v = hist_s[i];
            if (v) {
                const int slot = i / per_slot;
                const int r = i - slot * per_slot;                 // (j, bin, class) within the chunk
                const int j = r / (p.NB * p.C);
                if (j < nf) atomicAdd(p.hist + ((size_t)slot * p.F + f0) * p.NB * p.C + r, v);
            }
        }
    }
}

extern "C" int rdf_train_hist(const uint16_t* depth_dev, const uint16_t* labels_dev, const int32_t* nodes_by_pixel_dev,
                              int num_images, int dim_x, int dim_y, const int32_t* node_slot_dev, int num_slots,
                              const float* offsets_dev, const float* thresholds_dev, int num_features, int num_thresholds,
                              int num_classes, uint32_t* hist_dev, void* stream) {
    RDF_REQUIRE(depth_dev && labels_dev && nodes_by_pixel_dev && node_slot_dev && offsets_dev && thresholds_dev && hist_dev,
                "rdf_train_hist: NULL argument");
    RDF_REQUIRE(num_images >= 0 && dim_x > 0 && dim_y > 0 && num_slots >= 1 && num_features >= 1 && num_thresholds >= 1 &&
                    num_classes >= 1 && num_classes <= RDF_MAX_CLASSES,
                "rdf_train_hist: bad shape");
    rdf_hist_params p;
    p.depth = depth_dev; p.labels = labels_dev; p.nodes_by_pixel = nodes_by_pixel_dev; p.node_slot = node_slot_dev;
    p.offsets = offsets_dev; p.thresholds = thresholds_dev; p.hist = hist_dev;
    p.num_pixels = (int64_t)num_images * dim_x * dim_y;
    p.W = dim_x; p.H = dim_y; p.S = num_slots; p.F = num_features; p.NT = num_thresholds; p.NB = num_thresholds + 1;
    p.C = num_classes;
    if (p.num_pixels == 0) return RDF_OK;

    const size_t smem_budget = 200 * 1024;
    const size_t per_feature_static = sizeof(float) * (4 + (size_t)p.NT);
    const size_t per_feature_hist = sizeof(uint32_t) * (size_t)p.S * p.NB * p.C;
    int fc = (int)(smem_budget / (per_feature_static + per_feature_hist));
    p.privatise = fc >= 8 ? 1 : 0;
    if (!p.privatise) fc = (int)(smem_budget / per_feature_static);
    if (fc > TH_MAX_FC) fc = TH_MAX_FC;
    if (fc > p.F) fc = p.F;
    if (fc < 1) {
        rdf_set_error("rdf_train_hist: %d thresholds per feature do not fit shared memory", p.NT);
        return RDF_ERR_UNSUPPORTED;
    }
    p.FC = fc;
    const size_t smem = per_feature_static * fc + (p.privatise ? per_feature_hist * fc : 0);
    RDF_ENSURE_DYN_SMEM(rdf_train_hist_kernel, smem_budget + 1024);
    const int64_t tiles = (p.num_pixels + TH_PIX - 1) / TH_PIX;
    const int chunks = (p.F + fc - 1) / fc;
    RDF_REQUIRE(tiles <= 0x7fffffffLL && chunks <= 65535, "rdf_train_hist: grid too large (%lld tiles, %d feature chunks)",
                (long long)tiles, chunks);
    rdf_train_hist_kernel<<<dim3((unsigned)tiles, (unsigned)chunks), TH_THREADS, smem, rdf_stream(stream)>>>(p);
    RDF_LAUNCH_CHECK("rdf_train_hist_kernel");
    return RDF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// split histograms, bucketed: pixels grouped by node first
// ---------------------------------------------------------------------------------------------------------------
// The raster kernel above privatises [slots x feature chunk] histograms, so the feature chunk shrinks as 1 / slots and, past
// ~100 slots, every update is a global atomic (cfg4 on a B200: 311 / 418 / 207 / 210 ms per level at 1 / 16 / 256 / 4096 nodes,
// profiles/r01_train_cfg4.md).  Here the active pixels are first grouped by histogram slot (counting sort of pixel indices:
// rdf_train_bucket, once per level and node block, reused by every proposal block), so a CTA works on pixels of ONE node at
// a time: its shared-memory histogram is [feature chunk][NT+1][C] whatever the number of nodes, flushed with one global
// reduction per non-zero counter (pair) when the node changes.  Lanes of a warp are consecutive pixels of the same node evaluating
// the same feature: probes of neighbouring pixels share sectors, thresholds are shared-memory broadcasts, and every lane adds 1
// to its counter with red.shared - the shared-memory atomic unit resolves collisions faster than a match.any aggregation did.
// Tunables (overridable with -D for variant builds, tools/build_variant.sh): threads per CTA, CTAs per SM the shared-memory
// budget is split over, features evaluated together by one thread.
// Measured on cfg4 (profiles/r02_ncu_train.md, gpurun_out/r02_variants*.log).  While the kernel was latency-bound (match.any
// aggregation, exact-divide test in the loop) several small CTAs per SM won; with the lean loop the per-pixel setup of every
// feature chunk is what is left to amortise, and ONE 1024-thread CTA per SM with the whole 220 KB of shared memory (160 features
// per chunk at NT = 64, C = 4) is ahead: 63.8 / 62.9 / 64.2 ms at levels 0 / 8 / 12 against 64.4 / 63.8 / 64.8 (2 x 512 threads,
// 96 KB) and 65.7 / 65.3 / 66.2 (3 x 384, 60 KB).
#ifndef TB_THREADS
#define TB_THREADS 1024
#endif
#ifndef TB_CTAS_PER_SM
#define TB_CTAS_PER_SM 1
#endif
#ifndef TB_U
#define TB_U 8                  // features evaluated together by one thread (independent load chains)
#endif
static_assert((TB_U & (TB_U - 1)) == 0 && TB_U >= 1, "TB_U must be a power of two: feature chunks are rounded with bit masks");
#define TB_TILE 8192            // sorted pixels per CTA
#define TB_MAX_FC 256
#ifndef TB_SMEM_KB
#define TB_SMEM_KB (TB_CTAS_PER_SM == 1 ? 220 : TB_CTAS_PER_SM == 2 ? 110 : 60)   /* 3 x 60 KB leave 48 KB of L1 for the probes */
#endif
#define TB_SMEM_BUDGET ((size_t)TB_SMEM_KB * 1024)

struct rdf_bucket_ws {           // layout of the caller-provided workspace
    int total;                   // number of bucketed pixels (device-side value, never read by the host)
    int pad[3];
    // int starts[S + 1]; int cursor[S]; int list[num_pixels];
};

static size_t tb_align16(size_t v) { return (v + 15) & ~(size_t)15; }

extern "C" int rdf_train_bucket_workspace_bytes(int64_t num_pixels, int num_slots, size_t* bytes) {
    RDF_REQUIRE(bytes && num_pixels >= 0 && num_slots >= 1, "rdf_train_bucket_workspace_bytes: bad argument");
    *bytes = sizeof(rdf_bucket_ws) + tb_align16(sizeof(int) * ((size_t)num_slots + 1)) + tb_align16(sizeof(int) * (size_t)num_slots) +
             sizeof(int) * (size_t)num_pixels;
    return RDF_OK;
}

__device__ __forceinline__ int tb_slot_of(const int32_t* __restrict__ nodes, const int32_t* __restrict__ node_slot, int64_t i) {
    const int g = __ldg(nodes + i);
    return g < 0 ? -1 : __ldg(node_slot + g);
}

// counts[slot] += 1 per active pixel (MODE 0) / list[cursor[slot]++] = pixel (MODE 1), one atomic per (warp, slot)
template <int MODE>
__global__ void __launch_bounds__(256) rdf_train_bucket_kernel(const int32_t* __restrict__ nodes, const int32_t* __restrict__ node_slot,
                                                               int64_t n, int* __restrict__ counts_or_cursor, int* __restrict__ list) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) / 32 * 32;                              // whole warps stay converged for match.any
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const int slot = i < n ? tb_slot_of(nodes, node_slot, i) : -1;
        const unsigned grp = __match_any_sync(0xffffffffu, slot);
        if (slot >= 0) {
            const int leader = __ffs(grp) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(counts_or_cursor + slot, __popc(grp));
            if (MODE == 1) {
                base = __shfl_sync(grp, base, leader);
                list[base + __popc(grp & ((1u << lane) - 1u))] = (int)i;
            }
        }
    }
}

// The same two passes with the slot counters privatised per CTA in shared memory (S <= TBK_MAX_SLOTS): at shallow levels every
// pixel of the GPU lands in a handful of slots, and one global atomic per (warp, slot) on the same few addresses made the level-0
// bucket 0.74 ms (7 % of an 8-GPU level).  A CTA takes chunks of TBK_CHUNK pixels: count them per slot with shared-memory
// increments, reserve each slot's range with ONE global atomic, then (MODE 1) hand out positions inside the range with a second
// round of shared-memory increments.  The order inside a slot's list depends on scheduling; the histograms do not (integer sums).
#define TBK_THREADS 512
#define TBK_CHUNK 8192
#define TBK_MAX_SLOTS 8192
template <int MODE>
__global__ void __launch_bounds__(TBK_THREADS) rdf_train_bucket_cta_kernel(const int32_t* __restrict__ nodes, const int32_t* __restrict__ node_slot,
                                                                           int64_t n, int S, int* __restrict__ counts_or_cursor,
                                                                           int* __restrict__ list) {
    extern __shared__ int tbk_s[];                                            // cnt[S], base[S]
    int* cnt = tbk_s;
    int* base = tbk_s + S;
    for (int s = threadIdx.x; s < S; s += TBK_THREADS) cnt[s] = 0;
    __syncthreads();
    for (int64_t c0 = (int64_t)blockIdx.x * TBK_CHUNK; c0 < n; c0 += (int64_t)gridDim.x * TBK_CHUNK) {
        int slot[TBK_CHUNK / TBK_THREADS];
#pragma unroll
        for (int k = 0; k < TBK_CHUNK / TBK_THREADS; k++) {
            const int64_t i = c0 + k * TBK_THREADS + threadIdx.x;
            slot[k] = i < n ? tb_slot_of(nodes, node_slot, i) : -1;
            if (slot[k] >= 0) atomicAdd(&cnt[slot[k]], 1);
        }
        __syncthreads();
        for (int s = threadIdx.x; s < S; s += TBK_THREADS) {
            const int c = cnt[s];
            if (c) {
                const int b = atomicAdd(counts_or_cursor + s, c);
                if (MODE == 1) base[s] = b;
                cnt[s] = 0;
            }
        }
        __syncthreads();
        if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < TBK_CHUNK / TBK_THREADS; k++)
                if (slot[k] >= 0) list[base[slot[k]] + atomicAdd(&cnt[slot[k]], 1)] = (int)(c0 + k * TBK_THREADS + threadIdx.x);
            __syncthreads();
            for (int s = threadIdx.x; s < S; s += TBK_THREADS) cnt[s] = 0;
            __syncthreads();
        }
    }
}

// single CTA: exclusive scan of counts[S] (held in starts[]) -> starts[0..S], cursor[s] = starts[s], total
__global__ void __launch_bounds__(1024) rdf_train_bucket_scan_kernel(int* __restrict__ starts, int* __restrict__ cursor, int S,
                                                                      int* __restrict__ total) {
    __shared__ int warp_tot[32];
    __shared__ int running;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < S; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < S ? starts[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const int excl = running + (warp ? warp_tot[warp - 1] : 0) + incl - v;
        if (i < S) {
            starts[i] = excl;
            cursor[i] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 0) running += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        starts[S] = running;
        *total = running;
    }
}

extern "C" int rdf_train_bucket(const int32_t* nodes_by_pixel_dev, int64_t num_pixels, const int32_t* node_slot_dev, int num_slots,
                                void* workspace_dev, size_t workspace_bytes, void* stream) {
    RDF_REQUIRE(nodes_by_pixel_dev && node_slot_dev && workspace_dev, "rdf_train_bucket: NULL argument");
    RDF_REQUIRE(num_pixels >= 0 && num_pixels < ((int64_t)1 << 31) && num_slots >= 1, "rdf_train_bucket: bad shape");
    size_t need = 0;
    rdf_train_bucket_workspace_bytes(num_pixels, num_slots, &need);
    RDF_REQUIRE(workspace_bytes >= need, "rdf_train_bucket: workspace %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t st = rdf_stream(stream);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace_dev);
    int* total = reinterpret_cast<int*>(ws);
    int* starts = reinterpret_cast<int*>(ws + sizeof(rdf_bucket_ws));
    int* cursor = reinterpret_cast<int*>(ws + sizeof(rdf_bucket_ws) + tb_align16(sizeof(int) * ((size_t)num_slots + 1)));
    int* list = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(cursor) + tb_align16(sizeof(int) * (size_t)num_slots));
    RDF_CUDA(cudaMemsetAsync(starts, 0, sizeof(int) * ((size_t)num_slots + 1), st));
    int blocks = (int)((num_pixels + 255) / 256);
    if (blocks > rdf_sm_count() * 16) blocks = rdf_sm_count() * 16;
    if (blocks < 1) blocks = 1;
    if (num_slots <= TBK_MAX_SLOTS) {
        int cblocks = (int)((num_pixels + TBK_CHUNK - 1) / TBK_CHUNK);
        if (cblocks > rdf_sm_count() * 4) cblocks = rdf_sm_count() * 4;
        if (cblocks < 1) cblocks = 1;
        const size_t smem = sizeof(int) * 2 * (size_t)num_slots;
        RDF_ENSURE_DYN_SMEM(rdf_train_bucket_cta_kernel<0>, sizeof(int) * 2 * TBK_MAX_SLOTS);
        RDF_ENSURE_DYN_SMEM(rdf_train_bucket_cta_kernel<1>, sizeof(int) * 2 * TBK_MAX_SLOTS);
        rdf_train_bucket_cta_kernel<0><<<cblocks, TBK_THREADS, smem, st>>>(nodes_by_pixel_dev, node_slot_dev, num_pixels, num_slots, starts, nullptr);
        rdf_train_bucket_scan_kernel<<<1, 1024, 0, st>>>(starts, cursor, num_slots, total);
        rdf_train_bucket_cta_kernel<1><<<cblocks, TBK_THREADS, smem, st>>>(nodes_by_pixel_dev, node_slot_dev, num_pixels, num_slots, cursor, list);
    } else {
        rdf_train_bucket_kernel<0><<<blocks, 256, 0, st>>>(nodes_by_pixel_dev, node_slot_dev, num_pixels, starts, nullptr);
        rdf_train_bucket_scan_kernel<<<1, 1024, 0, st>>>(starts, cursor, num_slots, total);
        rdf_train_bucket_kernel<1><<<blocks, 256, 0, st>>>(nodes_by_pixel_dev, node_slot_dev, num_pixels, cursor, list);
    }
    RDF_LAUNCH_CHECK("rdf_train_bucket kernels");
    return RDF_OK;
}

struct rdf_histb_params {
    const uint16_t* depth;
    const uint16_t* labels;
    const int* total;
    const int* starts;           // [S + 1]
    const int* list;             // [total] pixel indices grouped by slot
    const float* offsets;        // [F][4]
    const float* thresholds;     // [F][NT]
    uint32_t* hist;              // [S][F][NB][C]
    int W, H, S, F, NT, NB, C, FC;
    // feature-sharded reduction fused into the flush (multi-GPU): feature f belongs to rank f / Fo, whose buffer
    // owner_hist[rank] is laid out [S][Fo][NB][C]; the pointers are peer-mapped (NVLink) device addresses
    uint32_t* const* owner_hist; // [world] or NULL
    int Fo;
    int pair64;                  // flush two neighbouring counters as one 64-bit add (even C, 8-byte aligned histogram buffers)
};

// ceil(t) for "t <= f" against an integer-valued feature: t <= f  <=>  ceil(t) <= f.  NaN never counts (-> INT_MAX).
__device__ __forceinline__ int tb_int_thresh(float t) {
    if (!(t == t)) return 0x7fffffff;
    if (t >= 2147483648.f) return 0x7fffffff;
    if (t <= -2147483648.f) return (int)0x80000000;
    return __float2int_ru(t);
}

// shared-memory u32 add through the .shared window (a generic-address atomic costs an address-space conversion per use)
__device__ __forceinline__ void tb_red_shared(unsigned smem_addr, unsigned v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_addr), "r"(v) : "memory");
}
// One step of the threshold search on a running shared address: if (*(int*)(addr + OFF) <= f) addr += INC.
// (Not volatile: the threshold table is read-only after the CTA's first barrier, and the TB_U chains should interleave freely.)
template <int OFF, int INC>
__device__ __forceinline__ void tb_step(unsigned& addr, int f) {
    asm("{\n\t.reg .pred p;\n\t.reg .s32 t;\n\tld.shared.s32 t, [%0+%1];\n\tsetp.le.s32 p, t, %3;\n\t@p add.u32 %0, %0, %2;\n\t}"
        : "+r"(addr)
        : "n"(OFF), "n"(INC), "r"(f));
}
// steps LG, LG-1, ..., 0 (thresholds[pos + 2^LG - 1] <= f ? pos += 2^LG) for TB_U interleaved searches, in lock step
template <int LG, int U>
__device__ __forceinline__ void tb_search_steps(unsigned (&pa)[U], const int (&f)[U]) {
    if constexpr (LG >= 0) {
#pragma unroll
        for (int u = 0; u < U; u++) tb_step<(4 << LG) - 4, (4 << LG)>(pa[u], f[u]);
        tb_search_steps<LG - 1, U>(pa, f);
    }
}
__device__ __forceinline__ int tb_lds(unsigned smem_addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_addr));
    return v;
}

// All features of the CTA's chunk at one pixel: probes -> integer feature -> threshold bin -> histogram increment.
// compute_feature with scale 1 (tree_train.cu:58).  d == 0 -> the feature is 0.f (decision_tree_common.hpp:12): on the reciprocal
// path that costs nothing per evaluation - with rcp = 0 every quotient is exactly 0, both probes land on the pixel itself and their
// difference is 0; the exact-divide path (features flagged at staging time) keeps the explicit select.
template <int LOG2NTP, bool ANY_EXACT>
__device__ __forceinline__ void tb_pixel_features(const uint16_t* __restrict__ img, int W, int H, int X, int Y, unsigned d,
                                                  const float4* __restrict__ off_s, const unsigned char* __restrict__ exact_s,
                                                  unsigned thr_base, unsigned hist_base, unsigned row_bytes, unsigned C, unsigned label4,
                                                  int nfp) {
    constexpr int NTP = 1 << LOG2NTP;
    const float df = (float)d;
    const float rcp = d != 0u ? __frcp_rn(df) : 0.f;
    const float xm = (float)X + RDF_MAGIC_F, ym = (float)Y + RDF_MAGIC_F;
    for (int j0 = 0; j0 < nfp; j0 += TB_U) {
        int f[TB_U];
        unsigned tha[TB_U];                                              // shared address of the feature's thresholds
#pragma unroll
        for (int u = 0; u < TB_U; u++) {
            const int j = j0 + u;
            const float4 o = off_s[j];
            tha[u] = thr_base + (unsigned)j * (NTP * 4u);
            if (ANY_EXACT && exact_s[j]) {
                f[u] = rdf_feature_i<true>(img, W, H, X, Y, df, rcp, xm, ym, o.x, o.y, o.z, o.w);
                f[u] = d != 0u ? f[u] : 0;
            } else {
                f[u] = rdf_feature_i<false>(img, W, H, X, Y, df, rcp, xm, ym, o.x, o.y, o.z, o.w);
            }
        }
        // bin = #{k : t_k <= f}: binary search by halving steps, then one last compare.  The running value is the shared ADDRESS of
        // thresholds[pos] (one predicated add per step: load, compare, add), not an index that would need a select and a second
        // add to become an address.
        unsigned pa[TB_U];
#pragma unroll
        for (int u = 0; u < TB_U; u++) pa[u] = tha[u];
        tb_search_steps<LOG2NTP - 1, TB_U>(pa, f);                               // thresholds[pos + step - 1] <= f ?
#pragma unroll
        for (int u = 0; u < TB_U; u++) tb_step<0, 4>(pa[u], f[u]);               // pos <= NTP - 1; then pos = bin in 0..NT
#pragma unroll
        for (int u = 0; u < TB_U; u++) {
            const unsigned key = (pa[u] - tha[u]) * C + label4;                  // byte offset of [bin][label]
            // every lane adds 1 and the shared-memory atomic unit resolves the collisions (ptxas: ATOMS.POPC.INC).  Measured against
            // one add per distinct counter of the warp (match.any + popc + leader test, 12 instructions and a MATCH latency per
            // evaluation): 76-78 ms per cfg4 level instead of 95-99, and still faster (115 vs 124 ms) when every pixel of the GPU
            // hits the same counter (all probes outside the image) - profiles/r02_ncu_train.md
            tb_red_shared(hist_base + (unsigned)(j0 + u) * row_bytes + key, 1u);
        }
    }
}

// LOG2NTP: thresholds of a feature are padded in shared memory to NTP = 2^LOG2NTP entries with INT_MAX, so the search is a
// fixed, fully unrolled sequence of LOG2NTP + 1 loads without bound checks.
template <int LOG2NTP>
__global__ void __launch_bounds__(TB_THREADS, TB_CTAS_PER_SM) rdf_train_hist_bucketed_kernel(const rdf_histb_params p) {
    constexpr int NTP = 1 << LOG2NTP;
    extern __shared__ __align__(16) unsigned char tb_smem[];
    float4* off_s = reinterpret_cast<float4*>(tb_smem);                          // [FC]
    int* thr_s = reinterpret_cast<int*>(off_s + p.FC);                           // [FC][NTP]
    uint32_t* hist_s = reinterpret_cast<uint32_t*>(thr_s + (size_t)p.FC * NTP);  // [FC][NB][C]
    unsigned char* exact_s = reinterpret_cast<unsigned char*>(hist_s + (size_t)p.FC * p.NB * p.C);   // [FC]
    __shared__ int s_slot;

    const int total = __ldg(p.total);
    const int tile0 = blockIdx.x * TB_TILE;
    if (tile0 >= total) return;
    const int tile1 = min(total, tile0 + TB_TILE);
    const int f0 = blockIdx.y * p.FC;
    const int nf = min(p.FC, p.F - f0);

    // the chunk is padded to a multiple of TB_U features (zero offsets, INT_MAX thresholds): padded rows are evaluated and
    // counted into their own (never flushed) histogram rows, which keeps the inner loop free of tail tests
    const int nfp = (nf + TB_U - 1) / TB_U * TB_U;                                // <= FC (FC is a multiple of 4)
    __shared__ int s_any_exact;
    if (threadIdx.x == 0) s_any_exact = (p.W > 65535 || p.H > 65535) ? 1 : 0;
    __syncthreads();
    for (int j = threadIdx.x; j < nfp; j += TB_THREADS) {
        const float4 o = j < nf ? __ldg(reinterpret_cast<const float4*>(p.offsets) + f0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        off_s[j] = o;
        const bool ex = !(rdf_fastfloor_domain(o.x) && rdf_fastfloor_domain(o.y) && rdf_fastfloor_domain(o.z) && rdf_fastfloor_domain(o.w));
        exact_s[j] = ex;
        if (ex) s_any_exact = 1;
    }
    for (int i = threadIdx.x; i < nfp * NTP; i += TB_THREADS) {
        const int j = i >> LOG2NTP, k = i & (NTP - 1);
        thr_s[i] = (j < nf && k < p.NT) ? tb_int_thresh(__ldg(p.thresholds + (size_t)(f0 + j) * p.NT + k)) : 0x7fffffff;
    }
    const int per_chunk = nf * p.NB * p.C;
    for (int i = threadIdx.x; i < nfp * p.NB * p.C; i += TB_THREADS) hist_s[i] = 0u;
    if (threadIdx.x == 0) {                                                      // slot of the first pixel of the tile
        int lo = 0, hi = p.S - 1;                                               // last s with starts[s] <= tile0
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(p.starts + mid) <= tile0) lo = mid; else hi = mid - 1;
        }
        s_slot = lo;
    }
    __syncthreads();

    const int per_img = p.W * p.H;
    const bool any_exact = s_any_exact != 0;                                     // uniform: some feature needs __fdiv_rn
    const unsigned thr_base = (unsigned)__cvta_generic_to_shared(thr_s);
    const unsigned hist_base = (unsigned)__cvta_generic_to_shared(hist_s);
    const unsigned row_bytes = (unsigned)(p.NB * p.C) * 4u;                      // one feature's histogram
    int cur = tile0;
    int slot = s_slot;
    while (cur < tile1) {
        while (__ldg(p.starts + slot + 1) <= cur) slot++;                        // skip empty slots (uniform across the CTA)
        const int run1 = min(tile1, __ldg(p.starts + slot + 1));
        for (int e0 = cur; e0 < run1; e0 += TB_THREADS) {
            const int e = e0 + threadIdx.x;
            const bool have = e < run1;
            int X = 0, Y = 0;
            unsigned d = 0, label = 0xffffffffu;
            const uint16_t* img = p.depth;
            if (have) {
                const int i = __ldg(p.list + e);
                const int n = i / per_img;
                const int rem = i - n * per_img;
                Y = rem / p.W;
                X = rem - Y * p.W;
                img = p.depth + (size_t)n * per_img;
                d = __ldg(img + rem);
                label = __ldg(p.labels + i);
            }
            const bool use = have && label < (unsigned)p.C;                      // labels outside 0..C-1 cannot be counted
            if (!use) continue;
            // TB_U features per trip: their 2 * TB_U probes are issued before any is consumed and the TB_U threshold searches
            // advance in lock step, so one warp keeps several independent load chains in flight.  Two instances of the loop: the
            // common one knows that no feature of the chunk needs the exact divide (CTA-uniform) and has no per-evaluation test.
            if (any_exact)
                tb_pixel_features<LOG2NTP, true>(img, p.W, p.H, X, Y, d, off_s, exact_s, thr_base, hist_base, row_bytes, (unsigned)p.C, label * 4u, nfp);
            else
                tb_pixel_features<LOG2NTP, false>(img, p.W, p.H, X, Y, d, off_s, exact_s, thr_base, hist_base, row_bytes, (unsigned)p.C, label * 4u, nfp);
        }
        __syncthreads();
        // flush this node's counters and clear them for the next node.  With an even class count two neighbouring counters
        // ([bin][c], [bin][c+1]: 8-byte aligned) travel as ONE 64-bit add: no counter of a histogram can reach 2^32 (there are fewer
        // than 2^31 pixels), so the low half never carries into the high one - a third fewer reductions on cfg4's three used classes,
        // and over NVLink the packet count, not the payload, is what the flush costs.
        const bool pair = p.pair64 != 0;
        if (p.owner_hist == nullptr) {
            uint32_t* out = p.hist + ((size_t)slot * p.F + f0) * p.NB * p.C;
            if (pair) {
                for (int i = 2 * threadIdx.x; i < per_chunk; i += 2 * TB_THREADS) {
                    const uint2 v = *reinterpret_cast<const uint2*>(hist_s + i);
                    if (v.x | v.y) {
                        atomicAdd(reinterpret_cast<unsigned long long*>(out + i), (unsigned long long)v.x | ((unsigned long long)v.y << 32));
                        *reinterpret_cast<uint2*>(hist_s + i) = make_uint2(0u, 0u);
                    }
                }
            } else {
                for (int i = threadIdx.x; i < per_chunk; i += TB_THREADS) {
                    const uint32_t v = hist_s[i];
                    if (v) {
                        atomicAdd(out + i, v);
                        hist_s[i] = 0u;
                    }
                }
            }
        } else {
            // reduce-scatter fused into the flush: every counter goes straight to the rank that owns its feature, as a
            // system-scope reduction over NVLink (no separate collective, and it overlaps the other CTAs' evaluation)
            const int row = p.NB * p.C;
            const int step = pair ? 2 : 1;
            for (int i = step * threadIdx.x; i < per_chunk; i += step * TB_THREADS) {
                const uint32_t v = hist_s[i], v2 = pair ? hist_s[i + 1] : 0u;
                if (v | v2) {
                    const int j = i / row;
                    const int f = f0 + j;
                    const int o = f / p.Fo;
                    uint32_t* dst = p.owner_hist[o] + ((size_t)slot * p.Fo + (f - o * p.Fo)) * row + (i - j * row);
                    if (pair) {
                        atomicAdd_system(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)v | ((unsigned long long)v2 << 32));
                        hist_s[i + 1] = 0u;
                    } else {
                        atomicAdd_system(dst, v);
                    }
                    hist_s[i] = 0u;
                }
            }
        }
        __syncthreads();
        cur = run1;
    }
}

static int rdf_hist_bucketed_launch(const uint16_t* depth_dev, const uint16_t* labels_dev, int num_images, int dim_x, int dim_y,
                                    const void* bucket_workspace_dev, int num_slots, const float* offsets_dev,
                                    const float* thresholds_dev, int num_features, int num_thresholds, int num_classes,
                                    uint32_t* hist_dev, uint32_t* const* owner_hist_dev, int world, void* stream) {
    RDF_REQUIRE(depth_dev && labels_dev && bucket_workspace_dev && offsets_dev && thresholds_dev && (hist_dev || owner_hist_dev),
                "rdf_train_hist_bucketed: NULL argument");
    RDF_REQUIRE(num_images >= 0 && dim_x > 0 && dim_y > 0 && num_slots >= 1 && num_features >= 1 && num_thresholds >= 1 &&
                    num_classes >= 1 && num_classes <= RDF_MAX_CLASSES,
                "rdf_train_hist_bucketed: bad shape");
    const int64_t num_pixels = (int64_t)num_images * dim_x * dim_y;
    RDF_REQUIRE(num_pixels < ((int64_t)1 << 31), "rdf_train_hist_bucketed: too many pixels");
    if (num_pixels == 0) return RDF_OK;
    const unsigned char* ws = reinterpret_cast<const unsigned char*>(bucket_workspace_dev);
    rdf_histb_params p;
    p.depth = depth_dev; p.labels = labels_dev;
    p.total = reinterpret_cast<const int*>(ws);
    p.starts = reinterpret_cast<const int*>(ws + sizeof(rdf_bucket_ws));
    p.list = reinterpret_cast<const int*>(ws + sizeof(rdf_bucket_ws) + tb_align16(sizeof(int) * ((size_t)num_slots + 1)) +
                                          tb_align16(sizeof(int) * (size_t)num_slots));
    p.offsets = offsets_dev; p.thresholds = thresholds_dev; p.hist = hist_dev;
    p.W = dim_x; p.H = dim_y; p.S = num_slots; p.F = num_features; p.NT = num_thresholds; p.NB = num_thresholds + 1;
    p.C = num_classes;
    p.owner_hist = owner_hist_dev;
    p.Fo = owner_hist_dev ? (num_features + world - 1) / world : num_features;
    // peer-mapped owner buffers come from a 256-byte aligned allocator (torch symmetric memory); a caller's local buffer is checked
    p.pair64 = (num_classes % 2 == 0) && (owner_hist_dev != nullptr || (reinterpret_cast<uintptr_t>(hist_dev) & 7u) == 0);
    int log2ntp = 0;
    while ((1 << log2ntp) < p.NT) log2ntp++;
    RDF_REQUIRE(log2ntp <= 10, "rdf_train_hist_bucketed: at most 1024 thresholds per feature (got %d)", p.NT);
    const size_t smem_budget = TB_SMEM_BUDGET;
    const size_t per_feature = sizeof(float4) + sizeof(int) * ((size_t)1 << log2ntp) + sizeof(uint32_t) * (size_t)p.NB * p.C + 1;
    // features per CTA: a multiple of TB_U (the loop evaluates TB_U at a time, the chunk is padded to it) that fits shared memory
    int fc_max = (int)((smem_budget - 64) / per_feature);
    if (fc_max > TB_MAX_FC) fc_max = TB_MAX_FC;
    fc_max &= ~(TB_U - 1);
    if (fc_max < TB_U) {
        rdf_set_error("rdf_train_hist_bucketed: %d thresholds x %d classes per feature do not fit shared memory (use rdf_train_hist)",
                      p.NT, p.C);
        return RDF_ERR_UNSUPPORTED;
    }
    const int chunks = (p.F + fc_max - 1) / fc_max;
    int fc = (p.F + chunks - 1) / chunks;                                        // equal chunks
    fc = (fc + TB_U - 1) & ~(TB_U - 1);                                          // <= fc_max (a multiple of TB_U itself)
    p.FC = fc;
    const size_t smem = per_feature * fc + 64;
    const int64_t tiles = (num_pixels + TB_TILE - 1) / TB_TILE;
    const int chunks2 = (p.F + fc - 1) / fc;
    RDF_REQUIRE(chunks2 <= 65535, "rdf_train_hist_bucketed: too many feature chunks (%d)", chunks2);
    const dim3 grid((unsigned)tiles, (unsigned)chunks2);
#define TB_CASE(L)                                                                                                      \
    case L:                                                                                                             \
        RDF_ENSURE_DYN_SMEM(rdf_train_hist_bucketed_kernel<L>, smem_budget + 4096);                                    \
        rdf_train_hist_bucketed_kernel<L><<<grid, TB_THREADS, smem, rdf_stream(stream)>>>(p);                           \
        break;
    switch (log2ntp) {
        TB_CASE(0) TB_CASE(1) TB_CASE(2) TB_CASE(3) TB_CASE(4) TB_CASE(5) TB_CASE(6) TB_CASE(7) TB_CASE(8) TB_CASE(9) TB_CASE(10)
    }
#undef TB_CASE
    RDF_LAUNCH_CHECK("rdf_train_hist_bucketed_kernel");
    return RDF_OK;
}

extern "C" int rdf_train_hist_bucketed(const uint16_t* depth_dev, const uint16_t* labels_dev, int num_images, int dim_x, int dim_y,
                                       const void* bucket_workspace_dev, int num_slots, const float* offsets_dev,
                                       const float* thresholds_dev, int num_features, int num_thresholds, int num_classes,
                                       uint32_t* hist_dev, void* stream) {
    RDF_REQUIRE(hist_dev != nullptr, "rdf_train_hist_bucketed: hist_dev is NULL");
    return rdf_hist_bucketed_launch(depth_dev, labels_dev, num_images, dim_x, dim_y, bucket_workspace_dev, num_slots, offsets_dev,
                                    thresholds_dev, num_features, num_thresholds, num_classes, hist_dev, nullptr, 1, stream);
}

extern "C" int rdf_train_hist_bucketed_p2p(const uint16_t* depth_dev, const uint16_t* labels_dev, int num_images, int dim_x, int dim_y,
                                           const void* bucket_workspace_dev, int num_slots, const float* offsets_dev,
                                           const float* thresholds_dev, int num_features, int num_thresholds, int num_classes,
                                           uint32_t* const* owner_hist_dev, int world, void* stream) {
    RDF_REQUIRE(owner_hist_dev != nullptr && world >= 1, "rdf_train_hist_bucketed_p2p: bad owner table");
    return rdf_hist_bucketed_launch(depth_dev, labels_dev, num_images, dim_x, dim_y, bucket_workspace_dev, num_slots, offsets_dev,
                                    thresholds_dev, num_features, num_thresholds, num_classes, nullptr, owner_hist_dev, world, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// pick best split per active node
// ---------------------------------------------------------------------------------------------------------------
// fp32 operation order = what nvcc 12.9 emits for the reference's gini helpers on sm_100a (checked in PTX):
//   gini(c)  : s = cvt.rn.f32.u64(sum); p = fma(c_i/s, c_i/s, p) in class order; 1 - p          (tree_train.cu:72-80)
//   gain     : lt = (l/p) * gini(l); rem = fma(r/p, gini(r), lt); gini(parent) - rem            (tree_train.cu:82-89)
__device__ __forceinline__ float rdf_gini(const unsigned long long* c, int C, unsigned long long sum) {
    const float s = __ull2float_rn(sum);
    float p = 0.f;
    for (int i = 0; i < C; i++) {
        const float pi = __fdiv_rn(__ull2float_rn(c[i]), s);
        p = __fmaf_rn(pi, pi, p);
    }
    return __fsub_rn(1.f, p);
}

#define PB_WARPS 16
#define PB_THREADS (PB_WARPS * 32)

struct rdf_pick_params {
    const int32_t* active_nodes;
    const int32_t* node_slot;
    const unsigned long long* parent_counts;   // [2^D][C] by node id
    const uint32_t* hist;                      // [S][F][NB][C]
    const float* offsets;
    const float* thresholds;
    float* tree;
    unsigned long long* next_counts;           // [2^D][C] by child node id
    float* best_gain;                          // [num_active]
    int num_active, S, F, NT, NB, C, level, D;
    // candidates mode (feature-sharded split search): hist holds features f_offset .. f_offset + F - 1 of the global proposal
    // block; the node's best local candidate is written here instead of being finalised
    float* cand_gain;                          // [num_active] or NULL
    int* cand_idx;                             // [num_active] global candidate index (f_offset + f) * NT + k
    unsigned long long* cand_counts;           // [num_active][2][C] child counts of that candidate
    int f_offset;
    int f_stride;                              // features per slot in hist's layout (>= F)
};

// Writes the node record for the winning candidate (tree_train.cu:170-235).  best_i indexes p.offsets / p.thresholds.
__device__ __forceinline__ void pb_finalize(const rdf_pick_params& p, int a, int node, float best_g, int best_i,
                                            const unsigned long long* left, const unsigned long long* right,
                                            const unsigned long long* par) {
    const int C = p.C;
    if (!(best_g > -1.f)) return;                                  // the reference asserts best_g > -1 (tree_train.cu:170)
    if (best_g <= p.best_gain[a]) return;                          // tree_train.cu:172
    p.best_gain[a] = best_g;
    unsigned long long ls = 0, rs = 0, ps = 0;
    for (int c = 0; c < C; c++) {
        ls += left[c];
        rs += right[c];
        ps += par[c];
    }
    const float p_sum_f = __ull2float_rn(ps);
    const int bf = best_i / p.NT, bk = best_i - bf * p.NT;
    const int E = 7 + 2 * C;
    float* out = p.tree + ((((size_t)1 << p.level) - 1) + node) * E;
    out[0] = p.offsets[4 * bf + 0];                                // tree_train.cu:183-186
    out[1] = p.offsets[4 * bf + 1];
    out[2] = p.offsets[4 * bf + 2];
    out[3] = p.offsets[4 * bf + 3];
    out[4] = p.thresholds[(size_t)bf * p.NT + bk];
    if (best_g <= 0.f) {                                           // tree_train.cu:190-198
        out[5] = 0.f;
        out[6] = 0.f;
        for (int c = 0; c < C; c++) {
            const float v = __fdiv_rn(__ull2float_rn(par[c]), p_sum_f);
            out[7 + c] = v;
            out[7 + C + c] = v;
        }
        return;
    }
    for (int side = 0; side < 2; side++) {                         // tree_train.cu:201-235
        const unsigned long long* cnt = side ? right : left;
        const unsigned long long sum = side ? rs : ls;
        const float sum_f = __ull2float_rn(sum);
        float* pdf = out + 7 + side * C;
        int cut = -1;
        for (int c = 0; c < C; c++)
            if (__fdiv_rn(__ull2float_rn(cnt[c]), sum_f) >= 0.999f) { cut = c; break; }      // count_above_cutoff (:92-97)
        if (cut > -1) {
            out[5 + side] = 0.f;
            pdf[cut] = 1.f;                                        // other entries keep whatever they held (:204-206)
        } else if (p.level == p.D - 1) {
            out[5 + side] = 0.f;
            for (int c = 0; c < C; c++) pdf[c] = __fdiv_rn(__ull2float_rn(cnt[c]), sum_f);
        } else {
            out[5 + side] = -1.f;
            unsigned long long* nc = p.next_counts + ((size_t)2 * node + side) * C;
            for (int c = 0; c < C; c++) nc[c] = cnt[c];
        }
    }
}

__device__ __forceinline__ unsigned pb_warp_incl_scan(unsigned v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ---- candidate screening -------------------------------------------------------------------------------------------------
// The exact gain costs two div.rn.f32 per (class, threshold) and was 10 % of HBM peak (profiles/r01_train_cfg4.md).  Almost all
// candidates lose by far, so each feature is first scored in cheap arithmetic and only features that could hold the winner run
// the exact sequence.  With ls, rs the side totals, lc / rc the per-class side counts, T = ls + rs and P the parent total,
//     gain = gini(parent) - (ls/P) (1 - sum (lc/ls)^2) - (rs/P) (1 - sum (rc/rs)^2) = gini(parent) - (T - Q) / P,
//     Q = sum lc^2 / ls + sum rc^2 / rs,
// i.e. for one node the gain is an increasing function of Q alone (an empty side contributes 0 to Q and the reference then
// defines gain = 0, which is the formula's value too when T == P; rows with T != P skip the screen).  Q~ is Q in fp32 with
// MUFU reciprocals: |Q~ - Q| <= d2 P with d2 < 1e-6 (a dozen roundings of 6e-8 on terms <= P; counts above 2^24 convert with
// an error of <= 2 counts out of P >= 2^24).  The exact fp32 gain g differs from the real one by d1 < 1e-6 (same count of
// roundings on values <= 1).  If candidate x has the greatest exact gain and y the greatest Q~ seen anywhere, then
// g(x) >= g(y) => Q(x) >= Q(y) - 2 d1 P => Q~(x) >= Q~(y) - 2 (d1 + d2) P.  A feature is therefore skipped only when all its
// candidates have Q~ < (greatest Q~ seen so far) - PB_SCREEN_MARGIN * P with PB_SCREEN_MARGIN = 2e-5 >= 5 x 2 (d1 + d2): no
// candidate that attains the greatest exact gain is ever skipped, every surviving feature is scored exactly as before, and the
// winner (greatest gain, ties -> smallest index) is the same candidate.  Tie-heavy nodes simply screen nothing out.
#define PB_SCREEN_MARGIN 2e-5f

__device__ __forceinline__ uint4 pb_ld4(const uint32_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ float pb_q_side(float s0, float s1, float s2, float s3, float n) {
    // sum c^2 / n, 0 for an empty side
    const float sq = __fmaf_rn(s3, s3, __fmaf_rn(s2, s2, __fmaf_rn(s1, s1, __fmul_rn(s0, s0))));
    return n > 0.f ? __fdividef(sq, n) : 0.f;
}

// Greatest Q~ over the two candidates this lane owns (k = 2 lane, 2 lane + 1) of a feature with NT <= 64 thresholds, -1 if it
// owns none.  *row_total receives T.  CT == 4: four classes, one 128-bit load per bin; CT == 0: any C, one class at a time.
struct pb_row4 {                                                   // the bins a lane owns of a 4-class row: 2 lane, 2 lane + 1, and bin 64
    uint4 a, b, last;
};

__device__ __forceinline__ pb_row4 pb_load_row4(const uint32_t* __restrict__ h, int NB, int lane) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    pb_row4 r;
    r.a = 2 * lane < NB ? pb_ld4(h + 8 * lane) : z;
    r.b = 2 * lane + 1 < NB ? pb_ld4(h + 8 * lane + 4) : z;
    r.last = NB > 64 ? pb_ld4(h + 4 * 64) : z;                     // bin 64 exists only for NT == 64
    return r;
}

template <int CT>
__device__ __forceinline__ float pb_screen_feature(const uint32_t* __restrict__ h, const pb_row4& row, int NT, int C, int lane,
                                                   bool small, unsigned* row_total) {
    const int NB = NT + 1, k0 = 2 * lane, k1 = k0 + 1;
    float q0, q1;
    unsigned T;
    if (CT == 4) {
        const uint4 a = row.a, b = row.b, last = row.last;
        uint4 in = make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);    // inclusive scan over lanes of the pair sums
        if (small) {
            // every count of this node is below 2^16 (the parent holds fewer than 65536 pixels): two classes share a register, so
            // the 4-class scan is 10 shuffles + 10 adds instead of 20 + 20 (no field can carry into its neighbour)
            unsigned xy = in.x | (in.y << 16), zw = in.z | (in.w << 16);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t0 = __shfl_up_sync(0xffffffffu, xy, o), t1 = __shfl_up_sync(0xffffffffu, zw, o);
                if (lane >= o) { xy += t0; zw += t1; }
            }
            in = make_uint4(xy & 0xffffu, xy >> 16, zw & 0xffffu, zw >> 16);
        } else {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned tx = __shfl_up_sync(0xffffffffu, in.x, o), ty = __shfl_up_sync(0xffffffffu, in.y, o);
                const unsigned tz = __shfl_up_sync(0xffffffffu, in.z, o), tw = __shfl_up_sync(0xffffffffu, in.w, o);
                if (lane >= o) { in.x += tx; in.y += ty; in.z += tz; in.w += tw; }
            }
        }
        const unsigned t0 = __shfl_sync(0xffffffffu, in.x, 31) + last.x, t1 = __shfl_sync(0xffffffffu, in.y, 31) + last.y;
        const unsigned t2 = __shfl_sync(0xffffffffu, in.z, 31) + last.z, t3 = __shfl_sync(0xffffffffu, in.w, 31) + last.w;
        T = t0 + t1 + t2 + t3;
        const float f0 = __uint2float_rn(t0), f1 = __uint2float_rn(t1), f2 = __uint2float_rn(t2), f3 = __uint2float_rn(t3);
        const float Tf = (f0 + f1) + (f2 + f3);
        // candidate k1: left = bins 0..k1 = the inclusive scan; candidate k0: that minus bin k1
        const float l10 = __uint2float_rn(in.x), l11 = __uint2float_rn(in.y), l12 = __uint2float_rn(in.z), l13 = __uint2float_rn(in.w);
        const float l00 = __uint2float_rn(in.x - b.x), l01 = __uint2float_rn(in.y - b.y), l02 = __uint2float_rn(in.z - b.z),
                    l03 = __uint2float_rn(in.w - b.w);
        const float ls1 = (l10 + l11) + (l12 + l13), ls0 = (l00 + l01) + (l02 + l03);
        q1 = pb_q_side(l10, l11, l12, l13, ls1) + pb_q_side(f0 - l10, f1 - l11, f2 - l12, f3 - l13, Tf - ls1);
        q0 = pb_q_side(l00, l01, l02, l03, ls0) + pb_q_side(f0 - l00, f1 - l01, f2 - l02, f3 - l03, Tf - ls0);
    } else {
        float sl0 = 0.f, sl1 = 0.f, sr0 = 0.f, sr1 = 0.f, ls0 = 0.f, ls1 = 0.f, Tf = 0.f;
        T = 0u;
        for (int c = 0; c < C; c++) {
            const unsigned a = k0 < NB ? __ldg(h + k0 * C + c) : 0u;
            const unsigned b = k1 < NB ? __ldg(h + k1 * C + c) : 0u;
            const unsigned last = NB > 64 ? __ldg(h + 64 * C + c) : 0u;
            unsigned in = a + b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, in, o);
                if (lane >= o) in += t;
            }
            const unsigned tc = __shfl_sync(0xffffffffu, in, 31) + last;
            T += tc;
            const float tf = __uint2float_rn(tc), l1 = __uint2float_rn(in), l0 = __uint2float_rn(in - b);
            Tf += tf;
            ls0 += l0; ls1 += l1;
            sl0 = __fmaf_rn(l0, l0, sl0); sl1 = __fmaf_rn(l1, l1, sl1);
            sr0 = __fmaf_rn(tf - l0, tf - l0, sr0); sr1 = __fmaf_rn(tf - l1, tf - l1, sr1);
        }
        q0 = (ls0 > 0.f ? __fdividef(sl0, ls0) : 0.f) + (Tf - ls0 > 0.f ? __fdividef(sr0, Tf - ls0) : 0.f);
        q1 = (ls1 > 0.f ? __fdividef(sl1, ls1) : 0.f) + (Tf - ls1 > 0.f ? __fdividef(sr1, Tf - ls1) : 0.f);
    }
    *row_total = T;
    q0 = k0 < NT ? q0 : -1.f;
    q1 = k1 < NT ? q1 : -1.f;
    return fmaxf(q0, q1);
}

// One CTA per active node; each WARP takes whole features (f = warp, warp + PB_WARPS, ...) with lanes across the threshold
// bins, so a feature's [NT+1][C] histogram (1040 B at NT=64, C=4) is read with coalesced loads; left counts per threshold come
// from warp prefix sums over the bins.  (The first version gave each thread a feature and walked its 1040 bytes serially:
// 32 different cache lines per load instruction, 285 GB/s at 4096 nodes - profiles/r01_train_cfg4.md.)
// Per class the Gini terms are accumulated in class order exactly as the reference's compiled helpers do
// (cvt.rn.f32.u64, div.rn, fma; tree_train.cu:72-89).  Winner = greatest gain, ties -> smallest candidate index
// (= first in proposal order: feature-major, threshold-minor).
// SCREEN: -1 = score every feature exactly; 4 / 0 = screen first (pb_screen_feature<4> for C == 4, <0> for any C; NT <= 64).
template <int SCREEN>
__global__ void __launch_bounds__(PB_THREADS) rdf_train_pick_best_kernel(const rdf_pick_params p) {
    const int a = blockIdx.x;
    if (a >= p.num_active) return;
    const int node = p.active_nodes[a];
    const int slot = p.node_slot[node];
    if (slot < 0) return;                                          // not in this node block (tree_train.cu:135)
    const int C = p.C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    extern __shared__ unsigned long long pb_smem[];
    unsigned long long* par = pb_smem;                             // [C]
    unsigned long long* left = par + C;                            // [C]  (winner's child counts, thread 0)
    unsigned long long* right = left + C;                          // [C]
    unsigned* tot_w = reinterpret_cast<unsigned*>(right + C) + (size_t)warp * 2 * C;   // [C] class totals of the warp's feature
    unsigned* carry_w = tot_w + C;                                 // [C] running prefix per class
    __shared__ float red_g[PB_WARPS];
    __shared__ int red_i[PB_WARPS];
    __shared__ float s_gini_parent;
    __shared__ unsigned long long s_parent_sum;
    __shared__ unsigned s_qmax;                                    // greatest Q~ any warp of the CTA has seen (float bits, Q~ >= 0)

    for (int c = threadIdx.x; c < C; c += PB_THREADS) par[c] = p.parent_counts[(size_t)node * C + c];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int c = 0; c < C; c++) s += par[c];
        s_parent_sum = s;
        s_gini_parent = rdf_gini(par, C, s);
        s_qmax = 0u;
    }
    __syncthreads();
    const unsigned long long parent_sum = s_parent_sum;
    const float p_sum_f = __ull2float_rn(parent_sum);
    const float gini_parent = s_gini_parent;

    float best_g = -1.f;
    int best_i = 0x7fffffff;
    float q_run = 0.f;                                             // greatest Q~ this warp has seen (warp-uniform)
    const float q_margin = PB_SCREEN_MARGIN * fmaxf(p_sum_f, 1.f);
    // SCREEN == 4: the next feature's row is loaded before the current one is scored (one row per warp in flight is ~33 KB per
    // SM, about what HBM latency x bandwidth needs; two rows leave slack)
    const uint32_t* hist_node = p.hist + (size_t)slot * p.f_stride * p.NB * C;
    pb_row4 row_next;
    if (SCREEN == 4 && warp < p.F) row_next = pb_load_row4(hist_node + (size_t)warp * p.NB * C, p.NB, lane);
    for (int f = warp; f < p.F; f += PB_WARPS) {
        const uint32_t* h = hist_node + (size_t)f * p.NB * C;
        if (SCREEN >= 0) {
            pb_row4 row;
            if (SCREEN == 4) {
                row = row_next;
                if (f + PB_WARPS < p.F) row_next = pb_load_row4(h + (size_t)PB_WARPS * p.NB * C, p.NB, lane);
            }
            unsigned row_total;
            const float qm = pb_screen_feature<SCREEN>(h, row, p.NT, C, lane, parent_sum < 65536ull, &row_total);
            const float q_seen = fmaxf(q_run, __uint_as_float(*reinterpret_cast<volatile unsigned*>(&s_qmax)));
            const float q_feat = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(qm, 0.f))));
            if (q_feat > q_seen && lane == 0) atomicMax(&s_qmax, __float_as_uint(q_feat));
            q_run = fmaxf(q_seen, q_feat);
            // skip when no candidate of this feature can attain the greatest exact gain (see PB_SCREEN_MARGIN); a row whose
            // total differs from the parent's count (never with this library's own histograms) is always scored exactly
            if (q_feat < q_run - q_margin && (unsigned long long)row_total == parent_sum) continue;
        }
        // class totals over all NT + 1 bins
        unsigned T = 0;
        for (int c = 0; c < C; c++) {
            unsigned sum = 0;
            for (int b = lane; b < p.NB; b += 32) sum += __ldg(h + b * C + c);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) {
                tot_w[c] = sum;
                carry_w[c] = 0u;
            }
            T += sum;
        }
        __syncwarp();
        unsigned l_carry = 0;
        for (int b0 = 0; b0 < p.NT; b0 += 32) {                    // candidates k = b0 + lane: left = bins 0..k  (f < t_k)
            const int k = b0 + lane;
            const bool cand = k < p.NT;
            unsigned n = 0;
            if (cand)
                for (int c = 0; c < C; c++) n += __ldg(h + k * C + c);
            const unsigned n_incl = pb_warp_incl_scan(n, lane);
            const unsigned ls = l_carry + n_incl, rs = T - ls;
            l_carry += __shfl_sync(0xffffffffu, n_incl, 31);
            const float lsf = __uint2float_rn(ls), rsf = __uint2float_rn(rs);
            float pl = 0.f, pr = 0.f;
            for (int c = 0; c < C; c++) {
                const unsigned v = cand ? __ldg(h + k * C + c) : 0u;
                const unsigned v_incl = pb_warp_incl_scan(v, lane);
                const unsigned base = carry_w[c];
                const unsigned lc = base + v_incl, rc = tot_w[c] - lc;
                const unsigned chunk = __shfl_sync(0xffffffffu, v_incl, 31);
                __syncwarp();
                if (lane == 0) carry_w[c] = base + chunk;
                const float pil = __fdiv_rn(__uint2float_rn(lc), lsf);
                pl = __fmaf_rn(pil, pil, pl);
                const float pir = __fdiv_rn(__uint2float_rn(rc), rsf);
                pr = __fmaf_rn(pir, pir, pr);
            }
            __syncwarp();
            float g = 0.f;                                         // a side empty -> 0 (tree_train.cu:158-160)
            if (ls && rs) {
                const float lt = __fmul_rn(__fdiv_rn(lsf, p_sum_f), __fsub_rn(1.f, pl));
                const float rem = __fmaf_rn(__fdiv_rn(rsf, p_sum_f), __fsub_rn(1.f, pr), lt);
                g = __fsub_rn(gini_parent, rem);
            }
            if (cand && g > best_g) {                              // strict >: a lane sees its candidates in ascending index
                best_g = g;
                best_i = f * p.NT + k;
            }
        }
    }
    // greatest gain, ties -> smallest candidate index: across lanes, then across warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float g2 = __shfl_xor_sync(0xffffffffu, best_g, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (g2 > best_g || (g2 == best_g && i2 < best_i)) {
            best_g = g2;
            best_i = i2;
        }
    }
    if (lane == 0) {
        red_g[warp] = best_g;
        red_i[warp] = best_i;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int wv = 1; wv < PB_WARPS; wv++) {
        if (red_g[wv] > red_g[0] || (red_g[wv] == red_g[0] && red_i[wv] < red_i[0])) {
            red_g[0] = red_g[wv];
            red_i[0] = red_i[wv];
        }
    }
    best_g = red_g[0];
    best_i = red_i[0];
    // child counts of the winner, from its histogram row
    if (best_i != 0x7fffffff) {
        const int bf = best_i / p.NT, bk = best_i - bf * p.NT;
        const uint32_t* h = p.hist + ((size_t)slot * p.f_stride + bf) * p.NB * C;
        for (int c = 0; c < C; c++) {
            unsigned long long l = 0, tot = 0;
            for (int b = 0; b < p.NB; b++) {
                const unsigned long long v = h[b * C + c];
                tot += v;
                if (b <= bk) l += v;
            }
            left[c] = l;
            right[c] = tot - l;
        }
    }
    if (p.cand_gain) {                                             // feature-sharded search: hand the local winner over
        p.cand_gain[a] = best_g;
        p.cand_idx[a] = best_i == 0x7fffffff ? best_i : best_i + p.f_offset * p.NT;
        unsigned long long* cc = p.cand_counts + (size_t)a * 2 * C;
        for (int c = 0; c < C; c++) {
            cc[c] = best_i == 0x7fffffff ? 0ull : left[c];
            cc[C + c] = best_i == 0x7fffffff ? 0ull : right[c];
        }
        return;
    }
    pb_finalize(p, a, node, best_g, best_i, left, right, par);
}

// Feature-sharded split search, second step: every rank holds the gathered local winners of all ranks
// (gain / idx [world][num_active], counts [world][num_active][2][C]); greatest gain wins, ties -> smallest candidate index,
// i.e. exactly the candidate a single GPU scanning all features in order would have kept.  One thread per active node.
struct rdf_pick_final_params {
    rdf_pick_params base;
    const float* all_gain;
    const int* all_idx;
    const unsigned long long* all_counts;
    int world;
};

__global__ void __launch_bounds__(128) rdf_train_pick_finalize_kernel(const rdf_pick_final_params q) {
    const rdf_pick_params& p = q.base;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= p.num_active) return;
    const int node = p.active_nodes[a];
    if (p.node_slot[node] < 0) return;
    float best_g = -1.f;
    int best_i = 0x7fffffff, best_r = 0;
    for (int r = 0; r < q.world; r++) {
        const float g = q.all_gain[(size_t)r * p.num_active + a];
        const int i = q.all_idx[(size_t)r * p.num_active + a];
        if (g > best_g || (g == best_g && i < best_i)) {
            best_g = g;
            best_i = i;
            best_r = r;
        }
    }
    const unsigned long long* cc = q.all_counts + ((size_t)best_r * p.num_active + a) * 2 * p.C;
    pb_finalize(p, a, node, best_g, best_i, cc, cc + p.C, p.parent_counts + (size_t)node * p.C);
}

// RDF_PICK_NO_SCREEN=1 scores every candidate exactly (experiments / A-B tests of the screen)
static void pb_launch(const rdf_pick_params& p, size_t smem, cudaStream_t st) {
    const bool screen = p.NT <= 64 && RDF_GETENV_ONCE("RDF_PICK_NO_SCREEN") == nullptr;
    if (!screen) rdf_train_pick_best_kernel<-1><<<p.num_active, PB_THREADS, smem, st>>>(p);
    else if (p.C == 4) rdf_train_pick_best_kernel<4><<<p.num_active, PB_THREADS, smem, st>>>(p);
    else rdf_train_pick_best_kernel<0><<<p.num_active, PB_THREADS, smem, st>>>(p);
}

extern "C" int rdf_train_pick_best(int num_active, const int32_t* active_nodes_dev, const int32_t* node_slot_dev,
                                   const uint64_t* parent_counts_dev, const uint32_t* hist_dev, int num_slots,
                                   const float* offsets_dev, const float* thresholds_dev, int num_features,
                                   int num_thresholds, int num_classes, int level, int max_depth, float* tree_dev,
                                   uint64_t* next_counts_dev, float* best_gain_dev, void* stream) {
    RDF_REQUIRE(active_nodes_dev && node_slot_dev && parent_counts_dev && hist_dev && offsets_dev && thresholds_dev && tree_dev &&
                    next_counts_dev && best_gain_dev,
                "rdf_train_pick_best: NULL argument");
    RDF_REQUIRE(num_active >= 0 && num_slots >= 1 && num_features >= 1 && num_thresholds >= 1 && num_classes >= 1 &&
                    num_classes <= RDF_MAX_CLASSES && level >= 0 && level < max_depth && max_depth <= RDF_MAX_DEPTH,
                "rdf_train_pick_best: bad argument");
    if (num_active == 0) return RDF_OK;
    rdf_pick_params p;
    p.active_nodes = active_nodes_dev; p.node_slot = node_slot_dev;
    p.parent_counts = reinterpret_cast<const unsigned long long*>(parent_counts_dev);
    p.hist = hist_dev; p.offsets = offsets_dev; p.thresholds = thresholds_dev; p.tree = tree_dev;
    p.next_counts = reinterpret_cast<unsigned long long*>(next_counts_dev);
    p.best_gain = best_gain_dev;
    p.num_active = num_active; p.S = num_slots; p.F = num_features; p.NT = num_thresholds; p.NB = num_thresholds + 1;
    p.C = num_classes; p.level = level; p.D = max_depth;
    p.cand_gain = nullptr; p.cand_idx = nullptr; p.cand_counts = nullptr; p.f_offset = 0; p.f_stride = num_features;
    const size_t smem = sizeof(unsigned long long) * 3 * (size_t)num_classes + sizeof(unsigned) * (size_t)PB_WARPS * 2 * num_classes;
    RDF_REQUIRE(smem <= 48 * 1024, "rdf_train_pick_best: %d classes exceed the shared-memory scratch", num_classes);
    pb_launch(p, smem, rdf_stream(stream));
    RDF_LAUNCH_CHECK("rdf_train_pick_best_kernel");
    return RDF_OK;
}

extern "C" int rdf_train_pick_candidates(int num_active, const int32_t* active_nodes_dev, const int32_t* node_slot_dev,
                                         const uint64_t* parent_counts_dev, const uint32_t* hist_local_dev, int num_slots,
                                         int num_local_features, int feature_stride, int feature_offset, int num_thresholds,
                                         int num_classes, float* cand_gain_dev, int32_t* cand_idx_dev, uint64_t* cand_counts_dev,
                                         void* stream) {
    RDF_REQUIRE(active_nodes_dev && node_slot_dev && parent_counts_dev && hist_local_dev && cand_gain_dev && cand_idx_dev && cand_counts_dev,
                "rdf_train_pick_candidates: NULL argument");
    RDF_REQUIRE(num_active >= 0 && num_slots >= 1 && num_local_features >= 0 && feature_stride >= num_local_features &&
                    feature_offset >= 0 && num_thresholds >= 1 && num_classes >= 1 && num_classes <= RDF_MAX_CLASSES,
                "rdf_train_pick_candidates: bad argument");
    if (num_active == 0) return RDF_OK;
    rdf_pick_params p;
    memset(&p, 0, sizeof(p));
    p.active_nodes = active_nodes_dev; p.node_slot = node_slot_dev;
    p.parent_counts = reinterpret_cast<const unsigned long long*>(parent_counts_dev);
    p.hist = hist_local_dev;
    p.num_active = num_active; p.S = num_slots; p.F = num_local_features; p.NT = num_thresholds; p.NB = num_thresholds + 1;
    p.C = num_classes;
    p.cand_gain = cand_gain_dev; p.cand_idx = cand_idx_dev;
    p.cand_counts = reinterpret_cast<unsigned long long*>(cand_counts_dev);
    p.f_offset = feature_offset;
    p.f_stride = feature_stride;
    const size_t smem = sizeof(unsigned long long) * 3 * (size_t)num_classes + sizeof(unsigned) * (size_t)PB_WARPS * 2 * num_classes;
    pb_launch(p, smem, rdf_stream(stream));
    RDF_LAUNCH_CHECK("rdf_train_pick_best_kernel (candidates)");
    return RDF_OK;
}

extern "C" int rdf_train_pick_finalize(int num_active, const int32_t* active_nodes_dev, const int32_t* node_slot_dev,
                                       const uint64_t* parent_counts_dev, int world, const float* all_gain_dev,
                                       const int32_t* all_idx_dev, const uint64_t* all_counts_dev, const float* offsets_dev,
                                       const float* thresholds_dev, int num_thresholds, int num_classes, int level, int max_depth,
                                       float* tree_dev, uint64_t* next_counts_dev, float* best_gain_dev, void* stream) {
    RDF_REQUIRE(active_nodes_dev && node_slot_dev && parent_counts_dev && all_gain_dev && all_idx_dev && all_counts_dev && offsets_dev &&
                    thresholds_dev && tree_dev && next_counts_dev && best_gain_dev,
                "rdf_train_pick_finalize: NULL argument");
    RDF_REQUIRE(num_active >= 0 && world >= 1 && num_thresholds >= 1 && num_classes >= 1 && num_classes <= RDF_MAX_CLASSES &&
                    level >= 0 && level < max_depth && max_depth <= RDF_MAX_DEPTH,
                "rdf_train_pick_finalize: bad argument");
    if (num_active == 0) return RDF_OK;
    rdf_pick_final_params q;
    memset(&q, 0, sizeof(q));
    q.base.active_nodes = active_nodes_dev; q.base.node_slot = node_slot_dev;
    q.base.parent_counts = reinterpret_cast<const unsigned long long*>(parent_counts_dev);
    q.base.offsets = offsets_dev; q.base.thresholds = thresholds_dev; q.base.tree = tree_dev;
    q.base.next_counts = reinterpret_cast<unsigned long long*>(next_counts_dev);
    q.base.best_gain = best_gain_dev;
    q.base.num_active = num_active; q.base.NT = num_thresholds; q.base.NB = num_thresholds + 1; q.base.C = num_classes;
    q.base.level = level; q.base.D = max_depth;
    q.all_gain = all_gain_dev; q.all_idx = all_idx_dev;
    q.all_counts = reinterpret_cast<const unsigned long long*>(all_counts_dev);
    q.world = world;
    rdf_train_pick_finalize_kernel<<<(num_active + 127) / 128, 128, 0, rdf_stream(stream)>>>(q);
    RDF_LAUNCH_CHECK("rdf_train_pick_finalize_kernel");
    return RDF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// next active nodes (deterministic compaction: ascending active index, left child before right)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) rdf_train_next_active_kernel(const float* __restrict__ tree, int level, int C,
                                                                      const int32_t* __restrict__ active, int num_active,
                                                                      int32_t* __restrict__ next_active,
                                                                      int32_t* __restrict__ num_next) {
    __shared__ int warp_tot[32];
    __shared__ int running;
    const int E = 7 + 2 * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < num_active; base += 1024) {
        const int i = base + threadIdx.x;
        int node = 0, nl = 0, nr = 0;
        if (i < num_active) {
            node = active[i];
            const float* nd = tree + ((((size_t)1 << level) - 1) + node) * E;
            nl = nd[5] == -1.f;                                    // tree_train.cu:263
            nr = nd[6] == -1.f;                                    // tree_train.cu:268
        }
        const int mine = nl + nr;
        int incl = mine;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += v;
            }
            warp_tot[lane] = t;                                    // inclusive over warps
        }
        __syncthreads();
        const int start = running + (warp ? warp_tot[warp - 1] : 0) + incl - mine;
        if (nl) next_active[start] = 2 * node;
        if (nr) next_active[start + nl] = 2 * node + 1;
        __syncthreads();
        if (threadIdx.x == 0) running += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *num_next = running;
}

extern "C" int rdf_train_next_active(const float* tree_dev, int level, int max_depth, int num_classes,
                                     const int32_t* active_nodes_dev, int num_active, int32_t* next_active_dev,
                                     int32_t* num_next_active_dev, void* stream) {
    RDF_REQUIRE(tree_dev && active_nodes_dev && next_active_dev && num_next_active_dev, "rdf_train_next_active: NULL argument");
    RDF_REQUIRE(level >= 0 && level < max_depth && num_active >= 0 && num_classes >= 1, "rdf_train_next_active: bad argument");
    rdf_train_next_active_kernel<<<1, 1024, 0, rdf_stream(stream)>>>(tree_dev, level, num_classes, active_nodes_dev, num_active,
                                                                     next_active_dev, num_next_active_dev);
    RDF_LAUNCH_CHECK("rdf_train_next_active_kernel");
    return RDF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// advance pixels to their child node
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rdf_train_advance_kernel(const uint16_t* __restrict__ depth, int32_t* __restrict__ nodes,
                                                                 int64_t num_pixels, int W, int H,
                                                                 const float* __restrict__ tree, int level, int C) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= num_pixels) return;
    const int g = nodes[i];
    if (g == -1) return;                                           // tree_train.cu:294
    const int per_img = W * H;
    const int n = (int)(i / per_img);
    const int rem = (int)(i - (int64_t)n * per_img);
    const int Y = rem / W, X = rem - Y * W;
    const uint16_t* img = depth + (size_t)n * per_img;
    const int E = 7 + 2 * C;
    const float* nd = tree + ((((size_t)1 << level) - 1) + g) * E;
    const unsigned d = __ldg(img + rem);
    const float f = d == 0u ? 0.f : rdf_feature<true>(img, W, H, X, Y, (float)d, 0.f, __ldg(nd + 0), __ldg(nd + 1), __ldg(nd + 2), __ldg(nd + 3));
    const bool left = f < __ldg(nd + 4);
    const int status = __float2int_rd(__ldg(nd + (left ? 5 : 6)));
    nodes[i] = status != -1 ? -1 : 2 * g + (left ? 0 : 1);         // tree_train.cu:316-323
}

extern "C" int rdf_train_advance_pixels(const uint16_t* depth_dev, int32_t* nodes_by_pixel_dev, int num_images, int dim_x,
                                        int dim_y, const float* tree_dev, int level, int max_depth, int num_classes,
                                        void* stream) {
    RDF_REQUIRE(depth_dev && nodes_by_pixel_dev && tree_dev, "rdf_train_advance_pixels: NULL argument");
    RDF_REQUIRE(num_images >= 0 && dim_x > 0 && dim_y > 0 && level >= 0 && level < max_depth && num_classes >= 1,
                "rdf_train_advance_pixels: bad argument");
    const int64_t n = (int64_t)num_images * dim_x * dim_y;
    if (n == 0) return RDF_OK;
    const int64_t blocks = (n + 255) / 256;
    RDF_REQUIRE(blocks <= 0x7fffffffLL, "rdf_train_advance_pixels: too many pixels");
    rdf_train_advance_kernel<<<(unsigned)blocks, 256, 0, rdf_stream(stream)>>>(depth_dev, nodes_by_pixel_dev, n, dim_x, dim_y,
                                                                                tree_dev, level, num_classes);
    RDF_LAUNCH_CHECK("rdf_train_advance_kernel");
    return RDF_OK;
}
