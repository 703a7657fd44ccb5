// Forest evaluation with the depth probes served by the texture units (experimental alternative to the global-load gather of
// rdf_eval.cu; SURVEY 8a: evaluate_image_using_forest, src/cuda/tree_eval.cu:24-137).
//
// The frames are copied once into a layered 2-D CUDA array holding 65535 - d.  A probe is then ONE texel fetch with integer
// coordinates (TLD): the texture unit does the address arithmetic and the bound test - outside the image it returns the
// border value 0, i.e. d = 65535, exactly Array3d::get's default (src/cuda/cu_utils.hpp:58-62,79-86) - and the block-linear
// array layout keeps 2-D neighbourhoods in the same sectors.  The feature becomes int(pv') - int(pu') (the complements cancel).
// Per probe that replaces two compares, an index multiply-add, a 64-bit address instruction, a default move and a predicated
// load of the global path by a single instruction.
#include "rdf_traverse.cuh"

struct rdf_depth_tex {
    cudaArray_t array;
    cudaTextureObject_t tex;
    cudaSurfaceObject_t surf;
    int max_images, W, H;
    int device;
};

__global__ void __launch_bounds__(256) rdf_tex_upload_kernel(const uint16_t* __restrict__ depth, cudaSurfaceObject_t surf, int W, int H,
                                                             int num_images) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    if (x >= W || y >= H || n >= num_images) return;
    const unsigned short v = (unsigned short)(65535u - __ldg(depth + ((size_t)n * H + y) * W + x));
    surf2DLayeredwrite(v, surf, x * 2, y, n);
}

extern "C" int rdf_depth_tex_create(int max_images, int dim_x, int dim_y, rdf_depth_tex** out) {
    RDF_REQUIRE(out != nullptr, "rdf_depth_tex_create: out is NULL");
    *out = nullptr;
    RDF_REQUIRE(max_images >= 1 && max_images <= 2048 && dim_x >= 1 && dim_x <= 32768 && dim_y >= 1 && dim_y <= 32768,
                "rdf_depth_tex_create: %d images of %dx%d outside the layered-array limits (2048 layers, 32768 x 32768)", max_images,
                dim_x, dim_y);
    rdf_depth_tex* t = new rdf_depth_tex();
    memset(t, 0, sizeof(*t));
    t->max_images = max_images; t->W = dim_x; t->H = dim_y;
    t->device = rdf_current_device();
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
    cudaError_t e = cudaMalloc3DArray(&t->array, &fmt, make_cudaExtent(dim_x, dim_y, max_images), cudaArrayLayered | cudaArraySurfaceLoadStore);
    if (e == cudaSuccess) {
        cudaResourceDesc res;
        memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypeArray;
        res.res.array.array = t->array;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;      // outside -> 0 = complement of 65535
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        e = cudaCreateTextureObject(&t->tex, &res, &td, nullptr);
        if (e == cudaSuccess) e = cudaCreateSurfaceObject(&t->surf, &res);
    }
    if (e != cudaSuccess) {
        rdf_set_error("rdf_depth_tex_create: %s", cudaGetErrorString(e));
        if (t->tex) cudaDestroyTextureObject(t->tex);
        if (t->array) cudaFreeArray(t->array);
        delete t;
        return RDF_ERR_CUDA;
    }
    *out = t;
    return RDF_OK;
}

extern "C" int rdf_depth_tex_destroy(rdf_depth_tex* t) {
    if (!t) return RDF_OK;
    if (t->surf) cudaDestroySurfaceObject(t->surf);
    if (t->tex) cudaDestroyTextureObject(t->tex);
    if (t->array) cudaFreeArray(t->array);
    delete t;
    return RDF_OK;
}

extern "C" int rdf_depth_tex_upload(rdf_depth_tex* t, const uint16_t* depth_dev, int num_images, void* stream) {
    RDF_REQUIRE(t && depth_dev, "rdf_depth_tex_upload: NULL argument");
    RDF_REQUIRE(num_images >= 0 && num_images <= t->max_images, "rdf_depth_tex_upload: %d images, capacity %d", num_images, t->max_images);
    if (num_images == 0) return RDF_OK;
    dim3 grid((t->W + 31) / 32, (t->H + 7) / 8, num_images);
    rdf_tex_upload_kernel<<<grid, 256, 0, rdf_stream(stream)>>>(depth_dev, t->surf, t->W, t->H, num_images);
    RDF_LAUNCH_CHECK("rdf_tex_upload_kernel");
    return RDF_OK;
}

__device__ __forceinline__ unsigned rdf_tld(cudaTextureObject_t tex, int layer, int x, int y) {
    unsigned r0, r1, r2, r3;
    asm("tex.a2d.v4.u32.s32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%8}];"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
        : "l"(tex), "r"(layer), "r"(x), "r"(y), "r"(0));
    (void)r1; (void)r2; (void)r3;
    return r0;
}

struct rdf_eval_tex_params {
    rdf_forest_view fv;
    const uint16_t* depth;       // original frames (centre depth)
    cudaTextureObject_t tex;     // 65535 - depth, layered
    const uint16_t* filter;
    uint16_t* labels;
    int W, H, w, h, r;
    int tiles_x;
    int filter_class;
    int smem_levels;
};

// Fast path only: scale == 1, every node in the reciprocal-divide / magic-floor domain (checked by the host), T <= 8.
template <int T, int WARP_W>
__global__ void __launch_bounds__(256, RDF_EVAL_MIN_BLOCKS_T(T)) rdf_eval_tex_kernel(const rdf_eval_tex_params p) {
    constexpr int WARP_H = 32 / WARP_W;
    constexpr int WARPS_X = 32 / WARP_W;
    __shared__ __align__(32) rdf_node_hdr hdr_s[T * 63];
    const int KS = min(p.fv.D, p.smem_levels);
    if (KS > 0) {
        rdf_stage_upper_levels(p.fv, KS, hdr_s);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x - tile_y * p.tiles_x;
    const int x = tile_x * 32 + (warp % WARPS_X) * WARP_W + (lane % WARP_W);
    const int y = tile_y * 8 + (warp / WARPS_X) * WARP_H + (lane / WARP_W);
    if (x >= p.w || y >= p.h) return;
    const int n = blockIdx.y;
    const size_t li = ((size_t)n * p.h + y) * p.w + x;
    if (p.filter_class != -1 && (int)__ldg(p.filter + li) != p.filter_class) return;
    const int X = x * p.r, Y = y * p.r;
    const unsigned d = __ldg(p.depth + ((size_t)n * p.H + Y) * p.W + X);
    if (d == 0u || d == RDF_NO_PIXEL) return;
    const float df = (float)d;
    const float rcp = __frcp_rn(df);
    const float xm = (float)X + RDF_MAGIC_F, ym = (float)Y + RDF_MAGIC_F;
    int state[T];
    const int M = (1 << KS) - 1;
#pragma unroll
    for (int t = 0; t < T; t++) state[t] = KS > 0 ? t * M : t * p.fv.nodes_per_tree;
    for (int j = 0; j < p.fv.D; j++) {
        const bool in_smem = j < KS;
        int all = state[0];
#pragma unroll
        for (int t = 1; t < T; t++) all &= state[t];
        if (all < 0) break;
        rdf_hdr_regs h[T];
#pragma unroll
        for (int t = 0; t < T; t++) {
            if (in_smem) {
                const int node = max(state[t], t * M);
                const float4* sp = reinterpret_cast<const float4*>(hdr_s + node);
                h[t].a = sp[0];
                h[t].b = *reinterpret_cast<const int4*>(sp + 1);
            } else {
                h[t] = rdf_load_hdr(p.fv.hdr, max(state[t], t * p.fv.nodes_per_tree));
            }
        }
        int f[T];
#pragma unroll
        for (int t = 0; t < T; t++) {
            int ux, uy, vx, vy;
            rdf_coord_fast2(h[t].a.x, h[t].a.y, df, rcp, xm, ym, ux, uy);
            rdf_coord_fast2(h[t].a.z, h[t].a.w, df, rcp, xm, ym, vx, vy);
            const unsigned pu = rdf_tld(p.tex, n, ux, uy);            // 65535 - d(u), 0 outside
            const unsigned pv = rdf_tld(p.tex, n, vx, vy);
            f[t] = (int)pv - (int)pu;                                 // = d(u) - d(v)
        }
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int next = (f[t] < h[t].b.x) ? h[t].b.y : h[t].b.z;
            state[t] = state[t] < 0 ? state[t] : next;
        }
    }
#pragma unroll
    for (int t = 0; t < T; t++)
        if (state[t] >= 0) state[t] = RDF_NO_LEAF;
    p.labels[li] = (uint16_t)rdf_vote<T>(p.fv, state, nullptr);
}

template <int T>
static void rdf_launch_tex(const rdf_eval_tex_params& p, dim3 grid, cudaStream_t st) {
    rdf_eval_tex_kernel<T, 16><<<grid, 256, 0, st>>>(p);
}

// Same contract as rdf_eval_forest for scale == 1 and no probability output; depth_dev are the frames that were uploaded into `tex`.
// Returns RDF_ERR_UNSUPPORTED when the forest holds nodes outside the fast-divide domain (use rdf_eval_forest).
extern "C" int rdf_eval_forest_tex(const rdf_forest_t* forest, const rdf_depth_tex* tex, const uint16_t* depth_dev, int num_images,
                                   const uint16_t* filter_dev, int filter_class, uint16_t* labels_dev, int labels_reduce, void* stream) {
    RDF_REQUIRE(forest && tex && depth_dev && labels_dev, "rdf_eval_forest_tex: NULL argument");
    RDF_REQUIRE(num_images >= 0 && num_images <= tex->max_images && labels_reduce >= 1, "rdf_eval_forest_tex: bad shape");
    if (forest->T > RDF_FAST_MAX_TREES || forest->has_exact_nodes) {
        rdf_set_error("rdf_eval_forest_tex: needs <= %d trees and offsets inside the fast-divide domain", RDF_FAST_MAX_TREES);
        return RDF_ERR_UNSUPPORTED;
    }
    rdf_eval_tex_params p;
    p.fv = rdf_view(forest);
    p.depth = depth_dev; p.tex = tex->tex; p.filter = filter_dev; p.labels = labels_dev;
    p.W = tex->W; p.H = tex->H; p.r = labels_reduce;
    p.w = tex->W / labels_reduce; p.h = tex->H / labels_reduce;
    if (num_images == 0 || p.w == 0 || p.h == 0) return RDF_OK;
    p.tiles_x = (p.w + 31) / 32;
    p.filter_class = filter_dev ? filter_class : -1;
    p.smem_levels = 6;
    dim3 grid((unsigned)(p.tiles_x * ((p.h + 7) / 8)), (unsigned)num_images);
    cudaStream_t st = rdf_stream(stream);
    switch (forest->T) {
        case 1: rdf_launch_tex<1>(p, grid, st); break;
        case 2: rdf_launch_tex<2>(p, grid, st); break;
        case 3: rdf_launch_tex<3>(p, grid, st); break;
        case 4: rdf_launch_tex<4>(p, grid, st); break;
        case 5: rdf_launch_tex<5>(p, grid, st); break;
        case 6: rdf_launch_tex<6>(p, grid, st); break;
        case 7: rdf_launch_tex<7>(p, grid, st); break;
        default: rdf_launch_tex<8>(p, grid, st); break;
    }
    RDF_LAUNCH_CHECK("rdf_eval_tex_kernel");
    return RDF_OK;
}
