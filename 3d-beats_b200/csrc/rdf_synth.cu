// Device-side synthetic inputs: bit-exact twins of rdf_b200/synth.py (depth_frames, hash_forest).
// Pure 32-bit integer hashing, so the CPU oracle and the GPU see identical frames and forests without any PCIe copy.
#include "rdf_common.cuh"

__device__ __forceinline__ uint32_t mix32(uint32_t h) {     // murmur3 finaliser
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256) rdf_synth_depth_kernel(uint16_t* __restrict__ out, int kind, int N, int W, int H,
                                                              uint32_t seed, int first_frame) {
    const int64_t total = (int64_t)N * H * W;
    // ellipse of the cfg-2 hand blob, scaled from 848x480 (synth.ellipse_geometry)
    const long long cx = W / 2, cy = H / 2;
    long long rx = (150LL * W) / 848, ry = (110LL * H) / 480;
    if (rx < 1) rx = 1;
    if (ry < 1) ry = 1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int64_t t = i / W;
        const int y = (int)(t % H);
        const uint32_t n = (uint32_t)(first_frame + (int)(t / H));
        const uint32_t h = mix32(seed ^ (n * 0x9E3779B1u) ^ ((uint32_t)y * 0x85EBCA77u) ^ ((uint32_t)x * 0xC2B2AE3Du));
        uint32_t d;
        if (kind == 1) {
            d = 1u + (h % 65534u);
        } else {
            d = 3000u + ((3u * (uint32_t)x + 2u * (uint32_t)y + 37u * n) % 1024u) + (h & 31u);
            if (kind == 2) {
                const long long dx = x - cx, dy = y - cy;
                if (dx * dx * ry * ry + dy * dy * rx * rx > rx * rx * ry * ry) d = RDF_NO_PIXEL;
            }
        }
        out[i] = (uint16_t)d;
    }
}

extern "C" int rdf_synth_depth(uint16_t* depth_dev, int kind, int num_images, int dim_x, int dim_y, uint32_t seed,
                               int first_frame, void* stream) {
    RDF_REQUIRE(depth_dev != nullptr && num_images >= 0 && dim_x > 0 && dim_y > 0 && kind >= 0 && kind <= 2,
                "rdf_synth_depth: bad argument");
    if (num_images == 0) return RDF_OK;
    rdf_synth_depth_kernel<<<rdf_sm_count() * 16, 256, 0, rdf_stream(stream)>>>(depth_dev, kind, num_images, dim_x, dim_y, seed, first_frame);
    RDF_LAUNCH_CHECK("rdf_synth_depth_kernel");
    return RDF_OK;
}

// canonical layout float32[T][2^D-1][7+2C]; full-depth forest (flags -1 above the last level, 0 on it), leaf pdfs k/1024.
__global__ void __launch_bounds__(256) rdf_synth_forest_kernel(float* __restrict__ canon, int T, int D, int C, uint32_t seed) {
    const int E = 7 + 2 * C;
    const int64_t NN = ((int64_t)1 << D) - 1;
    const int64_t first_last = ((int64_t)1 << (D - 1)) - 1;   // first row of level D-1
    const int64_t total = (int64_t)T * NN * E;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % E);
        const int64_t nr = i / E;
        const uint32_t row = (uint32_t)(nr % NN);
        const uint32_t t = (uint32_t)(nr / NN);
        const bool last = (int64_t)row >= first_last;
        const uint32_t h = mix32(seed ^ (t * 0x9E3779B1u) ^ (row * 0x85EBCA77u) ^ ((uint32_t)e * 0xC2B2AE3Du));
        float v;
        if (e < 5) {
            const uint32_t span = e < 4 ? 20u : 16u;
            const uint32_t expo = 127u + ((h >> 24) % span);
            v = __uint_as_float(((h & 1u) << 31) | (expo << 23) | ((h >> 1) & 0x7FFFFFu));
        } else if (e < 7) {
            v = last ? 0.f : -1.f;
        } else {
            v = last ? (float)(h & 1023u) / 1024.f : 0.f;
        }
        canon[i] = v;
    }
}

extern "C" int rdf_synth_forest(float* canon_dev, int num_trees, int max_depth, int num_classes, uint32_t seed, void* stream) {
    RDF_REQUIRE(canon_dev != nullptr && num_trees >= 1 && max_depth >= 1 && max_depth <= RDF_MAX_DEPTH && num_classes >= 1 &&
                    num_classes <= RDF_MAX_CLASSES,
                "rdf_synth_forest: bad argument");
    rdf_synth_forest_kernel<<<rdf_sm_count() * 16, 256, 0, rdf_stream(stream)>>>(canon_dev, num_trees, max_depth, num_classes, seed);
    RDF_LAUNCH_CHECK("rdf_synth_forest_kernel");
    return RDF_OK;
}
