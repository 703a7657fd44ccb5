/* CPU restatement of the reference's pre-/post-forest frame kernels (src/cuda/points_ops.cu, src/cuda/calibrated_plane.cu).
 * TEST INFRASTRUCTURE (oracle): only tests/, __graft_entry__.smoke() and bench.py's cpu/reference legs may load this.
 *
 * One function per reference kernel, so that the host sequence of src/3d_bz.py:159-220,390-456 can be replayed call by call
 * (oracle/frame_oracle.py).  Compiled with -ffp-contract=off: every fused multiply-add below is an explicit fmaf() placed where
 * nvcc 12.9 contracts the reference's expressions (checked in the SASS of oracle/_ref/libref_points.so), every other operation
 * is a separately rounded fp32 operation.
 *
 * Third-party arithmetic on this path: GLM (https://github.com/g-truc/glm, unpinned and un-vendored upstream,
 * src/cuda/deps/readme.md:1).  Its published mat4 * vec4 is (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w) per component, with
 * column-major m; transform_points multiplies by transpose(t) where t is the row-major numpy matrix reinterpreted as
 * column-major, so the product is the ordinary numpy M @ p.  Parity is pinned on the reference's kernels compiled against
 * oracle/ref_kernels/glm_min (tests/golden/frame.npz). */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define FO_EXPORT __attribute__((visibility("default")))

/* points_ops.cu:5-36 */
FO_EXPORT void fo_deproject_points(int W, int H, float ppx, float ppy, float f, const uint16_t* depth, float* pts) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const uint16_t d = depth[y * W + x];
            if (d > 0) {                                   /* d == 0: the point keeps whatever it held (points_ops.cu:24) */
                const float d_ = (float)d * 1.f;
                float* p = pts + 4 * (y * W + x);
                p[0] = (d_ * ((float)x - ppx)) / f;
                p[1] = (d_ * ((float)y - ppy)) / f;
                p[2] = d_;
                p[3] = 1.f;
            }
        }
}

/* points_ops.cu:66-75; t = row-major numpy float32[4][4] */
FO_EXPORT void fo_transform_points(int n, float* pts, const float* t) {
    for (int i = 0; i < n; i++) {
        float* p = pts + 4 * i;
        if (p[3] != 1.f) continue;
        float r[4];
        for (int c = 0; c < 4; c++) {
            const float* m = t + 4 * c;                    /* row c of M */
            const float add0 = fmaf(p[1], m[1], p[0] * m[0]);
            const float add1 = fmaf(p[2], m[2], m[3]);     /* m[3] * w with w == 1 */
            r[c] = add0 + add1;
        }
        memcpy(p, r, sizeof(r));
    }
}

/* calibrated_plane.cu:30-45 */
FO_EXPORT void fo_filter_points_by_plane(int n, float thresh, float* pts) {
    for (int i = 0; i < n; i++) {
        float* p = pts + 4 * i;
        if (p[3] != 1.f) continue;
        if (p[2] > -thresh) p[0] = p[1] = p[2] = p[3] = 0.f;
    }
}

/* points_ops.cu:131-146 */
FO_EXPORT void fo_remove_missing(int n, const float* pts, uint16_t* depth) {
    for (int i = 0; i < n; i++)
        if (pts[4 * i + 3] == 0.f) depth[i] = 0;
}

/* points_ops.cu:327-373 */
FO_EXPORT void fo_gaussian_depth_filter(int W, int H, int k, const float* wk, const uint16_t* in, uint16_t* out) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float w0 = 0.f, wn = 0.f, s = 0.f;
            for (int dy = 0; dy < k; dy++)
                for (int dx = 0; dx < k; dx++) {
                    const int cx = x + dx - k / 2, cy = y + dy - k / 2;
                    if (cy < 0 || cx < 0 || cy >= H || cx >= W) continue;
                    const uint16_t d = in[cy * W + cx];
                    const float w = wk[dy * k + dx];
                    if (d == 0) {
                        w0 += w;
                    } else {
                        wn += w;
                        s = fmaf((float)d, w, s);
                    }
                }
            uint16_t v = 0;
            if (!(w0 > wn)) {
                const float q = floorf(s / wn);            /* __float2uint_rd: NaN / negative -> 0, saturating */
                unsigned u = 0;
                if (q >= 4294967296.f) u = 0xffffffffu; else if (q > 0.f) u = (unsigned)q;
                v = (uint16_t)u;
            }
            out[y * W + x] = v;
        }
}

/* points_ops.cu:375-404 */
FO_EXPORT void fo_shrink_image(int W, int H, int level, const uint16_t* in, uint16_t* out) {
    const int f = 1 << level, wo = W / f, ho = H / f;
    for (int y = 0; y < ho; y++)
        for (int x = 0; x < wo; x++) out[y * wo + x] = (x * f >= W || y * f >= H) ? 0 : in[(y * f) * W + x * f];
}

/* points_ops.cu:407-438 */
FO_EXPORT void fo_grow_groups(int w, int h, const uint16_t* in, uint16_t* out) {
    static const int DX[4] = {-1, 1, 0, 0}, DY[4] = {0, 0, -1, 1};
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint16_t g = in[y * w + x];
            for (int i = 0; i < 4 && g == 0; i++) {
                const int xx = x + DX[i], yy = y + DY[i];
                if (xx >= 0 && xx < w && yy >= 0 && yy < h) g = in[yy * w + xx];
            }
            out[y * w + x] = g;
        }
}

/* points_ops.cu:441-463; d_out is only written where the group matches */
FO_EXPORT void fo_stencil_depth_image_by_group(int W, int H, int level, int group, const uint16_t* g_in, const uint16_t* d_in,
                                               uint16_t* d_out) {
    const int f = 1 << level, gw = W / f, gh = H / f;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int gx = x / f, gy = y / f;
            const int g = (gx < gw && gy < gh) ? g_in[gy * gw + gx] : 0;
            if (g != group) continue;
            d_out[y * W + x] = d_in[y * W + x];
        }
}

/* points_ops.cu:466-483 */
FO_EXPORT void fo_flip_x(int W, int H, const uint16_t* in, uint16_t* out) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) out[y * W + (W - (x + 1))] = in[y * W + x];
}

/* points_ops.cu:118-129 */
FO_EXPORT void fo_convert_0s_to_maxuint(int n, uint16_t* depth) {
    for (int i = 0; i < n; i++)
        if (depth[i] == 0) depth[i] = 65535;
}

/* points_ops.cu:258-281 */
FO_EXPORT void fo_make_rgba_from_labels(int W, int H, int num_colors, const uint16_t* labels, const uint8_t* colors, uint8_t* rgba) {
    (void)num_colors;
    for (int i = 0; i < W * H; i++) {
        const uint16_t l = labels[i];
        if (l == 0 || l == 65535) continue;
        memcpy(rgba + 4 * i, colors + 4 * (l - 1), 4);
    }
}

/* points_ops.cu:283-325 */
FO_EXPORT void fo_make_depth_rgba(int W, int H, uint16_t d_min, uint16_t d_max, const uint16_t* depth, uint8_t* rgba) {
    for (int i = 0; i < W * H; i++) {
        const uint16_t d = depth[i];
        uint8_t c[4] = {0, 0, 0, 255};
        if (d == 0) { c[0] = 195; c[1] = 157; c[2] = 152; }
        else if (d == 65535) { c[0] = 157; c[1] = 195; c[2] = 152; }
        else if (d < d_min || d > d_max) { c[0] = 157; c[1] = 152; c[2] = 195; }
        else {
            const float n_f = ((1.0f * d - d_min) * 255.f) / (float)(d_max - d_min);
            const float q = floorf(256.f - n_f);
            unsigned u = q > 0.f ? (unsigned)q : 0u;
            c[0] = c[1] = c[2] = (uint8_t)u;
        }
        memcpy(rgba + 4 * i, c, 4);
    }
}
