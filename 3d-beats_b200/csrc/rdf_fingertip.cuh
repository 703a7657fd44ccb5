// Fingertip read-out arithmetic (src/3d_bz.py:503-522), shared by rdf_fingertip_z (csrc/rdf_frame.cu) and the mean-shift kernel's
// fused tail (csrc/rdf_meanshift.cu).
#pragma once
#include <stdint.h>

#define RF_MAX_FINGERTIPS 32

struct rf_fingertip_spec {
    const uint16_t* raw;       // raw camera frame uint16[H,W] (device or pinned host)
    const float* plane;        // float32[4][4] row-major, device
    double* z_out;             // float64[num_images, n]
    double* means_copy;        // nullable, float64[num_images, num_labels, 2]
    int n, r, W, H;            // fingertips, labels_reduce, frame size
    float ppx, ppy, fx, fy;
    int idx[RF_MAX_FINGERTIPS];   // 1-based label ids
};

__device__ __forceinline__ long long rf_astype_int32(double v) {
    // numpy float64 -> int32 on x86-64 (cvttsd2si): truncation toward zero, NaN / out of range -> INT_MIN
    if (!(v > -2147483649.0 && v < 2147483648.0)) return -2147483648ll;
    return (long long)(int)v;
}

// plane-space -z of the raw depth sample under centroid (mx, my), NaN = the reference's reset_positions()
__device__ __forceinline__ double rf_fingertip_eval(const rf_fingertip_spec& p, double mx, double my) {
    // `px *= LABELS_REDUCE` on an np.int32 scalar promotes to int64 under the NumPy the reference needs (< 1.24, Linux), so a NaN
    // centroid (INT_MIN) stays negative and resets the fingertip instead of wrapping to pixel 0
    const long long px = rf_astype_int32(mx) * (long long)p.r;
    const long long py = rf_astype_int32(my) * (long long)p.r;
    if (px < 0 || py < 0 || px >= p.W || py >= p.H) return __longlong_as_double(0x7ff8000000000000ll);
    const float z = (float)p.raw[(size_t)py * p.W + px];
    const float x = __fdiv_rn(__fsub_rn((float)px, p.ppx), p.fx);        // rs2_deproject_pixel_to_point, no distortion
    const float y = __fdiv_rn(__fsub_rn((float)py, p.ppy), p.fy);
    const double ptx = (double)__fmul_rn(z, x), pty = (double)__fmul_rn(z, y), ptz = (double)z;
    const double m0 = (double)p.plane[8], m1 = (double)p.plane[9], m2 = (double)p.plane[10], m3 = (double)p.plane[11];
    return -__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m0, ptx), __dmul_rn(m1, pty)), __dmul_rn(m2, ptz)), m3);
}
