// Minimal stand-in for the subset of GLM that the reference's src/cuda/points_ops.cu and calibrated_plane.cu use.
// TEST INFRASTRUCTURE (oracle) - never included by the product.
//
// Upstream tells users to clone https://github.com/g-truc/glm (unpinned; src/cuda/deps/readme.md:1) and does not vendor it, so
// GLM is absent from /root/reference.  This header is written from GLM's published behaviour (0.9.9 series, scalar "pure" code
// path, which is what nvcc sees), not copied from it.  What matters for parity is the ORDER of the fp32 operations:
//   * mat4 is column-major, m[c] is column c;
//   * mat4 * vec4  =  (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w), component-wise (detail/type_mat4x4.inl);
//   * dot(vec3)    =  x*x' + y*y' + z*z' summed left to right;  normalize(v) = v * (1 / sqrt(dot(v, v)));
//   * cross(a, b)  =  (a.y*b.z - b.y*a.z,  a.z*b.x - b.z*a.x,  a.x*b.y - b.x*a.y).
// nvcc's default FMA contraction applies to these expressions exactly as it does to GLM's.
#pragma once
#include <assert.h>
#include <math.h>
#include <stdio.h>

#ifdef __CUDACC__
#define GLM_MIN_FN __host__ __device__ inline
#else
#define GLM_MIN_FN inline
#endif

namespace glm {

struct vec3 {
    float x, y, z;
    vec3() = default;
    GLM_MIN_FN vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

struct vec4 {
    float x, y, z, w;
    vec4() = default;
    template <typename A, typename B, typename C, typename D>
    GLM_MIN_FN vec4(A x_, B y_, C z_, D w_) : x(static_cast<float>(x_)), y(static_cast<float>(y_)), z(static_cast<float>(z_)), w(static_cast<float>(w_)) {}
    GLM_MIN_FN vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    GLM_MIN_FN vec3 xyz() const { return vec3(x, y, z); }   // GLM_SWIZZLE
};

GLM_MIN_FN vec3 operator-(const vec3& v) { return vec3(-v.x, -v.y, -v.z); }
GLM_MIN_FN vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
GLM_MIN_FN vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
GLM_MIN_FN vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
GLM_MIN_FN vec4 operator*(const vec4& v, float s) { return vec4(v.x * s, v.y * s, v.z * s, v.w * s); }

GLM_MIN_FN float dot(const vec3& a, const vec3& b) {
    const vec3 t(a.x * b.x, a.y * b.y, a.z * b.z);
    return t.x + t.y + t.z;
}
GLM_MIN_FN vec3 normalize(const vec3& v) { return v * (1.0f / sqrtf(dot(v, v))); }
GLM_MIN_FN vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

struct mat4 {
    vec4 c[4];
    mat4() = default;
    GLM_MIN_FN explicit mat4(float s) {
        c[0] = vec4(s, 0.f, 0.f, 0.f); c[1] = vec4(0.f, s, 0.f, 0.f); c[2] = vec4(0.f, 0.f, s, 0.f); c[3] = vec4(0.f, 0.f, 0.f, s);
    }
    GLM_MIN_FN vec4& operator[](int i) { return c[i]; }
    GLM_MIN_FN const vec4& operator[](int i) const { return c[i]; }
};

GLM_MIN_FN mat4 transpose(const mat4& m) {
    mat4 r;
    r[0] = vec4(m[0].x, m[1].x, m[2].x, m[3].x);
    r[1] = vec4(m[0].y, m[1].y, m[2].y, m[3].y);
    r[2] = vec4(m[0].z, m[1].z, m[2].z, m[3].z);
    r[3] = vec4(m[0].w, m[1].w, m[2].w, m[3].w);
    return r;
}

GLM_MIN_FN vec4 operator*(const mat4& m, const vec4& v) {
    const vec4 mul0 = m[0] * v.x;
    const vec4 mul1 = m[1] * v.y;
    const vec4 add0 = mul0 + mul1;
    const vec4 mul2 = m[2] * v.z;
    const vec4 mul3 = m[3] * v.w;
    const vec4 add1 = mul2 + mul3;
    return add0 + add1;
}

}  // namespace glm
