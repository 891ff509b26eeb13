// Minimal stand-in for the GLM header tree, used ONLY to compile the reference
// rasterizer (oracle/_ref).  TEST INFRASTRUCTURE, not product code.
//
// The reference includes <glm/glm.hpp> (DGR/cuda_rasterizer/forward.h:19,
// backward.h:19, rasterizer_impl.cu:22) but does not vendor GLM
// (DGR/setup.py:32-38 points at a third_party/glm directory that is absent),
// and GLM is not installed in this image.  The reference sources use only
// glm::vec3 / vec4 / mat3, transpose, dot, length, max and a handful of
// arithmetic operators, so this header restates exactly those, following the
// published GLM 0.9.9 semantics:
//   * mat3 is column-major, m[c][r]; the 9-scalar constructor fills column by
//     column; mat3(s) is s * identity.
//   * operator*(mat3 a, mat3 b): r[j][i] = a[0][i]*b[j][0] + a[1][i]*b[j][1]
//     + a[2][i]*b[j][2]   (left-to-right sum; nvcc contracts it into FMAs).
//   * operator*(float, mat3) and operator*(mat3, float) scale every entry.
//   * dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z (left-to-right), length = sqrt(dot).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define GLMS_FN __host__ __device__ inline
#else
#define GLMS_FN inline
#endif

namespace glm {

struct vec3 {
    float x, y, z;
    GLMS_FN vec3() : x(0.f), y(0.f), z(0.f) {}
    GLMS_FN vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    GLMS_FN explicit vec3(float s) : x(s), y(s), z(s) {}
    GLMS_FN float& operator[](int i) { return (&x)[i]; }
    GLMS_FN const float& operator[](int i) const { return (&x)[i]; }
    GLMS_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLMS_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
    GLMS_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
    float x, y, z, w;
    GLMS_FN vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    GLMS_FN vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    GLMS_FN float& operator[](int i) { return (&x)[i]; }
    GLMS_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLMS_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLMS_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLMS_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLMS_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLMS_FN vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLMS_FN vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLMS_FN vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
GLMS_FN vec4 operator/(const vec4& a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }

GLMS_FN float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GLMS_FN float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
GLMS_FN float length(const vec3& a) { return sqrtf(dot(a, a)); }
GLMS_FN float length(const vec4& a) { return sqrtf(dot(a, a)); }
GLMS_FN vec3 max(const vec3& a, float s) { return vec3(fmaxf(a.x, s), fmaxf(a.y, s), fmaxf(a.z, s)); }

struct mat3 {
    vec3 c[3];
    GLMS_FN mat3() {}
    GLMS_FN explicit mat3(float s) { c[0] = vec3(s, 0.f, 0.f); c[1] = vec3(0.f, s, 0.f); c[2] = vec3(0.f, 0.f, s); }
    // Column-major fill; templated so that the reference's double literals
    // (img_W/2.0, 0.0, 1.0 in forward.cu:93-96) convert to float per entry.
    template <typename A, typename B, typename C, typename D, typename E, typename F, typename G, typename H, typename I>
    GLMS_FN mat3(A x0, B y0, C z0, D x1, E y1, F z1, G x2, H y2, I z2) {
        c[0] = vec3((float)x0, (float)y0, (float)z0);
        c[1] = vec3((float)x1, (float)y1, (float)z1);
        c[2] = vec3((float)x2, (float)y2, (float)z2);
    }
    GLMS_FN vec3& operator[](int i) { return c[i]; }
    GLMS_FN const vec3& operator[](int i) const { return c[i]; }
};

GLMS_FN mat3 operator*(const mat3& a, const mat3& b) {
    mat3 r;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++)
            r[j][i] = a[0][i] * b[j][0] + a[1][i] * b[j][1] + a[2][i] * b[j][2];
    return r;
}
GLMS_FN mat3 operator*(float s, const mat3& a) {
    mat3 r;
    for (int j = 0; j < 3; j++) r[j] = vec3(a[j].x * s, a[j].y * s, a[j].z * s);
    return r;
}
GLMS_FN mat3 operator*(const mat3& a, float s) { return s * a; }
GLMS_FN mat3 transpose(const mat3& a) {
    mat3 r;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) r[j][i] = a[i][j];
    return r;
}

}  // namespace glm
