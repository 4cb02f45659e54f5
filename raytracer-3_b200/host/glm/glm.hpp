/* glm/glm.hpp — the small subset of the GLM interface the host backend uses.
 *
 * The reference vendors GLM 0.9.9.8 (reference src/lib/glm). When the backend is
 * built inside the reference tree the real headers are used
 * and this file is not on the include path; standalone, this subset keeps the
 * same names and the same fp32 arithmetic order (dot = (x+y)+z,
 * cross as glm/detail/func_geometric.inl:74-77, normalize = v * (1/sqrt(dot)))
 * so that flattened scenes are bit-identical either way.
 */
#ifndef RT3_HOST_GLM_SUBSET_HPP
#define RT3_HOST_GLM_SUBSET_HPP

#include <cmath>
#include <cstdint>

namespace glm {
    struct vec3 {
        float x, y, z;
        vec3() : x(0.0f), y(0.0f), z(0.0f) {}
        explicit vec3(float s) : x(s), y(s), z(s) {}
        vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
        vec3(double x_, double y_, double z_) : x((float) x_), y((float) y_), z((float) z_) {}
    };

    struct vec4 {
        float x, y, z, w;
        vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
        vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
        vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
        vec4(const vec3& v, double w_) : x(v.x), y(v.y), z(v.z), w((float) w_) {}
        explicit operator vec3() const { return vec3(x, y, z); }
    };

    inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
    inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
    inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
    inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
    inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
    inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
    inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
    inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
    inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }

    inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
    inline vec3 cross(const vec3& x, const vec3& y) {
        return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
    }
    inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
    inline vec3 normalize(const vec3& a) { return a * (1.0f / std::sqrt(dot(a, a))); }
    inline float abs(float v) { return std::fabs(v); }
}

#endif
