/* rtiow_book.cpp — SECOND, INDEPENDENT CPU checker for the part of the path the reference does not implement.
 *
 * TEST INFRASTRUCTURE (like everything under oracle/): loaded only by tests/. It exists because the first
 * checker (rt3_oracle.c, orc_render_pathtrace) was written alongside the CUDA kernel and shares its sampling
 * header (include/rt3_rng.h) and its formulation of the sphere test, so "GPU == rt3_oracle.c" proves the kernel
 * equals its twin, not that the twin renders what the reference intends. The reference's intention for rows
 * a6 / a9 / a10 of SURVEY.md section 8 is stated in three places:
 *   - README.md:2                       "Based on ... Ray Tracing in One Weekend" (Shirley, book 1)
 *   - raytracer_v4.glsl:157-178         hit_sphere in the abc form (b*b - 4*a*c, near root)
 *   - raytracer_v4.glsl:183-283         the bounce loop skeleton (obj_material, "let's not bounce just yet" :279)
 *   - reduce_v1.glsl:66-76              the (empty) per-pixel reduction over samples
 * This file restates the book's program the way the book writes it and shares NOTHING with the product or with
 * rt3_oracle.c: double precision, recursive ray_color, rejection sampling for random_in_unit_sphere /
 * random_in_unit_disk, std::mt19937 + libm, the abc-form discriminant with both roots, Moeller-Trumbore for
 * triangles (the reference's plane + inside-out test is pinned separately, bit for bit). It includes
 * rt3cuda.h only for the plain input structs (scene, camera) -- no arithmetic comes from there.
 *
 * Because the random numbers differ, agreement with the GPU path is statistical: tests/test_book_pin.py and
 * tests/test_gpu_book_pin.py compare converged images (PSNR >= 40 dB), per-material energy (white furnace) and
 * scenes that need the far root of the sphere test. Tolerances are stated there and in DESIGN.md section 6.
 */
#include <atomic>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <thread>
#include <vector>

#include "rt3cuda.h"

namespace {

struct vec3 {
    double e[3];
    vec3() : e{ 0, 0, 0 } {}
    vec3(double a, double b, double c) : e{ a, b, c } {}
    double x() const { return e[0]; }
    double y() const { return e[1]; }
    double z() const { return e[2]; }
    vec3 operator-() const { return vec3(-e[0], -e[1], -e[2]); }
    vec3& operator+=(const vec3& v) { e[0] += v.e[0]; e[1] += v.e[1]; e[2] += v.e[2]; return *this; }
    double length_squared() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
    double length() const { return std::sqrt(length_squared()); }
    bool near_zero() const { const double s = 1e-8; return std::fabs(e[0]) < s && std::fabs(e[1]) < s && std::fabs(e[2]) < s; }
};
using point3 = vec3;
using color = vec3;

inline vec3 operator+(const vec3& u, const vec3& v) { return vec3(u.e[0] + v.e[0], u.e[1] + v.e[1], u.e[2] + v.e[2]); }
inline vec3 operator-(const vec3& u, const vec3& v) { return vec3(u.e[0] - v.e[0], u.e[1] - v.e[1], u.e[2] - v.e[2]); }
inline vec3 operator*(const vec3& u, const vec3& v) { return vec3(u.e[0] * v.e[0], u.e[1] * v.e[1], u.e[2] * v.e[2]); }
inline vec3 operator*(double t, const vec3& v) { return vec3(t * v.e[0], t * v.e[1], t * v.e[2]); }
inline vec3 operator/(const vec3& v, double t) { return (1.0 / t) * v; }
inline double dot(const vec3& u, const vec3& v) { return u.e[0] * v.e[0] + u.e[1] * v.e[1] + u.e[2] * v.e[2]; }
inline vec3 cross(const vec3& u, const vec3& v) {
    return vec3(u.e[1] * v.e[2] - u.e[2] * v.e[1], u.e[2] * v.e[0] - u.e[0] * v.e[2], u.e[0] * v.e[1] - u.e[1] * v.e[0]);
}
inline vec3 unit_vector(const vec3& v) { return v / v.length(); }
inline vec3 from3(const float* p) { return vec3(p[0], p[1], p[2]); }

struct Rng {
    std::mt19937 gen;
    std::uniform_real_distribution<double> dist{ 0.0, 1.0 };
    explicit Rng(uint32_t seed) : gen(seed) {}
    double next() { return dist(gen); }
    double next(double lo, double hi) { return lo + (hi - lo) * next(); }
    vec3 in_unit_sphere() {
        for (;;) {
            vec3 p(next(-1, 1), next(-1, 1), next(-1, 1));
            if (p.length_squared() >= 1) { continue; }
            return p;
        }
    }
    vec3 unit() { return unit_vector(in_unit_sphere()); }
    vec3 in_unit_disk() {
        for (;;) {
            vec3 p(next(-1, 1), next(-1, 1), 0);
            if (p.length_squared() >= 1) { continue; }
            return p;
        }
    }
};

struct ray {
    point3 orig;
    vec3 dir;
    point3 at(double t) const { return orig + t * dir; }
};

struct material_rec { uint32_t kind; color albedo; double fuzz, ior; };

struct hit_record {
    point3 p;
    vec3 normal;
    double t = 0;
    bool front_face = false;
    material_rec mat{};
    void set_face_normal(const ray& r, const vec3& outward_normal) {
        front_face = dot(r.dir, outward_normal) < 0;
        normal = front_face ? outward_normal : -outward_normal;
    }
};

struct World {
    const rt3_scene* s;
    material_rec material_of_sphere(uint32_t i) const {
        if (s->sphere_material) { const rt3_material& m = s->materials[s->sphere_material[i]]; return { m.kind, from3(m.albedo), m.fuzz, m.ior }; }
        return { RT3_MAT_LAMBERTIAN, from3(&s->sphere_color[3 * i]), 0.0, 1.0 };
    }
    material_rec material_of_face(uint32_t i) const {
        if (s->face_material) { const rt3_material& m = s->materials[s->face_material[i]]; return { m.kind, from3(m.albedo), m.fuzz, m.ior }; }
        return { RT3_MAT_LAMBERTIAN, from3(s->faces[i].color), 0.0, 1.0 };
    }
    /* hit_sphere as raytracer_v4.glsl:157-178 writes it (a, b, c and the full discriminant), extended by the far
     * root and the [t_min, t_max] window of the book's sphere::hit. */
    bool hit_sphere(uint32_t i, const ray& r, double t_min, double t_max, hit_record& rec) const {
        const rt3_sphere& sp = s->spheres[i];
        const point3 center(sp.cx, sp.cy, sp.cz);
        const double radius = sp.r;
        const vec3 oc = r.orig - center;
        const double a = dot(r.dir, r.dir);
        const double b = 2.0 * dot(oc, r.dir);
        const double c = dot(oc, oc) - radius * radius;
        const double discriminant = b * b - 4 * a * c;
        if (discriminant < 0) { return false; }
        const double sqrtd = std::sqrt(discriminant);
        double root = (-b - sqrtd) / (2.0 * a);
        if (root < t_min || t_max < root) {
            root = (-b + sqrtd) / (2.0 * a);
            if (root < t_min || t_max < root) { return false; }
        }
        rec.t = root;
        rec.p = r.at(rec.t);
        rec.set_face_normal(r, (rec.p - center) / radius); /* a negative radius turns the normal inward: the hollow-glass trick */
        rec.mat = material_of_sphere(i);
        return true;
    }
    /* Moeller-Trumbore; the shading normal is the stored face normal, as in the reference (GFace.normal). */
    bool hit_triangle(uint32_t i, const ray& r, double t_min, double t_max, hit_record& rec) const {
        const rt3_face& f = s->faces[i];
        const point3 v0 = from3(&s->vertices[f.v1].x), v1 = from3(&s->vertices[f.v2].x), v2 = from3(&s->vertices[f.v3].x);
        const vec3 e1 = v1 - v0, e2 = v2 - v0;
        const vec3 pvec = cross(r.dir, e2);
        const double det = dot(e1, pvec);
        if (det == 0.0) { return false; }
        const double inv = 1.0 / det;
        const vec3 tvec = r.orig - v0;
        const double u = dot(tvec, pvec) * inv;
        if (u < 0.0 || u > 1.0) { return false; }
        const vec3 qvec = cross(tvec, e1);
        const double v = dot(r.dir, qvec) * inv;
        if (v < 0.0 || u + v > 1.0) { return false; }
        const double t = dot(e2, qvec) * inv;
        if (t < t_min || t_max < t) { return false; }
        rec.t = t;
        rec.p = r.at(t);
        rec.set_face_normal(r, from3(f.normal));
        rec.mat = material_of_face(i);
        return true;
    }
    bool hit(const ray& r, double t_min, double t_max, hit_record& rec) const {
        hit_record temp;
        bool any = false;
        double closest = t_max;
        for (uint32_t i = 0; i < s->n_faces; i++) { if (hit_triangle(i, r, t_min, closest, temp)) { any = true; closest = temp.t; rec = temp; } }
        for (uint32_t i = 0; i < s->n_spheres; i++) { if (hit_sphere(i, r, t_min, closest, temp)) { any = true; closest = temp.t; rec = temp; } }
        return any;
    }
};

inline vec3 reflect(const vec3& v, const vec3& n) { return v - 2 * dot(v, n) * n; }
inline vec3 refract(const vec3& uv, const vec3& n, double etai_over_etat) {
    const double cos_theta = std::fmin(dot(-uv, n), 1.0);
    const vec3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    const vec3 r_out_parallel = -std::sqrt(std::fabs(1.0 - r_out_perp.length_squared())) * n;
    return r_out_perp + r_out_parallel;
}
inline double reflectance(double cosine, double ref_idx) {
    double r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    return r0 + (1 - r0) * std::pow(1 - cosine, 5);
}

bool scatter(const material_rec& m, const ray& r_in, const hit_record& rec, color& attenuation, ray& scattered, Rng& rng) {
    if (m.kind == RT3_MAT_LAMBERTIAN) {
        vec3 scatter_direction = rec.normal + rng.unit();
        if (scatter_direction.near_zero()) { scatter_direction = rec.normal; }
        scattered = ray{ rec.p, scatter_direction };
        attenuation = m.albedo;
        return true;
    }
    if (m.kind == RT3_MAT_METAL) {
        const vec3 reflected = reflect(unit_vector(r_in.dir), rec.normal);
        const double fuzz = m.fuzz < 1 ? m.fuzz : 1;
        scattered = ray{ rec.p, reflected + fuzz * rng.in_unit_sphere() };
        attenuation = m.albedo;
        return dot(scattered.dir, rec.normal) > 0;
    }
    attenuation = color(1.0, 1.0, 1.0);
    const double refraction_ratio = rec.front_face ? (1.0 / m.ior) : m.ior;
    const vec3 unit_direction = unit_vector(r_in.dir);
    const double cos_theta = std::fmin(dot(-unit_direction, rec.normal), 1.0);
    const double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
    const bool cannot_refract = refraction_ratio * sin_theta > 1.0;
    vec3 direction;
    if (cannot_refract || reflectance(cos_theta, refraction_ratio) > rng.next()) { direction = reflect(unit_direction, rec.normal); }
    else { direction = refract(unit_direction, rec.normal, refraction_ratio); }
    scattered = ray{ rec.p, direction };
    return true;
}

color ray_color(const ray& r, const World& world, int depth, bool uniform_sky, Rng& rng) {
    hit_record rec;
    if (depth <= 0) { return color(0, 0, 0); }
    if (world.hit(r, 0.001, std::numeric_limits<double>::infinity(), rec)) {
        ray scattered;
        color attenuation;
        if (scatter(rec.mat, r, rec, attenuation, scattered, rng)) { return attenuation * ray_color(scattered, world, depth - 1, uniform_sky, rng); }
        return color(0, 0, 0);
    }
    if (uniform_sky) { return color(1, 1, 1); }
    /* the sky of SequentialRenderer.cpp:105-107 == the book's */
    const vec3 unit_direction = unit_vector(r.dir);
    const double t = 0.5 * (unit_direction.y() + 1.0);
    return (1.0 - t) * color(1.0, 1.0, 1.0) + t * color(0.5, 0.7, 1.0);
}

/* The book's camera::get_ray on the four vectors the reference Camera publishes (Camera.hpp:27-34) plus the lens. */
ray get_ray(const rt3_camera& cam, double s, double t, Rng& rng) {
    const point3 origin = from3(cam.origin), llc = from3(cam.lower_left_corner);
    const vec3 horizontal = from3(cam.horizontal), vertical = from3(cam.vertical);
    vec3 offset(0, 0, 0);
    if (cam.lens_radius > 0) {
        const vec3 rd = (double) cam.lens_radius * rng.in_unit_disk();
        offset = rd.x() * from3(cam.lens_u) + rd.y() * from3(cam.lens_v);
    }
    return ray{ origin + offset, llc + s * horizontal + t * vertical - origin - offset };
}

}  // namespace

extern "C" {

/* Renders rows [first_row, first_row + n_rows) of a width x height image, `spp` samples per pixel, `max_depth` ray
 * segments per path; rgb receives the per-pixel MEAN LINEAR radiance (3 floats per pixel, full-frame indexing,
 * untouched outside the rows). flags: RT3_FLAG_UNIFORM_SKY, RT3_FLAG_NO_JITTER. Row 0 is the top row
 * (SequentialRenderer.cpp:286-297). n_threads <= 0: all cores. */
int book_render(const rt3_scene* scene, const rt3_camera* cam, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_depth, uint32_t seed,
                uint32_t flags, uint32_t first_row, uint32_t n_rows, float* rgb, int n_threads) {
    if (!scene || !cam || !rgb || width < 2 || height < 2 || spp < 1 || max_depth < 1 || first_row + n_rows > height) { return -1; }
    if (n_threads <= 0) { n_threads = (int) std::thread::hardware_concurrency(); }
    if (n_threads <= 0) { n_threads = 1; }
    const World world{ scene };
    const bool uniform_sky = (flags & RT3_FLAG_UNIFORM_SKY) != 0, jitter = !(flags & RT3_FLAG_NO_JITTER);
    std::atomic<uint32_t> next_row{ 0 };
    auto worker = [&]() {
        for (;;) {
            const uint32_t k = next_row.fetch_add(1);
            if (k >= n_rows) { break; }
            const uint32_t j = first_row + k;
            Rng rng(seed * 9781u + j * 6271u + 1u);
            for (uint32_t i = 0; i < width; i++) {
                color pixel(0, 0, 0);
                for (uint32_t sidx = 0; sidx < spp; sidx++) {
                    const double u = (i + (jitter ? rng.next() : 0.0)) / (width - 1);
                    const double v = ((height - 1 - j) + (jitter ? rng.next() : 0.0)) / (height - 1);
                    pixel += ray_color(get_ray(*cam, u, v, rng), world, (int) max_depth, uniform_sky, rng);
                }
                float* out = rgb + 3 * ((size_t) j * width + i);
                for (int c = 0; c < 3; c++) { out[c] = (float) (pixel.e[c] / spp); }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++) { pool.emplace_back(worker); }
    worker();
    for (auto& t : pool) { t.join(); }
    return 0;
}

}  // extern "C"
