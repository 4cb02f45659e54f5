/* entities/Sphere.cpp — UV-sphere entity and its CPU tessellation.
 *
 * Produces the same vertex / face arrays, bit for bit, as reference
 * src/lib/entities/Sphere.cpp:69-79,87-115,120-351 (checked by
 * tests/test_host_backend.py against the compiled reference):
 *   vertex 0 = north pole, then rings 1..p-2 of m vertices, last = south pole;
 *   faces    = m north-cap triangles, 2m per band between rings, m south-cap;
 *   point(x, y) = centre + r * (sin(pi y/(p-1)) cos(2 pi x/m), cos(pi y/(p-1)), sin(..) sin(..))
 *                 with the trigonometry in double and the ratios in float;
 *   normal   = normalize(cross(c - a, b - a)), colour = colour * |n . (0,0,-1)|.
 */
#include <cmath>

#include "Sphere.hpp"

using namespace RayTracer;

namespace {
    const double kPi = 3.14159265358979323846;

    glm::vec3 sphere_point(const ECS::Sphere& s, uint32_t x, uint32_t y) {
        const float v = (float) y / (float) (s.n_parallels - 1);
        const float u = (float) x / (float) s.n_meridians;
        const double polar = kPi * v, azimuth = 2 * kPi * u;
        return s.center + s.radius * glm::vec3(std::sin(polar) * std::cos(azimuth), std::cos(polar), std::sin(polar) * std::sin(azimuth));
    }

    void emit_face(GFace& f, uint32_t ia, uint32_t ib, uint32_t ic, const glm::vec3& a, const glm::vec3& b, const glm::vec3& c, const glm::vec3& color) {
        f.v1 = ia; f.v2 = ib; f.v3 = ic;
        f.normal = glm::normalize(glm::cross(c - a, b - a));
        f.color = color * std::abs(glm::dot(f.normal, glm::vec3(0.0f, 0.0f, -1.0f)));
    }
}

ECS::Sphere* ECS::create_sphere(const glm::vec3& center, float radius, uint32_t n_meridians, uint32_t n_parallels, const glm::vec3& color) {
    Sphere* s = new Sphere;
    s->type = et_sphere;
    s->pre_render_mode = eprmf_cpu;
    s->pre_render_operation = epro_generate_sphere;
    /* two caps of m triangles, 2m per band, p - 3 bands; unsigned arithmetic as in the reference (needs p >= 3) */
    s->pre_render_faces = n_meridians + 2 * ((n_parallels - 3) * n_meridians) + n_meridians;
    s->pre_render_vertices = 2 + (n_parallels - 2) * n_meridians;
    s->center = center;
    s->radius = radius;
    s->n_meridians = n_meridians;
    s->n_parallels = n_parallels;
    s->color = color;
    return s;
}

void ECS::cpu_pre_render_sphere(Tools::Array<GFace>& faces_buffer, Tools::Array<glm::vec4>& vertex_buffer, Sphere* sphere) {
    const uint32_t m = sphere->n_meridians, p = sphere->n_parallels;
    faces_buffer.resize(sphere->pre_render_faces);
    vertex_buffer.resize(sphere->pre_render_vertices);

    /* vertices: pole, rings, pole */
    auto ring_index = [m](uint32_t y, uint32_t x) { return 1 + (y - 1) * m + x; };
    const uint32_t south = 1 + (p - 2) * m;
    vertex_buffer[0] = glm::vec4(sphere_point(*sphere, 0, 0), 0.0f);
    for (uint32_t y = 1; y + 1 < p; y++) {
        for (uint32_t x = 0; x < m; x++) { vertex_buffer[ring_index(y, x)] = glm::vec4(sphere_point(*sphere, x, y), 0.0f); }
    }
    vertex_buffer[south] = glm::vec4(sphere_point(*sphere, 0, p - 1), 0.0f);
    auto vtx = [&vertex_buffer](uint32_t i) { return glm::vec3(vertex_buffer[i].x, vertex_buffer[i].y, vertex_buffer[i].z); };

    uint32_t out = 0;
    /* north cap: (pole, previous, current) */
    for (uint32_t x = 0; x < m; x++) {
        const uint32_t prev = ring_index(1, x > 0 ? x - 1 : m - 1), cur = ring_index(1, x);
        emit_face(faces_buffer[out++], 0, prev, cur, vtx(0), vtx(prev), vtx(cur), sphere->color);
    }
    /* bands: the quad (a b / c d) between ring y-1 and ring y splits into (a, c, d) and (a, b, d) */
    for (uint32_t y = 2; y + 1 < p; y++) {
        for (uint32_t x = 0; x < m; x++) {
            const uint32_t xp = x > 0 ? x - 1 : m - 1;
            const uint32_t a = ring_index(y - 1, xp), b = ring_index(y - 1, x), c = ring_index(y, xp), d = ring_index(y, x);
            emit_face(faces_buffer[out++], a, c, d, vtx(a), vtx(c), vtx(d), sphere->color);
            emit_face(faces_buffer[out++], a, b, d, vtx(a), vtx(b), vtx(d), sphere->color);
        }
    }
    /* south cap: (pole, previous, current) on the last ring */
    for (uint32_t x = 0; x < m; x++) {
        const uint32_t prev = ring_index(p - 2, x > 0 ? x - 1 : m - 1), cur = ring_index(p - 2, x);
        emit_face(faces_buffer[out++], south, prev, cur, vtx(south), vtx(prev), vtx(cur), sphere->color);
    }
}
