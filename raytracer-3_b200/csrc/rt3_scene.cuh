/* rt3_scene.cuh — scene construction on the device.
 *
 * CUDA replacement for the reference's sphere pre-render: the CPU routine
 * cpu_pre_render_sphere (reference src/lib/entities/Sphere.cpp:69-79,120-351) and its
 * GPU twins pre_render_sphere_v2_vertices.glsl / pre_render_sphere_v2_faces.glsl
 * (src/lib/shaders/). One thread per vertex, then one thread per face, writing the
 * reference's flattened records (GFace 48 B, glm::vec4) at an offset, so that several
 * entities land in one vertex / face array like transfer_entity does
 * (SequentialRenderer.cpp:174-195: face indices are shifted by the entity's first vertex).
 *
 * Arithmetic follows the CPU routine, not the shader: the polar / azimuth angles and the
 * sin * cos products are evaluated in double and narrowed to float (Sphere.cpp:69-79) --
 * the reference's own shader uses float trigonometry and differs from its CPU path by
 * ~1e-6. Vertex layout: north pole, rings 1..p-2 of m vertices,
 * south pole; faces: m north-cap triangles, 2m per band, m south-cap triangles.
 */
#pragma once

#include "rt3_device.cuh"

struct rt3_uv_sphere_dev {
    float cx, cy, cz, radius;
    uint32_t m, p;            /* meridians, parallels (p >= 3, m >= 1) */
    float r, g, b;
    uint32_t first_vertex;    /* offset of this sphere's vertices in the output array */
    uint32_t first_face;
    uint32_t entity;
    uint32_t index_shift;     /* added to the vertex indices a face records (where the caller keeps the batch in its own vertex array) */
};

__host__ __device__ inline uint32_t uv_sphere_vertices(uint32_t m, uint32_t p) { return 2u + (p - 2u) * m; }
__host__ __device__ inline uint32_t uv_sphere_faces(uint32_t m, uint32_t p) { return 2u * m + 2u * (p - 3u) * m; }

/* Sphere.cpp:69-79. */
__device__ __forceinline__ rt3_vec3 uv_sphere_point(const rt3_uv_sphere_dev& s, uint32_t x, uint32_t y) {
    const double pi = 3.14159265358979323846;
    const float v = (float) y / (float) (s.p - 1u);
    const float u = (float) x / (float) s.m;
    const double polar = pi * (double) v, azimuth = 2 * pi * (double) u;
    const float dx = (float) (sin(polar) * cos(azimuth)), dy = (float) cos(polar), dz = (float) (sin(polar) * sin(azimuth));
    return v3(s.cx + s.radius * dx, s.cy + s.radius * dy, s.cz + s.radius * dz);
}

__global__ void uv_sphere_vertices_kernel(rt3_uv_sphere_dev s, float4* __restrict__ vertices) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = uv_sphere_vertices(s.m, s.p);
    if (i >= n) { return; }
    uint32_t x = 0, y = 0;
    if (i == n - 1u) { y = s.p - 1u; }
    else if (i > 0u) { y = 1u + (i - 1u) / s.m; x = (i - 1u) - (y - 1u) * s.m; }
    const rt3_vec3 pt = uv_sphere_point(s, x, y);
    vertices[s.first_vertex + i] = make_float4(pt.x, pt.y, pt.z, 0.0f);
}

/* The 48-byte GFace record (reference renderer/Vertex.hpp:39-51) as three 16-byte stores. */
__device__ __forceinline__ void store_face(uint4* __restrict__ faces, uint32_t f, uint32_t ia, uint32_t ib, uint32_t ic, rt3_vec3 n, rt3_vec3 color) {
    faces[3 * (size_t) f + 0] = make_uint4(ia, ib, ic, 0u);
    faces[3 * (size_t) f + 1] = make_uint4(__float_as_uint(n.x), __float_as_uint(n.y), __float_as_uint(n.z), 0u);
    faces[3 * (size_t) f + 2] = make_uint4(__float_as_uint(color.x), __float_as_uint(color.y), __float_as_uint(color.z), 0u);
}

__global__ void uv_sphere_faces_kernel(rt3_uv_sphere_dev s, const float4* __restrict__ vertices, uint4* __restrict__ faces,
                                       uint32_t* __restrict__ face_entity) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = s.m, p = s.p, n = uv_sphere_faces(m, p);
    if (f >= n) { return; }
    const uint32_t south = 1u + (p - 2u) * m;
    uint32_t ia, ib, ic;
    if (f < m) {                                   /* north cap: (pole, previous, current) */
        const uint32_t x = f, xp = x > 0u ? x - 1u : m - 1u;
        ia = 0u; ib = 1u + xp; ic = 1u + x;
    } else if (f >= n - m) {                       /* south cap on the last ring */
        const uint32_t x = f - (n - m), xp = x > 0u ? x - 1u : m - 1u, ring = 1u + (p - 3u) * m;
        ia = south; ib = ring + xp; ic = ring + x;
    } else {                                       /* bands: quad (a b / c d) -> (a, c, d), (a, b, d) */
        const uint32_t q = (f - m) / 2u, second = (f - m) & 1u;
        const uint32_t y = 2u + q / m, x = q - (y - 2u) * m, xp = x > 0u ? x - 1u : m - 1u;
        const uint32_t a = 1u + (y - 2u) * m + xp, b = 1u + (y - 2u) * m + x, c = 1u + (y - 1u) * m + xp, d = 1u + (y - 1u) * m + x;
        ia = a; ib = second ? b : c; ic = d;
    }
    const float4 A = vertices[s.first_vertex + ia], B = vertices[s.first_vertex + ib], C = vertices[s.first_vertex + ic];
    const rt3_vec3 a = v3(A.x, A.y, A.z), b = v3(B.x, B.y, B.z), c = v3(C.x, C.y, C.z);
    /* Sphere.cpp:153-155: normal = normalize(cross(c - a, b - a)), colour = colour * |n . (0, 0, -1)| */
    const rt3_vec3 nrm = normalize3(cross3(c - a, b - a));
    const float shade = fabsf((nrm.x * 0.0f + nrm.y * 0.0f) + nrm.z * -1.0f);
    store_face(faces, s.first_face + f, s.index_shift + s.first_vertex + ia, s.index_shift + s.first_vertex + ib, s.index_shift + s.first_vertex + ic, nrm,
               v3(s.r * shade, s.g * shade, s.b * shade));
    if (face_entity) { face_entity[s.first_face + f] = s.entity; }
}
