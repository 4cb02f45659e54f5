/* rt3_device.cuh — device-side building blocks of the render core (sm_100a).
 *
 * Arithmetic contract. This translation unit is compiled with -fmad=false and
 * without fast-math, so every C expression below is one IEEE-754 binary32
 * operation in the order written: the "exact" routines reproduce the CPU
 * reference bit for bit (reference src/lib/renderer/SequentialRenderer.cpp:47-109,
 * operation order in SURVEY.md appendix A). Fused multiply-adds appear only
 * where they are written explicitly (__fmaf_rn) — in the conservative
 * bounding-sphere prefilter, whose only job is to decide which primitives get
 * the exact test and which can never change the result.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rt3cuda.h"
#include "rt3_rng.h"

#define RT3_NO_HIT 0xFFFFFFFFu
#define RT3_TMIN 0.001f
#define RT3_ACC_SCALE 16777216.0f

#define RT3_BLOCK_PRIMS 32      /* primitives per candidate-mask block */
#define RT3_TILE_PRIMS 2048     /* primitives per streamed shared-memory tile (32 KB) */
#define RT3_RESIDENT_PRIMS 4096 /* scenes up to this size live in shared memory for the whole kernel (64 KB) */
#define RT3_CTA_THREADS 128
#define RT3_CTAS_PER_SM 5       /* register budget: 65536 / (128 * 5) = 102 per thread */
#define RT3_ITEM_CHUNK 1024u    /* path items a warp claims per global atomic */
#define RT3_CAND_CAP 16         /* per-ray deferred-candidate list entries (shared memory, 16-bit tile-relative ids) */

/* Relative slack folded into the prefilter (64 ulp of binary32): covers the
 * rounding of the prefilter's own FMA chains plus that of the exact sphere
 * test it stands in front of, both bounded by ~28 eps (|c|^2 + |o|^2). */
#define RT3_FILTER_SLACK 3.814697265625e-06f

struct rt3_vec3 { float x, y, z; };

__device__ __forceinline__ rt3_vec3 v3(float x, float y, float z) { rt3_vec3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ rt3_vec3 operator+(rt3_vec3 a, rt3_vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ rt3_vec3 operator-(rt3_vec3 a, rt3_vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ rt3_vec3 operator*(rt3_vec3 a, rt3_vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ rt3_vec3 operator*(float s, rt3_vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ rt3_vec3 operator-(rt3_vec3 a) { return v3(-a.x, -a.y, -a.z); }
/* dot3 macro, SequentialRenderer.cpp:32-33. */
__device__ __forceinline__ float dot3(rt3_vec3 a, rt3_vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* glm::cross, glm/detail/func_geometric.inl:74-77. */
__device__ __forceinline__ rt3_vec3 cross3(rt3_vec3 x, rt3_vec3 y) {
    return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
/* glm::normalize = v * (1 / sqrt(dot(v, v))). */
__device__ __forceinline__ rt3_vec3 normalize3(rt3_vec3 a) {
    float inv = 1.0f / sqrtf(dot3(a, a));
    return v3(a.x * inv, a.y * inv, a.z * inv);
}

/* glm::packUnorm4x8 channel (func_packing.inl:67-83): round(clamp(c,0,1)*255), half away from zero. */
__device__ __forceinline__ uint32_t unorm8(float c) {
    float m = (c < 0.0f) ? 0.0f : c;
    m = (1.0f < m) ? 1.0f : m;
    if (m != m) { m = 0.0f; }
    return (uint32_t) roundf(m * 255.0f);
}
__device__ __forceinline__ uint32_t pack_rgb(rt3_vec3 c) {
    return (unorm8(c.x) << 24) | (unorm8(c.y) << 16) | (unorm8(c.z) << 8) | 0xFFu;
}

/* Device view of the flattened scene (all pointers are global memory). */
struct rt3_scene_view {
    uint32_t n_faces;
    uint32_t n_spheres;
    uint32_t n_prims;        /* n_faces + n_spheres; primitive id = face index, then n_faces + sphere index */
    uint32_t n_prims_padded; /* rounded up to RT3_BLOCK_PRIMS with never-hit entries */
    const float4* bounds;    /* per primitive: prefilter sphere (cx, cy, cz, |c|^2 - r_eff^2 - slack) */
    const float4* face_n;    /* per face: (nx, ny, nz, dot3(n, p1)) */
    const float4* face_p1;   /* per face: p1.xyz */
    const float4* face_p2;
    const float4* face_p3;
    const float4* spheres;   /* per sphere: (cx, cy, cz, r) */
    const float4* prim_color;     /* per primitive: flat colour / albedo */
    const uint32_t* prim_material; /* per primitive: index into materials, or RT3_NO_HIT for Lambertian(prim_color) */
    const uint32_t* prim_entity;
    const float* prim_radius; /* per primitive (padded): sphere radius, 0 for faces */
    const float4* materials; /* 2 float4 per material: (kind bits, albedo rgb), (fuzz, ior, -, -) */
};

struct rt3_hit { float t; uint32_t prim; };

/* Per-ray constants of the prefilter. For a sphere (c, r) and a ray (o, unit dn):
 *   h    = (c - o) . dn            = c.dn - o.dn
 *   q    = |c - o|^2 - r^2         = (|c|^2 - r^2) + |o|^2 - 2 c.o
 *   disc = h^2 - q  >= 0  <=>  the line meets the sphere.
 * Three FFMA for h, one FADD + three FFMA for q, one FFMA for disc. */
struct rt3_ray_filter { float dx, dy, dz, nod, m2ox, m2oy, m2oz, oo; };

__device__ __forceinline__ rt3_ray_filter make_ray_filter(rt3_vec3 o, rt3_vec3 dn) {
    rt3_ray_filter f;
    f.dx = dn.x; f.dy = dn.y; f.dz = dn.z;
    f.nod = -dot3(o, dn);
    f.m2ox = -2.0f * o.x; f.m2oy = -2.0f * o.y; f.m2oz = -2.0f * o.z;
    float oo = dot3(o, o);
    f.oo = oo - RT3_FILTER_SLACK * oo; /* lowering q can only add candidates */
    return f;
}

/* One prefilter test; returns disc (sign bit set <=> certainly no hit). */
__device__ __forceinline__ float filter_disc(const float4 b, const rt3_ray_filter& f) {
    float h = __fmaf_rn(b.x, f.dx, __fmaf_rn(b.y, f.dy, __fmaf_rn(b.z, f.dz, f.nod)));
    float q = __fmaf_rn(b.x, f.m2ox, __fmaf_rn(b.y, f.m2oy, __fmaf_rn(b.z, f.m2oz, b.w + f.oo)));
    return __fmaf_rn(h, h, -q);
}

/* Exact ray-triangle test: the body of the reference's face loop,
 * SequentialRenderer.cpp:55-95, for one candidate face. tmin = 0 reproduces
 * the reference (`t < 0` rejected); the bounce loop passes 0.001. */
__device__ __forceinline__ void exact_face(const rt3_scene_view& S, uint32_t i, rt3_vec3 o, rt3_vec3 d, float tmin, rt3_hit& best) {
    float4 fn = __ldg(&S.face_n[i]);
    rt3_vec3 n = v3(fn.x, fn.y, fn.z);
    float nd = dot3(d, n);
    if (nd == 0) { return; }
    float4 a1 = __ldg(&S.face_p1[i]), a2 = __ldg(&S.face_p2[i]), a3 = __ldg(&S.face_p3[i]);
    rt3_vec3 p1 = v3(a1.x, a1.y, a1.z), p2 = v3(a2.x, a2.y, a2.z), p3 = v3(a3.x, a3.y, a3.z);
    float pd = fn.w; /* dot3(n, p1), evaluated once at upload in the same order */
    float t = (pd - dot3(n, o)) / dot3(n, d);
    if (t < tmin || t >= best.t) { return; }
    rt3_vec3 hp = o + t * d;
    rt3_vec3 a = cross3(p2 - p1, hp - p1);
    rt3_vec3 b = cross3(p3 - p2, hp - p2);
    rt3_vec3 c = cross3(p1 - p3, hp - p3);
    if (-dot3(n, a) >= 0.0f && -dot3(n, b) >= 0.0f && -dot3(n, c) >= 0.0f) { best.prim = i; best.t = t; }
}

/* Exact ray-sphere test, reference mode: WIP hit_sphere, raytracer_v4.glsl:157-178
 * (abc form, near root, t >= 0), un-normalised direction. */
__device__ __forceinline__ void exact_sphere_v4(const rt3_scene_view& S, uint32_t si, float4 sp, rt3_vec3 o, rt3_vec3 d, rt3_hit& best) {
    rt3_vec3 oc = o - v3(sp.x, sp.y, sp.z);
    float a = dot3(d, d);
    float b = 2.0f * dot3(oc, d);
    float c = dot3(oc, oc) - sp.w * sp.w;
    float D = b * b - (4.0f * a) * c;
    if (D >= 0) {
        float t = (-b - sqrtf(D)) / (2.0f * a);
        if (t >= 0.0f && t < best.t) { best.prim = S.n_faces + si; best.t = t; }
    }
}

/* Exact ray-sphere test, bounce loop: half-b form with a unit direction, near
 * then far root, accepted iff tmin <= t < best. */
__device__ __forceinline__ void exact_sphere_path(const rt3_scene_view& S, uint32_t si, float4 sp, rt3_vec3 o, rt3_vec3 d, rt3_hit& best) {
    rt3_vec3 oc = o - v3(sp.x, sp.y, sp.z);
    float h = dot3(oc, d);
    float c = dot3(oc, oc) - sp.w * sp.w;
    float disc = h * h - c;
    if (!(disc >= 0.0f)) { return; }
    float sq = sqrtf(disc);
    float t = -h - sq;
    if (!(t >= RT3_TMIN && t < best.t)) {
        t = -h + sq;
        if (!(t >= RT3_TMIN && t < best.t)) { return; }
    }
    best.prim = S.n_faces + si; best.t = t;
}

/* Exact test of one candidate primitive. `sp` is the sphere record (centre,
 * radius) when the primitive is a sphere; the caller fetches it from shared
 * memory when the scene is resident, from global memory otherwise. */
template <bool PATH_MODE>
__device__ __forceinline__ void exact_prim(const rt3_scene_view& S, uint32_t prim, float4 sp, rt3_vec3 o, rt3_vec3 d, rt3_hit& best) {
    if (prim < S.n_faces) {
        exact_face(S, prim, o, d, PATH_MODE ? RT3_TMIN : 0.0f, best);
    } else if (prim < S.n_prims) {
        if (PATH_MODE) { exact_sphere_path(S, prim - S.n_faces, sp, o, d, best); }
        else { exact_sphere_v4(S, prim - S.n_faces, sp, o, d, best); }
    }
}

/* Where the sweep finds its data: the prefilter tile in shared memory, the
 * per-primitive radii (shared memory for resident scenes, else NULL) and the
 * per-thread deferred-candidate lists. */
struct rt3_tile_view {
    const float4* bounds;   /* shared: prefilter records of this tile */
    const float* radius;    /* shared: sphere radii of this tile, or NULL (fetch S.spheres from global) */
    uint16_t* cand;         /* shared: [R][RT3_CAND_CAP][RT3_CTA_THREADS] tile-relative candidate ids */
    uint32_t first_prim;    /* global id of the tile's first primitive */
};

template <bool PATH_MODE>
__device__ __forceinline__ void exact_candidate(const rt3_scene_view& S, const rt3_tile_view& T, uint32_t rel, rt3_vec3 o, rt3_vec3 d, rt3_hit& best) {
    const uint32_t prim = T.first_prim + rel;
    float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (prim >= S.n_faces && prim < S.n_prims) {
        if (T.radius) { sp = T.bounds[rel]; sp.w = T.radius[rel]; }   /* the prefilter centre of a sphere is its own centre */
        else { sp = __ldg(&S.spheres[prim - S.n_faces]); }
    }
    exact_prim<PATH_MODE>(S, prim, sp, o, d, best);
}

/* Filters one block of RT3_BLOCK_PRIMS primitives (shared memory, broadcast
 * LDS.128) against R rays. Survivors are appended to the ray's deferred list
 * (ascending primitive order); when a list is full the exact test runs at
 * once, which is still in ascending order. */
template <int R, bool PATH_MODE>
__device__ __forceinline__ void sweep_block(const rt3_scene_view& S, const rt3_tile_view& T, uint32_t rel_base,
                                            const rt3_ray_filter (&f)[R], const rt3_vec3 (&o)[R], const rt3_vec3 (&d)[R],
                                            const bool (&live)[R], uint32_t (&n_cand)[R], rt3_hit (&best)[R]) {
    const float4* __restrict__ blk = T.bounds + rel_base;
    uint32_t miss[R];
#pragma unroll
    for (int r = 0; r < R; r++) { miss[r] = 0u; }
#pragma unroll 8
    for (int j = 0; j < RT3_BLOCK_PRIMS; j++) {
        const float4 b = blk[j];
#pragma unroll
        for (int r = 0; r < R; r++) {
            float disc = filter_disc(b, f[r]);
            /* shift the sign bit of disc into the mask: bit (31 - j) set <=> primitive j cannot be hit */
            miss[r] = __funnelshift_l(__float_as_uint(disc), miss[r], 1);
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        uint32_t cand = live[r] ? ~miss[r] : 0u;
        while (cand) {
            int j = __clz(cand);
            cand &= ~(0x80000000u >> j);
            const uint32_t rel = rel_base + (uint32_t) j;
            if (n_cand[r] < RT3_CAND_CAP) {
                T.cand[(r * RT3_CAND_CAP + n_cand[r]) * RT3_CTA_THREADS + threadIdx.x] = (uint16_t) rel;
                n_cand[r]++;
            } else {
                /* list full: drain it first so that the order of exact tests stays ascending */
                for (uint32_t e = 0; e < RT3_CAND_CAP; e++) {
                    exact_candidate<PATH_MODE>(S, T, T.cand[(r * RT3_CAND_CAP + e) * RT3_CTA_THREADS + threadIdx.x], o[r], d[r], best[r]);
                }
                T.cand[(r * RT3_CAND_CAP) * RT3_CTA_THREADS + threadIdx.x] = (uint16_t) rel;
                n_cand[r] = 1;
            }
        }
    }
}

/* Runs the exact tests of every deferred candidate of this tile. */
template <int R, bool PATH_MODE>
__device__ __forceinline__ void drain_candidates(const rt3_scene_view& S, const rt3_tile_view& T, const rt3_vec3 (&o)[R],
                                                 const rt3_vec3 (&d)[R], uint32_t (&n_cand)[R], rt3_hit (&best)[R]) {
#pragma unroll
    for (int r = 0; r < R; r++) {
        for (uint32_t e = 0; e < n_cand[r]; e++) {
            exact_candidate<PATH_MODE>(S, T, T.cand[(r * RT3_CAND_CAP + e) * RT3_CTA_THREADS + threadIdx.x], o[r], d[r], best[r]);
        }
        n_cand[r] = 0;
    }
}

/* ---- mbarrier / bulk-copy (TMA) helpers for streamed tiles ---------------- */

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* 1-D bulk asynchronous copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP). */
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
