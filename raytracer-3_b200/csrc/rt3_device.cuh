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

#define RT3_RAYS 2              /* rays (path slots) per thread */
#define RT3_BLOCK_PRIMS 32      /* primitives per candidate-mask block */
#define RT3_PAD_PRIMS 8         /* the primitive array is padded to a multiple of this with never-hit records */
#define RT3_REC_BYTES 16        /* prefilter record: (cx, cy, cz, -R^2) */
#define RT3_TILE_PRIMS 2048     /* primitives per streamed shared-memory tile (32 KB) */
#define RT3_RESIDENT_PRIMS 4096 /* scenes up to this size live in shared memory for the whole kernel (64 KB) */
#define RT3_CTA_THREADS 128
#define RT3_CTAS_PER_SM 5       /* register budget: 65536 / (128 * 5) = 102 per thread */
#define RT3_ITEM_CHUNK 1024u    /* path items a warp claims per global atomic */
#define RT3_CAND_CAP 32         /* per-ray deferred-candidate list entries (shared memory, 16-bit tile-relative ids) */

/* Relative slack of the prefilter (64 ulp of binary32), applied to |c|^2 per
 * primitive (host, folded into R^2) and to |o|^2 per ray (folded into the
 * scale of the slab basis). It covers the rounding of the prefilter's own FMA
 * chains plus that of the exact sphere test it stands in front of, which is
 * bounded by ~14 eps |c - o|^2 <= 28 eps (|c|^2 + |o|^2). */
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
    uint32_t n_prims_padded; /* rounded up to RT3_PAD_PRIMS with never-hit records */
    const float4* bounds;    /* per primitive: (cx, cy, cz, -R^2), R = inflated bounding radius (R^2 includes the per-primitive slack) */
    const float4* face_n;    /* per face: (nx, ny, nz, dot3(n, p1)) */
    const float4* face_p1;   /* per face: p1.xyz */
    const float4* face_p2;
    const float4* face_p3;
    const float4* spheres;   /* per sphere: (cx, cy, cz, r) */
    const float4* prim_color;     /* per primitive: flat colour / albedo */
    const uint32_t* prim_material; /* per primitive: index into materials, or RT3_NO_HIT for Lambertian(prim_color) */
    const uint32_t* prim_entity;
    const float* prim_radius; /* per primitive (padded): sphere radius, 0 for faces */
    float axis[3];            /* unit vector along which the scene is thinnest (PCA of the primitive centres) */
    float axis_alt[3];        /* a unit vector perpendicular to axis */
    float ray_slack;          /* RT3_FILTER_SLACK / min R^2: per-ray inflation of the slab is 1 + ray_slack |o|^2 */
    const float4* materials; /* 2 float4 per material: (kind bits, albedo rgb), (fuzz, ior, -, -) */
};

struct rt3_hit { float t; uint32_t prim; };

/* Two-level conservative prefilter in a per-ray orthonormal basis (u, v) of the
 * plane perpendicular to the unit direction dn. With a = (c - o).u and
 * b = (c - o).v the distance of a centre c from the ray's line is
 * sqrt(a^2 + b^2), so
 *   level 1 (every primitive, 4 FMA):   a^2     - R^2 < 0   (the slab |a| < R)
 *   level 2 (level-1 survivors, 8 FMA): a^2+b^2 - R^2 < 0   (the line meets the bounding sphere)
 * are both necessary for any hit. u is chosen perpendicular to the scene's
 * thinnest axis, so the slab (a plane through the ray containing that axis)
 * cuts across the scene's long extent and few primitives survive level 1.
 * The per-ray slack (oracle rounding ~ eps |c-o|^2) is folded into the basis:
 * u and v are scaled by s = 1/sqrt(1 + ray_slack |o|^2) <= 1, which inflates
 * every R^2 by at least RT3_FILTER_SLACK |o|^2 at no per-test cost. */
struct rt3_ray_slab { float ux, uy, uz, nou, vx, vy, vz, nov; };

__device__ __forceinline__ rt3_ray_slab make_ray_slab(const rt3_scene_view& S, rt3_vec3 o, rt3_vec3 dn) {
    rt3_vec3 u = cross3(dn, v3(S.axis[0], S.axis[1], S.axis[2]));
    float uu = dot3(u, u);
    if (!(uu >= 0.01f)) { u = cross3(dn, v3(S.axis_alt[0], S.axis_alt[1], S.axis_alt[2])); uu = dot3(u, u); }
    const float scale = 1.0f / sqrtf(uu * (1.0f + S.ray_slack * dot3(o, o)));
    u = scale * u;
    const rt3_vec3 v = cross3(dn, u); /* |v| = |u| (dn is unit and perpendicular to u) */
    rt3_ray_slab f;
    f.ux = u.x; f.uy = u.y; f.uz = u.z; f.nou = -dot3(o, u);
    f.vx = v.x; f.vy = v.y; f.vz = v.z; f.nov = -dot3(o, v);
    return f;
}

/* Level 1: sign bit set <=> the primitive survives (|a| < R). Four FMA-pipe instructions. */
__device__ __forceinline__ float slab_test(const float4 b, const rt3_ray_slab& f) {
    const float a = __fmaf_rn(b.x, f.ux, __fmaf_rn(b.y, f.uy, __fmaf_rn(b.z, f.uz, f.nou)));
    return __fmaf_rn(a, a, b.w);
}

/* Level 2: true <=> the ray's line may meet the primitive's bounding sphere. */
__device__ __forceinline__ bool line_test(const float4 b, const rt3_ray_slab& f) {
    const float a = __fmaf_rn(b.x, f.ux, __fmaf_rn(b.y, f.uy, __fmaf_rn(b.z, f.uz, f.nou)));
    const float c = __fmaf_rn(b.x, f.vx, __fmaf_rn(b.y, f.vy, __fmaf_rn(b.z, f.vz, f.nov)));
    const float d2 = __fmaf_rn(c, c, __fmaf_rn(a, a, b.w));
    return !(d2 > 0.0f);
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
    const float4* recs;     /* shared: prefilter records of this tile */
    const float* radius;    /* shared: sphere radii of this tile, or NULL (fetch S.spheres from global) */
    uint16_t* cand;         /* shared: [RT3_RAYS][RT3_CAND_CAP][RT3_CTA_THREADS] tile-relative candidate ids */
    uint32_t first_prim;    /* global id of the tile's first primitive */
    uint32_t n;             /* primitives in this tile (multiple of RT3_PAD_PRIMS) */
};

/* Exact test of one level-2 survivor. */
template <bool PATH_MODE>
__device__ __forceinline__ void exact_candidate(const rt3_scene_view& S, const rt3_tile_view& T, uint32_t rel, rt3_vec3 o, rt3_vec3 d, rt3_hit& best) {
    const uint32_t prim = T.first_prim + rel;
    float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (prim >= S.n_faces && prim < S.n_prims) {
        /* the prefilter centre of a sphere is its own centre */
        if (T.radius) { const float4 rec = T.recs[rel]; sp = make_float4(rec.x, rec.y, rec.z, T.radius[rel]); }
        else { sp = __ldg(&S.spheres[prim - S.n_faces]); }
    }
    exact_prim<PATH_MODE>(S, prim, sp, o, d, best);
}

/* Closest hit of the thread's rays against one shared-memory tile.
 *
 * Per block of 32 primitives and per ray: one broadcast LDS.128 per primitive
 * (shared by the rays), four FMA-pipe instructions (level 1) and one funnel
 * shift that collects the sign bit into a 32-primitive survivor mask.
 * Survivors go to the ray's deferred list in ascending primitive order; the
 * list is drained once per tile (or earlier if it would overflow) from a single
 * call site: level 2, then the exact test, still in ascending order -- so the
 * strict `t < best` rule keeps the lowest index on ties exactly like the
 * reference loop (SequentialRenderer.cpp:71). */
template <bool PATH_MODE>
__device__ __forceinline__ void sweep_tile(const rt3_scene_view& S, const rt3_tile_view& T, const rt3_ray_slab (&f)[RT3_RAYS],
                                           const rt3_vec3 (&o)[RT3_RAYS], const rt3_vec3 (&d)[RT3_RAYS], const bool (&live)[RT3_RAYS],
                                           uint32_t (&n_cand)[RT3_RAYS], rt3_hit (&best)[RT3_RAYS]) {
    for (uint32_t base = 0;; base += RT3_BLOCK_PRIMS) {
        const bool last = base >= T.n;
        uint32_t cand[RT3_RAYS];
#pragma unroll
        for (int r = 0; r < RT3_RAYS; r++) { cand[r] = 0u; }
        uint32_t nb = 0;
        if (!last) {
            nb = T.n - base < RT3_BLOCK_PRIMS ? T.n - base : RT3_BLOCK_PRIMS;
            const float4* __restrict__ blk = T.recs + base;
            for (uint32_t j = 0; j < nb; j += RT3_PAD_PRIMS) {
#pragma unroll
                for (int u = 0; u < RT3_PAD_PRIMS; u++) {
                    const float4 b = blk[j + u];
#pragma unroll
                    for (int r = 0; r < RT3_RAYS; r++) {
                        cand[r] = __funnelshift_l(__float_as_uint(slab_test(b, f[r])), cand[r], 1);
                    }
                }
            }
            /* after nb shifts bit (nb - 1 - j) belongs to primitive j */
#pragma unroll
            for (int r = 0; r < RT3_RAYS; r++) { if (!live[r]) { cand[r] = 0u; } }
        }
        bool drain = last;
#pragma unroll
        for (int r = 0; r < RT3_RAYS; r++) { drain = drain || (n_cand[r] + (uint32_t) __popc(cand[r]) > RT3_CAND_CAP); }
        if (drain) {
            /* level 2 over the list, compacting the survivors in place (order preserved) */
            uint32_t nmax = 0, n_keep[RT3_RAYS];
#pragma unroll
            for (int r = 0; r < RT3_RAYS; r++) { nmax = n_cand[r] > nmax ? n_cand[r] : nmax; n_keep[r] = 0; }
            for (uint32_t e = 0; e < nmax; e++) {
#pragma unroll
                for (int r = 0; r < RT3_RAYS; r++) {
                    if (e < n_cand[r]) {
                        uint16_t* list = T.cand + (r * RT3_CAND_CAP) * RT3_CTA_THREADS + threadIdx.x;
                        const uint16_t rel = list[e * RT3_CTA_THREADS];
                        if (line_test(T.recs[rel], f[r])) { list[n_keep[r] * RT3_CTA_THREADS] = rel; n_keep[r]++; }
                    }
                }
            }
            /* exact tests of what is left, ascending primitive order */
            nmax = 0;
#pragma unroll
            for (int r = 0; r < RT3_RAYS; r++) { nmax = n_keep[r] > nmax ? n_keep[r] : nmax; }
            for (uint32_t e = 0; e < nmax; e++) {
#pragma unroll
                for (int r = 0; r < RT3_RAYS; r++) {
                    if (e < n_keep[r]) {
                        exact_candidate<PATH_MODE>(S, T, T.cand[(r * RT3_CAND_CAP + e) * RT3_CTA_THREADS + threadIdx.x], o[r], d[r], best[r]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RT3_RAYS; r++) { n_cand[r] = 0; }
        }
        if (last) { break; }
#pragma unroll
        for (int r = 0; r < RT3_RAYS; r++) {
            uint32_t c = cand[r];
            while (c) {
                const uint32_t bit = 31u - (uint32_t) __clz(c);
                c &= ~(1u << bit);
                T.cand[(r * RT3_CAND_CAP + n_cand[r]) * RT3_CTA_THREADS + threadIdx.x] = (uint16_t) (base + (nb - 1u - bit));
                n_cand[r]++;
            }
        }
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
