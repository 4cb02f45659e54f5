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

#define RT3_RAYS 2              /* rays per thread: one packed-FP32 (FFMA2) pair */
#define RT3_BLOCK_PRIMS 32      /* primitives per candidate-mask block */
#define RT3_GROUP 4             /* primitives filtered abreast (independent FMA chains in flight) */
#define RT3_PAD_PRIMS 8         /* the primitive array is padded to a multiple of this with never-hit records */
#define RT3_REC_BYTES 32        /* prefilter record: (cx,cx,cy,cy) (cz,cz,-k,-k), duplicated for the packed operands */
#define RT3_TILE_PRIMS 1024     /* primitives per streamed shared-memory tile (32 KB) */
#define RT3_RESIDENT_PRIMS 2048 /* scenes up to this size live in shared memory for the whole kernel (64 KB) */
#define RT3_CTA_THREADS 128
#define RT3_CTAS_PER_SM 5       /* register budget: 65536 / (128 * 5) = 102 per thread */
#define RT3_ITEM_CHUNK 1024u    /* path items a warp claims per global atomic */
#define RT3_CAND_CAP 32         /* per-ray deferred-candidate list entries (shared memory, 16-bit tile-relative ids) */

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
    uint32_t n_prims_padded; /* rounded up to RT3_PAD_PRIMS with never-hit records */
    const float4* bounds;    /* per primitive 2 x float4: (cx,cx,cy,cy) (cz,cz,-k,-k), k = |c|^2 - r_eff^2 - slack */
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

/* Prefilter, two rays at a time. For a sphere (c, r) and a ray (o, unit dn):
 *   h    = (c - o) . dn            = c.dn - o.dn
 *   q    = |c - o|^2 - r^2         = (|c|^2 - r^2) + |o|^2 - 2 c.o
 *   disc = h^2 - q  >= 0  <=>  the line meets the sphere.
 * Each operation is one packed fma.rn.f32x2 / add.rn.f32x2 over the ray pair
 * (SASS FFMA2 / FADD2): three for h, one add + three for -q, one for disc --
 * eight FMA-pipe instructions per primitive for two rays. */
struct rt3_pair_filter { float2 dx, dy, dz, nod, p2ox, p2oy, p2oz, noo; };

__device__ __forceinline__ rt3_pair_filter make_pair_filter(const rt3_vec3 (&o)[RT3_RAYS], const rt3_vec3 (&dn)[RT3_RAYS]) {
    rt3_pair_filter f;
    f.dx = make_float2(dn[0].x, dn[1].x); f.dy = make_float2(dn[0].y, dn[1].y); f.dz = make_float2(dn[0].z, dn[1].z);
    f.nod = make_float2(-dot3(o[0], dn[0]), -dot3(o[1], dn[1]));
    f.p2ox = make_float2(2.0f * o[0].x, 2.0f * o[1].x);
    f.p2oy = make_float2(2.0f * o[0].y, 2.0f * o[1].y);
    f.p2oz = make_float2(2.0f * o[0].z, 2.0f * o[1].z);
    float oo0 = dot3(o[0], o[0]), oo1 = dot3(o[1], o[1]);
    /* -(|o|^2 - slack |o|^2): lowering q can only add candidates */
    f.noo = make_float2(RT3_FILTER_SLACK * oo0 - oo0, RT3_FILTER_SLACK * oo1 - oo1);
    return f;
}

/* One prefilter test for the ray pair; the sign bit of each half is set <=> that ray certainly misses. */
__device__ __forceinline__ float2 filter_pair(const float4 A, const float4 B, const rt3_pair_filter& f) {
    const float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w), cz = make_float2(B.x, B.y), nk = make_float2(B.z, B.w);
    float2 h = __ffma2_rn(cx, f.dx, __ffma2_rn(cy, f.dy, __ffma2_rn(cz, f.dz, f.nod)));
    float2 nq = __ffma2_rn(cx, f.p2ox, __ffma2_rn(cy, f.p2oy, __ffma2_rn(cz, f.p2oz, __fadd2_rn(nk, f.noo))));
    return __ffma2_rn(h, h, nq);
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
    const float4* recs;     /* shared: prefilter records of this tile, 2 float4 per primitive */
    const float* radius;    /* shared: sphere radii of this tile, or NULL (fetch S.spheres from global) */
    uint16_t* cand;         /* shared: [RT3_RAYS][RT3_CAND_CAP][RT3_CTA_THREADS] tile-relative candidate ids */
    uint32_t first_prim;    /* global id of the tile's first primitive */
    uint32_t n;             /* primitives in this tile (multiple of RT3_PAD_PRIMS) */
};

template <bool PATH_MODE>
__device__ __forceinline__ void exact_candidate(const rt3_scene_view& S, const rt3_tile_view& T, uint32_t rel, rt3_vec3 o, rt3_vec3 d, rt3_hit& best) {
    const uint32_t prim = T.first_prim + rel;
    float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (prim >= S.n_faces && prim < S.n_prims) {
        if (T.radius) {
            /* the prefilter centre of a sphere is its own centre */
            const float4 A = T.recs[2 * rel], B = T.recs[2 * rel + 1];
            sp = make_float4(A.x, A.z, B.x, T.radius[rel]);
        } else {
            sp = __ldg(&S.spheres[prim - S.n_faces]);
        }
    }
    exact_prim<PATH_MODE>(S, prim, sp, o, d, best);
}

/* Closest hit of the ray pair against one shared-memory tile.
 *
 * Per block of 32 primitives: one broadcast LDS.128 pair per primitive, eight
 * packed FMA-pipe instructions and two funnel shifts that collect the sign
 * bits of disc into a per-ray miss mask. Survivors go to the ray's deferred
 * list in ascending primitive order; the exact tests run once per tile (or
 * earlier if a list would overflow), from a single call site, still in
 * ascending order -- so the strict `t < best` rule keeps the lowest index on
 * ties exactly like the reference loop (SequentialRenderer.cpp:71). */
template <bool PATH_MODE>
__device__ __forceinline__ void sweep_tile(const rt3_scene_view& S, const rt3_tile_view& T, const rt3_pair_filter& f,
                                           const rt3_vec3 (&o)[RT3_RAYS], const rt3_vec3 (&d)[RT3_RAYS], const bool (&live)[RT3_RAYS],
                                           uint32_t (&n_cand)[RT3_RAYS], rt3_hit (&best)[RT3_RAYS]) {
    for (uint32_t base = 0;; base += RT3_BLOCK_PRIMS) {
        const bool last = base >= T.n;
        uint32_t cand[RT3_RAYS] = { 0u, 0u };
        uint32_t nb = 0;
        if (!last) {
            nb = T.n - base < RT3_BLOCK_PRIMS ? T.n - base : RT3_BLOCK_PRIMS;
            const float4* __restrict__ blk = T.recs + 2 * base;
            uint32_t miss0 = 0xFFFFFFFFu, miss1 = 0xFFFFFFFFu;
            for (uint32_t j = 0; j < nb; j += RT3_PAD_PRIMS) {
                /* Stage order, RT3_GROUP primitives abreast: each packed FMA of a stage shares its ray-constant
                 * operand with the previous one (operand-reuse cache), so only the primitive pair and the
                 * accumulator pair come from the register file -- two registers per pipe cycle, evenly split
                 * over both banks -- and dependent instructions sit RT3_GROUP issues apart. */
#pragma unroll
                for (int g = 0; g < RT3_PAD_PRIMS; g += RT3_GROUP) {
                    float4 A[RT3_GROUP], B[RT3_GROUP];
                    float2 h[RT3_GROUP], nq[RT3_GROUP];
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) { A[u] = blk[2 * (j + g + u)]; B[u] = blk[2 * (j + g + u) + 1]; }
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) { nq[u] = __fadd2_rn(make_float2(B[u].z, B[u].w), f.noo); }
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) { h[u] = __ffma2_rn(make_float2(B[u].x, B[u].y), f.dz, f.nod); }
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) { nq[u] = __ffma2_rn(make_float2(B[u].x, B[u].y), f.p2oz, nq[u]); }
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) { h[u] = __ffma2_rn(make_float2(A[u].z, A[u].w), f.dy, h[u]); }
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) { nq[u] = __ffma2_rn(make_float2(A[u].z, A[u].w), f.p2oy, nq[u]); }
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) { h[u] = __ffma2_rn(make_float2(A[u].x, A[u].y), f.dx, h[u]); }
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) { nq[u] = __ffma2_rn(make_float2(A[u].x, A[u].y), f.p2ox, nq[u]); }
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) { h[u] = __ffma2_rn(h[u], h[u], nq[u]); }
#pragma unroll
                    for (int u = 0; u < RT3_GROUP; u++) {
                        miss0 = __funnelshift_l(__float_as_uint(h[u].x), miss0, 1);
                        miss1 = __funnelshift_l(__float_as_uint(h[u].y), miss1, 1);
                    }
                }
            }
            /* after nb shifts bit (nb - 1 - j) belongs to primitive j; higher bits keep their initial 1 (= miss) */
            cand[0] = live[0] ? ~miss0 : 0u;
            cand[1] = live[1] ? ~miss1 : 0u;
        }
        bool drain = last;
#pragma unroll
        for (int r = 0; r < RT3_RAYS; r++) { drain = drain || (n_cand[r] + (uint32_t) __popc(cand[r]) > RT3_CAND_CAP); }
        if (drain) {
            const uint32_t nmax = n_cand[0] > n_cand[1] ? n_cand[0] : n_cand[1];
            for (uint32_t e = 0; e < nmax; e++) {
#pragma unroll
                for (int r = 0; r < RT3_RAYS; r++) {
                    if (e < n_cand[r]) {
                        exact_candidate<PATH_MODE>(S, T, T.cand[(r * RT3_CAND_CAP + e) * RT3_CTA_THREADS + threadIdx.x], o[r], d[r], best[r]);
                    }
                }
            }
            n_cand[0] = 0; n_cand[1] = 0;
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
