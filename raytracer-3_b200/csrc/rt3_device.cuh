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

/* -DRT3_DEBUG_ASSERTS: device-side index checks at every data-dependent array access of the render kernels (compute-sanitizer is
 * closed on the B200 pool, so the bounds are checked by the code itself: profiles/sanitize_probe.py runs this build). */
#ifdef RT3_DEBUG_ASSERTS
#include <assert.h>
#define RT3_ASSERT(cond) assert(cond)
#else
#define RT3_ASSERT(cond) ((void) 0)
#endif

#define RT3_NO_HIT 0xFFFFFFFFu
#define RT3_TMIN 0.001f
#define RT3_ACC_SCALE 16777216.0f

#ifndef RT3_UNROLL_PAIRS
#define RT3_UNROLL_PAIRS 16
#endif
#define RT3_PRAGMA_STR(x) _Pragma(#x)
#define RT3_PRAGMA_UNROLL(n) RT3_PRAGMA_STR(unroll n)
#ifndef RT3_RAYS
#define RT3_RAYS 2              /* rays (path slots) per thread of the path tracer */
#endif
#ifndef RT3_REF_RAYS
#define RT3_REF_RAYS 2          /* pixels per thread of the reference-mode ray caster (the sweep is templated on it). Four rays per thread
                                 * halve the record loads per test but need 122-128 registers (4 CTAs per SM): measured slower on the
                                 * B200, 6.13 against 5.56 SMSP-cycles per warp-test at 65 536 spheres (profiles/r02b_sweep_rate_*.jsonl) */
#endif
#define RT3_WORD_PRIMS 32       /* primitives per survivor-mask word */
#define RT3_PAD_PRIMS 8         /* the primitive array is padded to a multiple of this with never-surviving records */
#ifndef RT3_CHUNK_WORDS
#define RT3_CHUNK_WORDS 16      /* mask words swept before the survivors are drained (512 primitives) */
#endif
#ifndef RT3_CONST_PRIMS
#define RT3_CONST_PRIMS 768     /* scenes up to this size are swept out of the constant bank: 9 KB of records, about what stays
                                 * resident in an SM's constant cache (measured crossover with the streamed path, profiles/crossover.py) */
#endif
#define RT3_TILE_PRIMS 1024     /* larger scenes: primitives per streamed shared-memory tile (12 KB per stage) */
#define RT3_CTA_THREADS 128
#ifndef RT3_CTAS_PER_SM
#define RT3_CTAS_PER_SM 5       /* register budget: 65536 / (128 * 5) = 102 per thread */
#endif
#define RT3_ITEM_CHUNK 256u     /* path items a warp claims per global atomic while plenty are left ... */
#define RT3_ITEM_CHUNK_MIN 32u  /* ... shrinking to this towards the end of the frame (rt3_kernels.cuh chunk_size) */

/* Relative slack of the prefilter (64 units of 2^-24), applied to |c|^2 + r^2
 * per primitive (host, folded into R^2) and to |o|^2 per ray (folded into the
 * scale of the slab direction). It covers the rounding of the prefilter's own
 * FMA chain (a few 2^-24 (|c| + |o|) on a) plus that of the exact tests it
 * stands in front of -- the exact sphere discriminant is off by up to
 * ~15 2^-24 |c - o|^2 -- see DESIGN.md section 3.1 for the bound. */
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
    uint32_t n_prims_padded; /* rounded up to RT3_PAD_PRIMS with never-surviving records */
    /* Prefilter records in the scene basis (e1, e2, e3): p_k = c . e_k for the bounding-sphere centre c,
     * w = -R^2 (inflated radius, per-primitive slack included). */
    const float4* pair_xy;   /* per primitive PAIR (2j, 2j+1): (p1a, p1b, p2a, p2b) */
    const float2* pair_w;    /* per primitive pair: (wa, wb) */
    const float4* filt3;     /* per primitive: (p1, p2, p3, w), level 2 */
    const float4* face_rec;  /* per face one 64-byte record, 4 float4: (nx, ny, nz, dot3(n, p1)), p1.xyz, p2.xyz, p3.xyz -- everything the exact
                              * test reads lies in two 32-byte sectors of one line (four separate arrays cost four lines per test) */
    const float4* spheres;   /* per sphere: (cx, cy, cz, r) */
    const float4* prim_color;     /* per primitive: flat colour / albedo */
    const uint32_t* prim_material; /* per primitive: index into materials, or RT3_NO_HIT for Lambertian(prim_color) */
    const uint32_t* prim_entity;
    float e1[3], e2[3], e3[3]; /* scene basis: e3 is the direction along which the primitive centres spread least */
    float ray_slack;           /* RT3_FILTER_SLACK / min R^2: the slab of a ray at o is widened by 1 + ray_slack |o|^2 */
    const float4* materials; /* 2 float4 per material: (kind bits, albedo rgb), (fuzz, ior, -, -) */
};

struct rt3_hit { float t; uint32_t prim; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

/* Two-level conservative prefilter.
 *
 * Every primitive has a bounding sphere (c, R). In the scene basis (e1, e2, e3)
 * let u be the unit vector of the (e1, e2) plane perpendicular to the ray
 * direction dn: u = (-g2, g1) / |(g1, g2)| with g_k = dn . e_k. u is
 * perpendicular to dn, so with a = (c - o) . u the distance of c from the ray's
 * line is at least |a|, and
 *   level 1 (every primitive):    a^2 - R^2 < 0        -- the slab |a| < R
 * is necessary for a hit. Because u has no e3 component, a needs only the two
 * in-plane coordinates of c: a = p1 u1 + p2 u2 - o.u, two FMAs, and one more for
 * a^2 - R^2. The slab is the plane through the ray that contains e3, the
 * direction along which the scene is thinnest, so it cuts across the scene's
 * long extent and few primitives survive.
 *   level 2 (faces that survive): a^2 + b^2 - R^2 < 0, b = (c - o) . (dn x u)
 * is the full line / bounding-sphere test; spheres go straight to their exact
 * test, which starts with the same discriminant.
 * The per-ray slack is folded into the scale of u: |u| = 1/sqrt(1 + ray_slack |o|^2),
 * which widens every R^2 by at least RT3_FILTER_SLACK |o|^2 at no per-test cost. */
struct rt3_ray_filter {
    float u1, u2, nou;   /* level 1 */
    float g1, g2, g3;    /* dn in the scene basis (level 2 builds v = dn x u from it) */
    float o1, o2, o3;    /* o in the scene basis */
};

__device__ __forceinline__ rt3_ray_filter make_ray_filter(const rt3_scene_view& S, rt3_vec3 o, rt3_vec3 dn) {
    const rt3_vec3 e1 = v3(S.e1[0], S.e1[1], S.e1[2]), e2 = v3(S.e2[0], S.e2[1], S.e2[2]), e3 = v3(S.e3[0], S.e3[1], S.e3[2]);
    rt3_ray_filter f;
    f.g1 = dot3(dn, e1); f.g2 = dot3(dn, e2); f.g3 = dot3(dn, e3);
    f.o1 = dot3(o, e1); f.o2 = dot3(o, e2); f.o3 = dot3(o, e3);
    const float s2 = f.g1 * f.g1 + f.g2 * f.g2;
    const float widen = 1.0f + S.ray_slack * ((f.o1 * f.o1 + f.o2 * f.o2) + f.o3 * f.o3);
    if (s2 > 1e-30f) {
        const float inv = 1.0f / sqrtf(s2 * widen);
        f.u1 = -f.g2 * inv; f.u2 = f.g1 * inv;
    } else {
        /* dn is along e3: every in-plane direction is perpendicular to it */
        f.u1 = 1.0f / sqrtf(widen); f.u2 = 0.0f;
    }
    f.nou = -(f.o1 * f.u1 + f.o2 * f.u2);
    if (!(f.nou == f.nou) || !(f.u1 == f.u1) || !(f.u2 == f.u2)) { f.u1 = 0.0f; f.u2 = 0.0f; f.nou = 0.0f; } /* degenerate ray: keep everything */
    return f;
}

/* Level 2: true <=> the ray's line may meet the bounding sphere of `rec` = (p1, p2, p3, -R^2). */
__device__ __forceinline__ bool line_test(const float4 rec, const rt3_ray_filter& f) {
    /* v = dn x u in basis coordinates; u = (u1, u2, 0) */
    const float v1 = -f.g3 * f.u2, v2 = f.g3 * f.u1, v3c = f.g1 * f.u2 - f.g2 * f.u1;
    const float nov = -((f.o1 * v1 + f.o2 * v2) + f.o3 * v3c);
    const float a = __fmaf_rn(rec.x, f.u1, __fmaf_rn(rec.y, f.u2, f.nou));
    const float b = __fmaf_rn(rec.x, v1, __fmaf_rn(rec.y, v2, __fmaf_rn(rec.z, v3c, nov)));
    const float d2 = __fmaf_rn(b, b, __fmaf_rn(a, a, rec.w));
    return !(d2 > 0.0f);
}

/* Does a hit (t, prim) replace `best`? The reference walks the primitives in ascending order and
 * rejects t >= min_t (SequentialRenderer.cpp:71), so among equal distances the lowest id wins.
 * ORDERED callers visit primitives in ascending order themselves and need only the strict compare;
 * the others (hierarchy traversal) break ties on the id explicitly. */
template <bool ORDERED>
__device__ __forceinline__ bool closer(float t, uint32_t prim, const rt3_hit& best) {
    return ORDERED ? t < best.t : (t < best.t || (t == best.t && prim < best.prim));
}

/* Exact ray-triangle test: the body of the reference's face loop,
 * SequentialRenderer.cpp:55-95, for one candidate face. tmin = 0 reproduces
 * the reference (`t < 0` rejected); the bounce loop passes 0.001. */
template <bool ORDERED>
__device__ __forceinline__ void exact_face(const rt3_scene_view& S, uint32_t i, rt3_vec3 o, rt3_vec3 d, float tmin, rt3_hit& best) {
    RT3_ASSERT(i < S.n_faces);
    const float4* rec = S.face_rec + 4 * (size_t) i;
    float4 fn = __ldg(&rec[0]);
    rt3_vec3 n = v3(fn.x, fn.y, fn.z);
    float nd = dot3(d, n);
    if (nd == 0) { return; }
    float4 a1 = __ldg(&rec[1]), a2 = __ldg(&rec[2]), a3 = __ldg(&rec[3]);
    rt3_vec3 p1 = v3(a1.x, a1.y, a1.z), p2 = v3(a2.x, a2.y, a2.z), p3 = v3(a3.x, a3.y, a3.z);
    float pd = fn.w; /* dot3(n, p1), evaluated once at upload in the same order */
    float t = (pd - dot3(n, o)) / dot3(n, d);
    if (t < tmin || !closer<ORDERED>(t, i, best)) { return; } /* a NaN t ends here too; the reference's falls through to the inside tests, which fail */
    rt3_vec3 hp = o + t * d;
    rt3_vec3 a = cross3(p2 - p1, hp - p1);
    rt3_vec3 b = cross3(p3 - p2, hp - p2);
    rt3_vec3 c = cross3(p1 - p3, hp - p3);
    if (-dot3(n, a) >= 0.0f && -dot3(n, b) >= 0.0f && -dot3(n, c) >= 0.0f) { best.prim = i; best.t = t; }
}

/* Exact ray-sphere test, reference mode: WIP hit_sphere, raytracer_v4.glsl:157-178
 * (abc form, near root, t >= 0), un-normalised direction. Written without
 * branches: the lanes of a warp walk different survivor lists, and some lane
 * hits in nearly every step. */
template <bool ORDERED>
__device__ __forceinline__ void exact_sphere_v4(uint32_t prim, float4 sp, rt3_vec3 o, rt3_vec3 d, rt3_hit& best) {
    rt3_vec3 oc = o - v3(sp.x, sp.y, sp.z);
    float a = dot3(d, d);
    float b = 2.0f * dot3(oc, d);
    float c = dot3(oc, oc) - sp.w * sp.w;
    float D = b * b - (4.0f * a) * c;
    /* a miss takes the root of 1 instead: sqrtf of zero or a negative leaves the fast path of its implementation */
    float t = (-b - sqrtf(D < 0.0f ? 1.0f : D)) / (2.0f * a);
    if (D >= 0 && t >= 0.0f && closer<ORDERED>(t, prim, best)) { best.prim = prim; best.t = t; }
}

/* Exact ray-sphere test, bounce loop: half-b form with a unit direction, near
 * then far root, accepted iff tmin <= t < best. This part
 * of the path has no implementation in the reference; its arithmetic is the
 * oracle's, which spells the dot products out as fused multiply-adds. */
template <bool ORDERED>
__device__ __forceinline__ void exact_sphere_path(uint32_t prim, float4 sp, rt3_vec3 o, rt3_vec3 d, rt3_hit& best) {
    rt3_vec3 oc = o - v3(sp.x, sp.y, sp.z);
    /* explicit fused multiply-adds, in the oracle's order (oracle/rt3_oracle.c closest_sphere_path) */
    float h = __fmaf_rn(oc.x, d.x, __fmaf_rn(oc.y, d.y, oc.z * d.z));
    float c = __fmaf_rn(oc.x, oc.x, __fmaf_rn(oc.y, oc.y, __fmaf_rn(oc.z, oc.z, -(sp.w * sp.w))));
    float disc = __fmaf_rn(h, h, -c);
    float sq = sqrtf(disc < 0.0f ? 1.0f : disc); /* a miss takes the root of 1: see exact_sphere_v4 */
    /* near root if it is in front of tmin, else the far one; the oracle's "near, then far" gives the same:
     * t1 <= t2, so a near root that is in front but not closer than `best` rules the far one out too */
    const float t1 = -h - sq, t2 = -h + sq;
    const float t = t1 >= RT3_TMIN ? t1 : t2;
    if (disc >= 0.0f && t >= RT3_TMIN && closer<ORDERED>(t, prim, best)) { best.prim = prim; best.t = t; }
}

/* The same test for callers whose lanes test the SAME sphere at the same time (the candidate lists of a chunk's primary rays,
 * rt3_kernels.cuh): there most candidates are missed by every lane of the warp, and a branch on the discriminant skips the
 * square root and the root selection for all of them -- 13 instructions instead of 37 per missed candidate. Same arithmetic,
 * same result. */
template <bool ORDERED>
__device__ __forceinline__ void exact_sphere_path_coherent(uint32_t prim, float4 sp, rt3_vec3 o, rt3_vec3 d, rt3_hit& best) {
    rt3_vec3 oc = o - v3(sp.x, sp.y, sp.z);
    float h = __fmaf_rn(oc.x, d.x, __fmaf_rn(oc.y, d.y, oc.z * d.z));
    float c = __fmaf_rn(oc.x, oc.x, __fmaf_rn(oc.y, oc.y, __fmaf_rn(oc.z, oc.z, -(sp.w * sp.w))));
    float disc = __fmaf_rn(h, h, -c);
    if (disc >= 0.0f) {
        float sq = sqrtf(disc);
        const float t1 = -h - sq, t2 = -h + sq;
        const float t = t1 >= RT3_TMIN ? t1 : t2;
        if (t >= RT3_TMIN && closer<ORDERED>(t, prim, best)) { best.prim = prim; best.t = t; }
    }
}

/* Records of scenes up to RT3_CONST_PRIMS live in the constant bank: the sweep
 * reads them through uniform registers (SASS LDCU), so the packed FMAs take the
 * primitive operand from the uniform datapath and only the ray's constants and
 * the accumulator from the vector register file. The host copies a context's
 * records here before a launch when another scene owned the bank (rt3_core.cu). */
__constant__ float4 c_pair_xy[RT3_CONST_PRIMS / 2];
__constant__ float2 c_pair_w[RT3_CONST_PRIMS / 2];

/* Per-thread survivor masks of the chunk being swept: [rays][RT3_CHUNK_WORDS][RT3_CTA_THREADS] words. */
#define RT3_MASK_BYTES_FOR(rays) ((rays) * RT3_CHUNK_WORDS * RT3_CTA_THREADS * 4)
#define RT3_MASK_BYTES RT3_MASK_BYTES_FOR(RT3_RAYS)

/* Level 1 for one primitive pair and every ray of the thread: three packed FMAs
 * (fma.rn.f32x2, SASS FFMA2: both primitives of the pair at once) and two funnel
 * shifts that push the sign bits of a^2 - R^2 into the ray's mask word. */
template <int R>
__device__ __forceinline__ void slab_pair(const float4 A, const float2 B, const float2 (&u1)[R], const float2 (&u2)[R],
                                          const float2 (&nou2)[R], uint32_t (&m)[R]) {
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float2 a = __ffma2_rn(make_float2(A.x, A.y), u1[r], __ffma2_rn(make_float2(A.z, A.w), u2[r], nou2[r]));
        const float2 d = __ffma2_rn(a, a, B);
        m[r] = __funnelshift_l(__float_as_uint(d.x), m[r], 1);
        m[r] = __funnelshift_l(__float_as_uint(d.y), m[r], 1);
    }
}

/* Level 1 over `n_pairs` primitive pairs (a multiple of RT3_PAD_PRIMS / 2; at most
 * RT3_CHUNK_WORDS words of 32 primitives) starting at pair index `first_pair` of
 * `xy` / `w` (constant bank or a shared-memory tile). Bit 31 - k of a mask word
 * belongs to its k-th primitive. Words go to shared memory; `nz` gets one bit per
 * word, set when the word is not empty, the first word of the chunk in the highest
 * of the bits used (bit n_words - 1). */
template <bool CONST_BANK, int R>
__device__ __forceinline__ void sweep_chunk(const float4* __restrict__ xy, const float2* __restrict__ w, uint32_t first_pair, uint32_t n_pairs,
                                            const rt3_ray_filter (&f)[R], uint32_t* __restrict__ masks, uint32_t (&nz)[R]) {
    constexpr uint32_t WORD_PAIRS = RT3_WORD_PRIMS / 2, PAD_PAIRS = RT3_PAD_PRIMS / 2;
    float2 nou2[R], u1[R], u2[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        nou2[r] = make_float2(f[r].nou, f[r].nou); u1[r] = make_float2(f[r].u1, f[r].u1); u2[r] = make_float2(f[r].u2, f[r].u2);
        nz[r] = 0u;
    }
    uint32_t waddr = smem_u32(masks) + threadIdx.x * 4u; /* this thread's word 0 of ray 0 */
    /* One mask word per ray out: the word goes to shared memory, its "not empty" bit into nz: nz = 2 nz + (m != 0), as the carry of
     * m + 0xFFFFFFFF added into nz + nz -- two integer adds per ray (IADD3 with carry out, IADD3.X) where the compiler's own
     * rendering of the C expression takes four. */
    auto emit = [&](const uint32_t (&m)[R]) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(waddr + (uint32_t) (r * RT3_CHUNK_WORDS * RT3_CTA_THREADS * 4)), "r"(m[r]) : "memory");
            asm("{ .reg .u32 t; add.cc.u32 t, %1, 0xFFFFFFFF; addc.u32 %0, %0, %0; }" : "+r"(nz[r]) : "r"(m[r])); /* nz = 2 nz + (m != 0) */
        }
        waddr += RT3_CTA_THREADS * 4u;
    };
    /* full words first (no partial-word test inside the loop), then the last, partial word of the scene */
    const uint32_t n_full = n_pairs / WORD_PAIRS;
    uint32_t base = first_pair;
    for (uint32_t wd = 0; wd < n_full; wd++, base += WORD_PAIRS) {
        uint32_t m[R];
#pragma unroll
        for (int r = 0; r < R; r++) { m[r] = 0u; }
RT3_PRAGMA_UNROLL(RT3_UNROLL_PAIRS)
        for (int j = 0; j < (int) WORD_PAIRS; j++) {
            slab_pair<R>(CONST_BANK ? c_pair_xy[base + j] : xy[base + j], CONST_BANK ? c_pair_w[base + j] : w[base + j], u1, u2, nou2, m);
        }
        /* (Shared-memory tiles: fetching the -R^2 of two pairs with one 16-byte load saves a quarter of the LDS instructions but was
         * measured slower, 20.2 against 19.85 ms at 65 536 spheres -- ten more registers, fewer loads in flight; profiles/r02k.) */
        emit(m);
    }
    const uint32_t left = n_pairs - n_full * WORD_PAIRS;
    if (left) {
        uint32_t m[R];
#pragma unroll
        for (int r = 0; r < R; r++) { m[r] = 0u; }
        for (uint32_t g = 0; g < left; g += PAD_PAIRS) {
#pragma unroll
            for (int j = 0; j < (int) PAD_PAIRS; j++) {
                slab_pair<R>(CONST_BANK ? c_pair_xy[base + g + j] : xy[base + g + j], CONST_BANK ? c_pair_w[base + g + j] : w[base + g + j], u1, u2, nou2, m);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) { m[r] <<= 32u - 2u * left; }
        emit(m);
    }
}

/* Exact tests of one ray's level-1 survivors of a chunk, in ascending primitive
 * order (words ascending, bits from the top), so the strict `t < best` rule
 * keeps the lowest index on ties exactly like the reference loop
 * (SequentialRenderer.cpp:71). Every lane walks its own survivor list; `masks`
 * points at this thread's first word of the ray. Padding records never survive
 * level 1, so every bit is a real primitive. */
template <bool PATH_MODE, bool SPHERES_ONLY>
__device__ __forceinline__ void drain_chunk(const rt3_scene_view& S, uint32_t first_prim, uint32_t n_words, const rt3_ray_filter& f, rt3_vec3 o,
                                            rt3_vec3 d, const uint32_t* __restrict__ masks, uint32_t nz, rt3_hit& best) {
    uint32_t m = 0u, word_prim = 0u;
    const uint32_t maddr = smem_u32(masks);
    /* (The bit walk is written with __clz. An inline-asm `bfind` would save the two 31 - x, but an asm statement here makes ptxas read
     * the records of the sweep next to it with per-thread LDC instead of uniform LDCU -- tests/test_sass_evidence.py caught it.) */
    for (;;) {
        if (m == 0u) {
            if (nz == 0u) { break; }
            const uint32_t b = 31u - (uint32_t) __clz((int) nz); /* highest bit = first non-empty word */
            nz ^= 1u << b;
            const uint32_t wd = n_words - 1u - b;
            RT3_ASSERT(b < n_words && n_words <= RT3_CHUNK_WORDS);
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m) : "r"(maddr + wd * (RT3_CTA_THREADS * 4u)));
            word_prim = first_prim + wd * RT3_WORD_PRIMS;
        }
        const uint32_t k = (uint32_t) __clz((int) m);
        m ^= 0x80000000u >> k;
        const uint32_t prim = word_prim + k;
        RT3_ASSERT(prim < S.n_prims); /* padding records never survive level 1 */
        if (SPHERES_ONLY || prim >= S.n_faces) {
            const float4 sp = __ldg(&S.spheres[SPHERES_ONLY ? prim : prim - S.n_faces]);
            if (PATH_MODE) { exact_sphere_path<true>(prim, sp, o, d, best); }
            else { exact_sphere_v4<true>(prim, sp, o, d, best); }
        } else {
            if (line_test(__ldg(&S.filt3[prim]), f)) { exact_face<true>(S, prim, o, d, PATH_MODE ? RT3_TMIN : 0.0f, best); }
        }
    }
}

/* Closest hit of the thread's rays against `n_prims` primitives (a multiple of
 * RT3_PAD_PRIMS) whose records start at pair index `first_pair`; `first_prim` is
 * the global id of the first one. */
template <bool PATH_MODE, bool CONST_BANK, bool SPHERES_ONLY, int R>
__device__ __forceinline__ void sweep_range(const rt3_scene_view& S, const float4* __restrict__ xy, const float2* __restrict__ w, uint32_t first_pair,
                                            uint32_t first_prim, uint32_t n_prims, const rt3_ray_filter (&f)[R],
                                            const rt3_vec3 (&o)[R], const rt3_vec3 (&d)[R], const bool (&live)[R],
                                            uint32_t* __restrict__ masks, rt3_hit (&best)[R]) {
    constexpr uint32_t CHUNK_PAIRS = RT3_CHUNK_WORDS * RT3_WORD_PRIMS / 2;
    const uint32_t n_pairs = n_prims / 2;
    for (uint32_t p0 = 0, w0 = 0; p0 < n_pairs; p0 += CHUNK_PAIRS, w0 += RT3_CHUNK_WORDS) {
        uint32_t nz[R];
        const uint32_t np = n_pairs - p0 < CHUNK_PAIRS ? n_pairs - p0 : CHUNK_PAIRS;
        sweep_chunk<CONST_BANK, R>(xy, w, first_pair + p0, np, f, masks, nz);
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (!live[r]) { nz[r] = 0u; }
            drain_chunk<PATH_MODE, SPHERES_ONLY>(S, first_prim + w0 * RT3_WORD_PRIMS, (np + RT3_WORD_PRIMS / 2 - 1) / (RT3_WORD_PRIMS / 2), f[r], o[r], d[r],
                                                 masks + r * RT3_CHUNK_WORDS * RT3_CTA_THREADS + threadIdx.x, nz[r], best[r]);
        }
    }
}

/* ---- mbarrier / bulk-copy (TMA) helpers for streamed tiles ---------------- */


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
