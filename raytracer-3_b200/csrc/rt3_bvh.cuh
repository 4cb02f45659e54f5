/* rt3_bvh.cuh — device-built bounding-volume hierarchy behind the same closest-hit contract.
 *
 * The brute-force sweep (rt3_device.cuh) is the reference's algorithm: every ray
 * against every primitive (SequentialRenderer.cpp:55-95). This file is the
 * "next" row of SURVEY.md section 8(f): an LBVH over the same flattened
 * primitives, built on the GPU (Morton codes -> radix sort -> Karras' radix
 * tree -> bottom-up refit), and a per-thread stack traversal whose result is
 * *identical* to the sweep's, bit for bit:
 *   - leaves run the same exact tests (exact_face / exact_sphere_*);
 *   - ties in t go to the lower primitive id, which is what the reference's
 *     ascending loop with `t >= min_t -> reject` does (SequentialRenderer.cpp:71);
 *   - boxes are widened like the sweep's bounding spheres (per primitive on the
 *     host, per ray by ray_margin |o| here), and the slab test uses the robust
 *     comparisons below, so a subtree is only skipped when none of its
 *     primitives could be reported by the exact tests at t <= best.
 *
 * Node layout (64 B, one per internal node, 4 x LDG.128): the boxes of both
 * children and their references; a reference >= 0 is an internal node, < 0 is
 * the primitive ~reference.
 */
#pragma once

#include "rt3_device.cuh"

#define RT3_BVH_STACK 128 /* >= depth of a radix tree over 63-bit keys + index bits */
#define RT3_BVH_OVERFLOW_BIT 0x80000000u /* in a thread's leaf-test counter: it had to drop a subtree */
#define RT3_BIN_MAX_SPP 8u      /* ... when a call renders at most this many samples per pixel */
#define RT3_BIN_MIN_FACES 1024u /* face trees at least this large get the binned traversal of the path tracer (rt3_kernels.cuh) */

/* Two trees, one over the faces and one over the analytic spheres, walked one after the other with the closest hit
 * carried over. They differ in how far a box must be widened for a ray starting at o: the exact sphere test's
 * discriminant is off by ~15 2^-24 |c - o|^2, so it can report a sphere that the ray's line misses by up to
 * sqrt(2^-18) |o| (what matters for tiny, distant spheres); the exact triangle test only moves the hit point by a few
 * 2^-24 (|o| + |p|). With one common margin the sphere bound (2^-9 |o|) would swell the boxes of a finely tessellated
 * mesh by a third of a triangle and make rays that start on it walk hundreds of nodes. */
struct rt3_bvh_tree {
    int32_t root;        /* reference of the root (a leaf when the tree has one primitive) */
    uint32_t n_prims;    /* 0: nothing to hit */
    float ray_margin;    /* every box is widened by ray_margin * |o| for a ray starting at o */
};
struct rt3_bvh_view {
    const float4* nodes; /* 4 float4 per internal node: (lo0.xyz, hi0.x) (hi0.yz, lo1.xy) (lo1.z, hi1.xyz) (ref0, ref1, -, -) */
    rt3_bvh_tree tree[2]; /* faces, spheres */
    uint32_t n_prims;
};

/* ---- build ---------------------------------------------------------------- */

__device__ __forceinline__ unsigned long long bvh_spread21(uint32_t v) {
    unsigned long long x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

/* 63-bit Morton code of every primitive's box centre, normalised to the box of all centres. */
__global__ void bvh_morton_kernel(uint32_t n, uint32_t first_prim, const float4* __restrict__ lo, const float4* __restrict__ hi, float3 cmin,
                                  float3 cscale, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const float4 a = lo[first_prim + i], b = hi[first_prim + i];
    float c[3] = { 0.5f * a.x + 0.5f * b.x, 0.5f * a.y + 0.5f * b.y, 0.5f * a.z + 0.5f * b.z };
    const float mn[3] = { cmin.x, cmin.y, cmin.z }, sc[3] = { cscale.x, cscale.y, cscale.z };
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float x = (c[k] - mn[k]) * sc[k] * 2097152.0f;
        if (!(x >= 0.0f)) { x = 0.0f; } /* also NaN: primitives with non-finite data */
        if (x > 2097151.0f) { x = 2097151.0f; }
        q[k] = (uint32_t) x;
    }
    keys[i] = bvh_spread21(q[0]) | (bvh_spread21(q[1]) << 1) | (bvh_spread21(q[2]) << 2);
    vals[i] = first_prim + i;
}

/* Length of the common prefix of sorted keys i and j (index bits break ties); -1 outside the array. */
__device__ __forceinline__ int bvh_delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) { return -1; }
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) { return 64 + __clz(i ^ j); }
    return __clzll((long long) (a ^ b));
}

/* Karras 2012: internal node i covers the maximal range of keys around i that share a longer prefix
 * than i shares with its other neighbour; its children split that range at the highest differing bit.
 * child[2i], child[2i+1]: >= 0 internal node, < 0 leaf ~position. */
__global__ void bvh_tree_kernel(int n, const unsigned long long* __restrict__ keys, int* __restrict__ child, int* __restrict__ node_parent,
                                int* __restrict__ leaf_parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) { return; }
    const int d = bvh_delta(keys, n, i, i + 1) - bvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = bvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (bvh_delta(keys, n, i, i + lmax * d) > dmin) { lmax *= 2; }
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) {
        if (bvh_delta(keys, n, i, i + (l + t) * d) > dmin) { l += t; }
    }
    const int j = i + l * d;
    const int dnode = bvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (bvh_delta(keys, n, i, i + (s + t) * d) > dnode) { s += t; }
        if (t == 1) { break; }
    }
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int left = lo == gamma ? ~gamma : gamma;
    const int right = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    child[2 * i] = left; child[2 * i + 1] = right;
    if (left < 0) { leaf_parent[~left] = i; } else { node_parent[left] = i; }
    if (right < 0) { leaf_parent[~right] = i; } else { node_parent[right] = i; }
    if (i == 0) { node_parent[0] = -1; }
}

/* Bottom-up refit, one thread per leaf: the second thread to arrive at a node owns it, unions its
 * children's boxes, writes the node record and climbs on. */
__global__ void bvh_refit_kernel(int n, const uint32_t* __restrict__ vals, const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi,
                                 const int* __restrict__ child, const int* __restrict__ node_parent, const int* __restrict__ leaf_parent,
                                 float4* box_lo, float4* box_hi, unsigned int* __restrict__ arrived, float4* __restrict__ nodes, int node_offset) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) { return; }
    int p = leaf_parent[k];
    while (p >= 0) {
        __threadfence();
        if (atomicAdd(&arrived[p], 1u) == 0u) { return; }
        __threadfence();
        float4 lo[2], hi[2];
        int ref[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int ch = child[2 * p + c];
            if (ch < 0) {
                const uint32_t prim = vals[~ch];
                lo[c] = prim_lo[prim]; hi[c] = prim_hi[prim];
                ref[c] = ~(int) prim;
            } else {
                lo[c] = __ldcg(&box_lo[ch]); hi[c] = __ldcg(&box_hi[ch]);
                ref[c] = ch + node_offset; /* node records of all trees share one array */
            }
        }
        box_lo[p] = make_float4(fminf(lo[0].x, lo[1].x), fminf(lo[0].y, lo[1].y), fminf(lo[0].z, lo[1].z), 0.f);
        box_hi[p] = make_float4(fmaxf(hi[0].x, hi[1].x), fmaxf(hi[0].y, hi[1].y), fmaxf(hi[0].z, hi[1].z), 0.f);
        float4* out = nodes + 4 * (size_t) (p + node_offset);
        out[0] = make_float4(lo[0].x, lo[0].y, lo[0].z, hi[0].x);
        out[1] = make_float4(hi[0].y, hi[0].z, lo[1].x, lo[1].y);
        out[2] = make_float4(lo[1].z, hi[1].x, hi[1].y, hi[1].z);
        out[3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.f, 0.f);
        p = node_parent[p];
    }
}

/* ---- traversal ------------------------------------------------------------- */

/* Robust slab test (Ize, "Robust BVH ray traversal", 2013): every computed entry / exit distance is
 * within a few ulps of its real value, so comparing with the far side scaled by 1 + 2^-20 can only
 * admit more. A box is entered iff [t_in, t_out] meets [0, limit]. */
#define RT3_BVH_ROBUST 1.00000095367431640625f

#ifndef RT3_BVH_FMA_SLAB
#define RT3_BVH_FMA_SLAB 1
#endif
#define RT3_BVH_NONE 0x7fffffff /* neither a node (>= 0) nor a leaf (~primitive < 0) */

struct rt3_bvh_ray {
    rt3_vec3 o_lo, o_hi, inv; /* o + grow, o - grow: the per-ray widening of every box, moved onto the origin */
};

/* 1 / d per axis for the slab test. An axis whose |d| is below 2^-60 (or whose o / d leaves the float
 * range) gets NaN: both of its slab distances are then NaN, fminf / fmaxf drop them and the axis is
 * ignored, which can only admit more boxes. */
__device__ __forceinline__ float bvh_axis_inv(float d, float o_a, float o_b) {
    const float inv = 1.0f / d;
    const float big = fmaxf(fabsf(o_a * inv), fabsf(o_b * inv));
    return (fabsf(d) >= 8.673617379884035e-19f && big < 3.0e38f) ? inv : __int_as_float(0x7fc00000);
}

__device__ __forceinline__ bool bvh_slab(const rt3_bvh_ray& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float limit,
                                         float& t_in) {
#if RT3_BVH_FMA_SLAB
    /* o_lo / o_hi hold -(o + grow) / d and -(o - grow) / d: one FMA per plane. The rounding of that product is the
     * distance to a plane moved by at most 2^-24 |o|, a small part of grow (>= 2^-19 |o|); the rest is relative. */
    const float x0 = __fmaf_rn(lox, r.inv.x, r.o_lo.x), x1 = __fmaf_rn(hix, r.inv.x, r.o_hi.x);
    const float y0 = __fmaf_rn(loy, r.inv.y, r.o_lo.y), y1 = __fmaf_rn(hiy, r.inv.y, r.o_hi.y);
    const float z0 = __fmaf_rn(loz, r.inv.z, r.o_lo.z), z1 = __fmaf_rn(hiz, r.inv.z, r.o_hi.z);
#else
    /* (lo - grow) - o == lo - (o + grow) up to one rounding of o + grow, which is 2^-24 |o| against grow >= 2^-19 |o| */
    const float x0 = (lox - r.o_lo.x) * r.inv.x, x1 = (hix - r.o_hi.x) * r.inv.x;
    const float y0 = (loy - r.o_lo.y) * r.inv.y, y1 = (hiy - r.o_hi.y) * r.inv.y;
    const float z0 = (loz - r.o_lo.z) * r.inv.z, z1 = (hiz - r.o_hi.z) * r.inv.z;
#endif
    /* fminf / fmaxf return the other operand for a NaN (0 * inf: origin on a slab plane of an axis the ray is parallel to) */
    const float tin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
    const float tout = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    t_in = tin;
    return tin <= tout * RT3_BVH_ROBUST && tout >= 0.0f && tin <= limit;
}

/* Closest hit of ray (o, d) through the hierarchy; same result as the sweep. `visits` counts node records
 * read, `tests` exact tests run. */
template <bool PATH_MODE>
__device__ __forceinline__ void bvh_closest_hit(const rt3_scene_view& S, const rt3_bvh_view& B, rt3_vec3 o, rt3_vec3 d, rt3_hit& best,
                                                uint32_t& visits, uint32_t& tests) {
    best.t = __int_as_float(0x7f800000);
    best.prim = RT3_NO_HIT;
    uint2 stack[RT3_BVH_STACK]; /* (subtree reference, its entry distance): one 8-byte local access per push / pop */
    const float o_len = sqrtf(dot3(o, o));
    rt3_bvh_ray r;
#if RT3_BVH_FMA_SLAB
    {
        const float g = B.tree[1].ray_margin * o_len; /* the larger of the two widenings */
        r.inv = v3(bvh_axis_inv(d.x, o.x + g, o.x - g), bvh_axis_inv(d.y, o.y + g, o.y - g), bvh_axis_inv(d.z, o.z + g, o.z - g));
    }
#else
    r.inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
#endif
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
        const rt3_bvh_tree T = B.tree[which];
        if (T.n_prims == 0u) { continue; }
        const float grow = T.ray_margin * o_len;
#if RT3_BVH_FMA_SLAB
        r.o_lo = v3(-(o.x + grow) * r.inv.x, -(o.y + grow) * r.inv.y, -(o.z + grow) * r.inv.z);
        r.o_hi = v3(-(o.x - grow) * r.inv.x, -(o.y - grow) * r.inv.y, -(o.z - grow) * r.inv.z);
#else
        r.o_lo = v3(o.x + grow, o.y + grow, o.z + grow);
        r.o_hi = v3(o.x - grow, o.y - grow, o.z - grow);
#endif
        int sp = 0;
        int32_t ref = T.root;
        /* next stacked subtree still worth entering, or RT3_BVH_NONE */
        auto pop = [&]() -> int32_t {
            while (sp > 0) {
                const uint2 e = stack[--sp];
                if (__uint_as_float(e.y) <= best.t * RT3_BVH_ROBUST) { return (int32_t) e.x; }
            }
            return RT3_BVH_NONE;
        };
        /* one node record: both child boxes against the ray; returns the subtree to enter next, stacks the farther one */
        auto visit = [&](int32_t node) -> int32_t {
            visits++;
            RT3_ASSERT(node >= 0 && (uint32_t) node + 2u <= B.tree[0].n_prims + B.tree[1].n_prims); /* n - 1 nodes per tree */
            const float4 n0 = __ldg(&B.nodes[4 * node + 0]), n1 = __ldg(&B.nodes[4 * node + 1]), n2 = __ldg(&B.nodes[4 * node + 2]),
                         n3 = __ldg(&B.nodes[4 * node + 3]);
            const float limit = best.t * RT3_BVH_ROBUST;
            float t0, t1;
            const bool h0 = bvh_slab(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, limit, t0);
            const bool h1 = bvh_slab(r, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, limit, t1);
            const int32_t c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool first0 = t0 <= t1;
                /* a full stack cannot happen for a radix tree over 63-bit keys + 31 index bits (depth <= 94 of 128); if it ever does, the top
                 * bit of the thread's test counter remembers it (no branch, no atomic in this loop: an atomic here cost 40 % on C3),
                 * count_accel reports it and the render fails (rt3_core.cu) instead of dropping the subtree silently */
                if (sp < RT3_BVH_STACK) { stack[sp++] = make_uint2((uint32_t) (first0 ? c1 : c0), __float_as_uint(first0 ? t1 : t0)); }
                else { tests |= RT3_BVH_OVERFLOW_BIT; }
                return first0 ? c0 : c1;
            }
            if (h0) { return c0; }
            if (h1) { return c1; }
            return pop();
        };
        auto test = [&](int32_t leaf) {
            const uint32_t prim = (uint32_t) ~leaf;
            RT3_ASSERT(prim < S.n_prims);
            tests++;
            if (prim < S.n_faces) {
                exact_face<false>(S, prim, o, d, PATH_MODE ? RT3_TMIN : 0.0f, best);
            } else {
                const float4 sp4 = __ldg(&S.spheres[prim - S.n_faces]);
                if (PATH_MODE) { exact_sphere_path<false>(prim, sp4, o, d, best); }
                else { exact_sphere_v4<false>(prim, sp4, o, d, best); }
            }
        };
        while (ref != RT3_BVH_NONE) {
            if (ref < 0) { test(ref); ref = pop(); }
            else { ref = visit(ref); }
        }
    }
}
