/* rt3_upload.cuh — the derived scene arrays, built on the device.
 *
 * The reference's GPU backend flattens the entities and derives nothing (it is brute force,
 * VulkanRenderer.cpp:266-399); this core also needs, per primitive, a bounding sphere (prefilter records in
 * the scene basis, rt3_device.cuh), a box (hierarchy leaves, rt3_bvh.cuh) and, per scene, the basis itself
 * (PCA of the centres), the smallest bounding radius and the box of the box centres. Round 1 computed all
 * of that in host loops inside rt3_scene_upload; here it is a handful of kernels over the flattened arrays
 * where they lie in HBM, so that a scene tessellated on the device (rt3_scene.cuh) never visits the host,
 * and a host scene costs one copy per input array and nothing per primitive on the CPU.
 *
 *   faces_kernel / spheres_kernel : one thread per primitive -> exact-test arrays (face records with the plane offset in
 *                                   the reference's operation order, p1..p3, spheres), colour / material / entity,
 *                                   bounding sphere (double), box, validation (first error in input order wins)
 *   (CUB radix sort of the R^2 keys)  -> median and minimum of the bounding radii
 *   moments_kernel + basis_kernel : fixed-order (deterministic) reduction of the first and second moments of the
 *                                   centres -> covariance -> Jacobi -> scene basis, ray slack
 *   records_kernel                : one thread per primitive PAIR -> level-1 / level-2 prefilter records, box of
 *                                   the box centres (Morton grid of the hierarchy build)
 *
 * Everything that decides a pixel stays in the exact tests; these records only have to be conservative
 * (DESIGN.md 3.1), and they follow the host formulas of round 1 operation for operation in double precision.
 */
#pragma once

#include "rt3_device.cuh"

/* What the host needs back from a scene build: one small device -> host copy. */
struct rt3_build_info {
    unsigned long long error_key; /* ~0: no error; else (stage << 62) | (index << 2) | kind, the smallest = first in input order */
    uint32_t error_detail[4];
    unsigned long long n_valid;   /* primitives with a finite bounding sphere */
    float basis[3][3];
    float ray_slack;
    int cmin_bits[3], cmax_bits[3]; /* box of the finite box centres, as order-preserving integers (float_to_ordered) */
    double r2_median, r2_min;
};

#define RT3_BUILD_ERR_VERTEX 0u
#define RT3_BUILD_ERR_MATERIAL 1u
#define RT3_BUILD_ERR_KIND 2u

struct rt3_bound { double cx, cy, cz, R2; }; /* R2 < 0: can never be hit (non-finite data) */

__device__ __forceinline__ int float_to_ordered(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__host__ __device__ inline float ordered_to_float(int i) {
    const int b = i >= 0 ? i : i ^ 0x7fffffff;
#if defined(__CUDA_ARCH__)
    return __int_as_float(b);
#else
    float f; memcpy(&f, &b, sizeof f); return f;
#endif
}

__device__ __forceinline__ float round_down_f(double v) { return __double2float_rd(v); }
__device__ __forceinline__ float round_up_f(double v) { return __double2float_ru(v); }

__device__ __forceinline__ void report_error(rt3_build_info* info, unsigned stage, uint32_t index, uint32_t kind, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const unsigned long long key = ((unsigned long long) stage << 62) | ((unsigned long long) index << 2) | kind;
    const unsigned long long old = atomicMin(&info->error_key, key);
    if (key < old) { info->error_detail[0] = a; info->error_detail[1] = b; info->error_detail[2] = c; info->error_detail[3] = d; } /* a later, smaller key overwrites: benign for a message */
}

/* Bounding sphere as the prefilter sees it: R^2 = (r^2 + RT3_FILTER_SLACK (|c|^2 + r^2)) (1 + 1e-6) + 1e-30. */
__device__ __forceinline__ rt3_bound make_bound_dev(double cx, double cy, double cz, double r) {
    rt3_bound b = { 0.0, 0.0, 0.0, -1.0 };
    if (!(isfinite(cx) && isfinite(cy) && isfinite(cz) && isfinite(r))) { return b; }
    const double cc = cx * cx + cy * cy + cz * cz;
    const double R2 = (r * r + (double) RT3_FILTER_SLACK * (cc + r * r)) * (1.0 + 1e-6) + 1e-30;
    if (!isfinite(R2) || R2 > 1e37) { return b; }
    b.cx = cx; b.cy = cy; b.cz = cz; b.R2 = R2;
    return b;
}

__device__ __forceinline__ double dist3(const double* p, const double* q) {
    return sqrt((p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2]));
}

/* Smallest sphere through / around a triangle (double precision). */
__device__ __forceinline__ void triangle_bound_dev(const double a[3], const double b[3], const double c[3], double centre[3], double* radius) {
    const double* v[3] = { a, b, c };
    int e0 = 0;
    double best = -1.0;
    for (int e = 0; e < 3; e++) { const double l = dist3(v[e], v[(e + 1) % 3]); if (l > best) { best = l; e0 = e; } }
    const double *p = v[e0], *q = v[(e0 + 1) % 3], *o = v[(e0 + 2) % 3];
    const double mid[3] = { 0.5 * (p[0] + q[0]), 0.5 * (p[1] + q[1]), 0.5 * (p[2] + q[2]) };
    /* longest edge first: if the opposite vertex lies inside its diameter sphere, that sphere is minimal */
    if (dist3(mid, o) <= 0.5 * best) { centre[0] = mid[0]; centre[1] = mid[1]; centre[2] = mid[2]; *radius = 0.5 * best; return; }
    const double ab[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, ac[3] = { c[0] - a[0], c[1] - a[1], c[2] - a[2] };
    const double n[3] = { ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0] };
    const double n2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    const double ab2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2], ac2 = ac[0] * ac[0] + ac[1] * ac[1] + ac[2] * ac[2];
    if (n2 > 0.0 && isfinite(n2)) {
        /* acute: circumsphere, centre = a + (|ac|^2 (n x ab) + |ab|^2 (ac x n)) / (2 |n|^2) */
        const double nxab[3] = { n[1] * ab[2] - n[2] * ab[1], n[2] * ab[0] - n[0] * ab[2], n[0] * ab[1] - n[1] * ab[0] };
        const double acxn[3] = { ac[1] * n[2] - ac[2] * n[1], ac[2] * n[0] - ac[0] * n[2], ac[0] * n[1] - ac[1] * n[0] };
        for (int i = 0; i < 3; i++) { centre[i] = a[i] + (ac2 * nxab[i] + ab2 * acxn[i]) / (2.0 * n2); }
    } else {
        for (int i = 0; i < 3; i++) { centre[i] = (a[i] + b[i] + c[i]) / 3.0; }
    }
    double r = 0.0;
    for (int i = 0; i < 3; i++) { const double l = dist3(centre, v[i]); if (l > r) { r = l; } }
    *radius = r;
}

struct rt3_build_out {
    float4 *face_rec, *spheres, *prim_color, *prim_lo, *prim_hi;
    uint32_t *prim_material, *prim_entity;
    rt3_bound* bounds;
    float* r2_keys; /* (float) R^2, +inf for primitives that can never be hit: sorted for the median */
};

__device__ __forceinline__ void store_bound(const rt3_build_out& o, uint32_t prim, const rt3_bound& b, rt3_build_info* info) {
    o.bounds[prim] = b;
    o.r2_keys[prim] = b.R2 >= 0 ? (float) b.R2 : __int_as_float(0x7f800000);
    if (b.R2 >= 0) { atomicAdd(&info->n_valid, 1ull); }
}

/* One thread per face of the flattened scene (GFace records, reference renderer/Vertex.hpp:39-51). */
__global__ void build_faces_kernel(uint32_t n_faces, uint32_t n_vertices, const uint4* __restrict__ faces, const float4* __restrict__ vertices,
                                   const uint32_t* __restrict__ face_material, const uint32_t* __restrict__ face_entity, uint32_t n_materials,
                                   rt3_build_out o, rt3_build_info* info) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_faces) { return; }
    const uint4 idx = faces[3 * (size_t) i], nb = faces[3 * (size_t) i + 1], cb = faces[3 * (size_t) i + 2];
    const float nx = __uint_as_float(nb.x), ny = __uint_as_float(nb.y), nz = __uint_as_float(nb.z);
    const float fnan = __int_as_float(0x7fc00000);
    if (idx.x >= n_vertices || idx.y >= n_vertices || idx.z >= n_vertices) {
        report_error(info, 0u, i, RT3_BUILD_ERR_VERTEX, idx.x, idx.y, idx.z, n_vertices);
        return;
    }
    const float4 a = vertices[idx.x], b = vertices[idx.y], c = vertices[idx.z];
    /* plane offset dot3(n, p1) in the reference's order (SequentialRenderer.cpp:32-33,67); this unit is built without contraction */
    const float pd = (nx * a.x + ny * a.y) + nz * a.z;
    float4* rec = o.face_rec + 4 * (size_t) i;
    rec[0] = make_float4(nx, ny, nz, pd);
    rec[1] = make_float4(a.x, a.y, a.z, 0.f);
    rec[2] = make_float4(b.x, b.y, b.z, 0.f);
    rec[3] = make_float4(c.x, c.y, c.z, 0.f);
    const double da[3] = { a.x, a.y, a.z }, db[3] = { b.x, b.y, b.z }, dc[3] = { c.x, c.y, c.z };
    double centre[3], radius;
    triangle_bound_dev(da, db, dc, centre, &radius);
    const double r_geom = radius;
    /* faces: the exact test accepts hit points up to a few ulps of the coordinates outside the triangle; widen the
     * bounding sphere by 2^-10 relative and 2^-16 (|c| + r) absolute on top of the common slack */
    radius = radius * (1.0 + 1.0 / 1024.0) + (sqrt(centre[0] * centre[0] + centre[1] * centre[1] + centre[2] * centre[2]) + radius) / 65536.0;
    const rt3_bound bd = make_bound_dev(centre[0], centre[1], centre[2], radius);
    store_bound(o, i, bd, info);
    if (bd.R2 >= 0) {
        /* around the triangle's own box: the geometric widening of its bounding sphere, without the slack term of the sphere discriminant */
        const double m = radius - r_geom;
        o.prim_lo[i] = make_float4(round_down_f(fmin(fmin(da[0], db[0]), dc[0]) - m), round_down_f(fmin(fmin(da[1], db[1]), dc[1]) - m),
                                   round_down_f(fmin(fmin(da[2], db[2]), dc[2]) - m), 0.f);
        o.prim_hi[i] = make_float4(round_up_f(fmax(fmax(da[0], db[0]), dc[0]) + m), round_up_f(fmax(fmax(da[1], db[1]), dc[1]) + m),
                                   round_up_f(fmax(fmax(da[2], db[2]), dc[2]) + m), 0.f);
    } else {
        /* never entered (every comparison of the slab test fails), ignored by fminf / fmaxf unions */
        o.prim_lo[i] = make_float4(fnan, fnan, fnan, 0.f); o.prim_hi[i] = make_float4(fnan, fnan, fnan, 0.f);
    }
    o.prim_color[i] = make_float4(__uint_as_float(cb.x), __uint_as_float(cb.y), __uint_as_float(cb.z), 0.f);
    uint32_t mat = RT3_NO_HIT;
    if (face_material) {
        mat = face_material[i];
        if (mat >= n_materials) { report_error(info, 0u, i, RT3_BUILD_ERR_MATERIAL, mat, n_materials, 0u, 0u); mat = RT3_NO_HIT; }
    }
    o.prim_material[i] = mat;
    o.prim_entity[i] = face_entity ? face_entity[i] : 0u;
}

__global__ void build_spheres_kernel(uint32_t n_faces, uint32_t n_spheres, const float4* __restrict__ spheres, const float* __restrict__ sphere_color,
                                     const uint32_t* __restrict__ sphere_material, const uint32_t* __restrict__ sphere_entity, uint32_t n_materials,
                                     rt3_build_out o, rt3_build_info* info) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_spheres) { return; }
    const uint32_t prim = n_faces + i;
    const float4 sp = spheres[i];
    const float fnan = __int_as_float(0x7fc00000);
    o.spheres[i] = sp;
    const rt3_bound bd = make_bound_dev(sp.x, sp.y, sp.z, fabs((double) sp.w));
    store_bound(o, prim, bd, info);
    if (bd.R2 >= 0) {
        const double R = sqrt(bd.R2);
        o.prim_lo[prim] = make_float4(round_down_f(bd.cx - R), round_down_f(bd.cy - R), round_down_f(bd.cz - R), 0.f);
        o.prim_hi[prim] = make_float4(round_up_f(bd.cx + R), round_up_f(bd.cy + R), round_up_f(bd.cz + R), 0.f);
    } else {
        o.prim_lo[prim] = make_float4(fnan, fnan, fnan, 0.f); o.prim_hi[prim] = make_float4(fnan, fnan, fnan, 0.f);
    }
    o.prim_color[prim] = sphere_color ? make_float4(sphere_color[3 * (size_t) i], sphere_color[3 * (size_t) i + 1], sphere_color[3 * (size_t) i + 2], 0.f)
                                      : make_float4(1.f, 1.f, 1.f, 0.f);
    uint32_t mat = RT3_NO_HIT;
    if (sphere_material) {
        mat = sphere_material[i];
        if (mat >= n_materials) { report_error(info, 1u, i, RT3_BUILD_ERR_MATERIAL, mat, n_materials, 0u, 0u); mat = RT3_NO_HIT; }
    }
    o.prim_material[prim] = mat;
    o.prim_entity[prim] = sphere_entity ? sphere_entity[i] : 0u;
}

/* rt3_material (32 B) -> 2 float4 per material: (kind bits, albedo rgb), (fuzz, ior, -, -). */
__global__ void build_materials_kernel(uint32_t n, const uint4* __restrict__ in, float4* __restrict__ out, rt3_build_info* info) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    const uint4 a = in[2 * (size_t) i], b = in[2 * (size_t) i + 1]; /* kind, albedo[3] | fuzz, ior, pad, pad */
    if (a.x > RT3_MAT_DIELECTRIC) { report_error(info, 2u, i, RT3_BUILD_ERR_KIND, a.x, 0u, 0u, 0u); }
    out[2 * (size_t) i] = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
    out[2 * (size_t) i + 1] = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), 0.f, 0.f);
}

/* Median and minimum of the bounding radii from the sorted keys (one thread). Primitives far larger than the
 * typical one (a ground sphere of radius 1000) survive every slab anyway and would only skew the statistics:
 * the basis is taken over those with R^2 <= 100 x median. The smallest R^2 is floored at 1/400 of the median
 * so that one sliver cannot widen every slab; records below the floor are raised to it (admits more, never less). */
__global__ void build_radius_stats_kernel(const float* __restrict__ sorted_keys, rt3_build_info* info) {
    const unsigned long long n = info->n_valid;
    double median = 1.0, lowest = 1.0;
    if (n > 0) { median = (double) sorted_keys[n / 2]; lowest = (double) sorted_keys[0]; }
    double r2min = n > 0 ? fmax(lowest, median / 400.0) : 1.0;
    if (!(r2min > 0.0)) { r2min = 1e-30; }
    info->r2_median = median;
    info->r2_min = r2min;
    info->ray_slack = (float) ((double) RT3_FILTER_SLACK / r2min * (1.0 + 1e-6));
}

#define RT3_MOMENT_BLOCKS 256
#define RT3_MOMENT_THREADS 256
#define RT3_MOMENTS 10 /* n, sum x y z, sum xx xy xz yy yz zz */

/* First and second moments of the bounding-sphere centres (those with R^2 <= cap), reduced in a fixed order:
 * thread-strided partial sums, a shared-memory tree per block, one row of partials per block. */
__global__ void __launch_bounds__(RT3_MOMENT_THREADS) build_moments_kernel(uint32_t n_prims, const rt3_bound* __restrict__ bounds, const rt3_build_info* __restrict__ info,
                                                                          double* __restrict__ partials) {
    __shared__ double sh[RT3_MOMENT_THREADS];
    const double cap = 100.0 * info->r2_median;
    double m[RT3_MOMENTS];
#pragma unroll
    for (int k = 0; k < RT3_MOMENTS; k++) { m[k] = 0.0; }
    for (uint32_t i = blockIdx.x * RT3_MOMENT_THREADS + threadIdx.x; i < n_prims; i += RT3_MOMENT_BLOCKS * RT3_MOMENT_THREADS) {
        const rt3_bound b = bounds[i];
        if (!(b.R2 >= 0 && b.R2 <= cap)) { continue; }
        m[0] += 1.0; m[1] += b.cx; m[2] += b.cy; m[3] += b.cz;
        m[4] += b.cx * b.cx; m[5] += b.cx * b.cy; m[6] += b.cx * b.cz; m[7] += b.cy * b.cy; m[8] += b.cy * b.cz; m[9] += b.cz * b.cz;
    }
    for (int k = 0; k < RT3_MOMENTS; k++) {
        sh[threadIdx.x] = m[k];
        __syncthreads();
        for (int off = RT3_MOMENT_THREADS / 2; off > 0; off >>= 1) {
            if ((int) threadIdx.x < off) { sh[threadIdx.x] += sh[threadIdx.x + off]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) { partials[blockIdx.x * RT3_MOMENTS + k] = sh[0]; }
        __syncthreads();
    }
}

/* Jacobi eigen-decomposition of a symmetric 3x3 matrix; eigenvectors in the columns of v. */
__device__ inline void jacobi3_dev(double a[3][3], double v[3][3], double w[3]) {
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) { v[i][j] = i == j ? 1.0 : 0.0; } }
    for (int sweep = 0; sweep < 32; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-300) { break; }
        for (int p = 0; p < 2; p++) {
            for (int q = p + 1; q < 3; q++) {
                if (fabs(a[p][q]) < 1e-300) { continue; }
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; k++) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; k++) { const double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq; }
            }
        }
    }
    for (int i = 0; i < 3; i++) { w[i] = a[i][i]; }
}

/* Scene basis for the slab prefilter (one thread): e3 = the direction along which the primitive centres spread
 * least (PCA), e1 = the direction of largest spread, e2 = e3 x e1; rounded to float. Default: e3 = y. */
__global__ void build_basis_kernel(const double* __restrict__ partials, rt3_build_info* info) {
    const float dflt[3][3] = { { 1.f, 0.f, 0.f }, { 0.f, 0.f, -1.f }, { 0.f, 1.f, 0.f } };
    for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) { info->basis[r][k] = dflt[r][k]; } }
    for (int k = 0; k < 3; k++) { info->cmin_bits[k] = 0x7fffffff; info->cmax_bits[k] = (int) 0x80000000; }
    double m[RT3_MOMENTS];
    for (int k = 0; k < RT3_MOMENTS; k++) { m[k] = 0.0; for (int b = 0; b < RT3_MOMENT_BLOCKS; b++) { m[k] += partials[b * RT3_MOMENTS + k]; } }
    if (info->n_valid < 2 || m[0] < 2.0) { return; }
    const double n = m[0], mean[3] = { m[1] / n, m[2] / n, m[3] / n };
    double cov[3][3];
    cov[0][0] = m[4] - n * mean[0] * mean[0]; cov[0][1] = cov[1][0] = m[5] - n * mean[0] * mean[1]; cov[0][2] = cov[2][0] = m[6] - n * mean[0] * mean[2];
    cov[1][1] = m[7] - n * mean[1] * mean[1]; cov[1][2] = cov[2][1] = m[8] - n * mean[1] * mean[2]; cov[2][2] = m[9] - n * mean[2] * mean[2];
    double vec[3][3], val[3];
    jacobi3_dev(cov, vec, val);
    if (!(isfinite(val[0]) && isfinite(val[1]) && isfinite(val[2]))) { return; }
    int lo = 0, hi = 0;
    for (int k = 1; k < 3; k++) { if (val[k] < val[lo]) { lo = k; } if (val[k] > val[hi]) { hi = k; } }
    if (lo == hi) { return; }
    double a3[3], a1[3], la = 0, lb = 0;
    for (int k = 0; k < 3; k++) { a3[k] = vec[k][lo]; a1[k] = vec[k][hi]; la += a3[k] * a3[k]; lb += a1[k] * a1[k]; }
    if (!(la > 0.5 && lb > 0.5)) { return; }
    for (int k = 0; k < 3; k++) { a3[k] /= sqrt(la); a1[k] /= sqrt(lb); }
    /* re-orthogonalise e1 against e3 (eigenvectors of a symmetric matrix already are, up to rounding) */
    double dp = a1[0] * a3[0] + a1[1] * a3[1] + a1[2] * a3[2], l1 = 0;
    for (int k = 0; k < 3; k++) { a1[k] -= dp * a3[k]; l1 += a1[k] * a1[k]; }
    if (!(l1 > 0.25)) { return; }
    for (int k = 0; k < 3; k++) { a1[k] /= sqrt(l1); }
    const double a2[3] = { a3[1] * a1[2] - a3[2] * a1[1], a3[2] * a1[0] - a3[0] * a1[2], a3[0] * a1[1] - a3[1] * a1[0] };
    for (int k = 0; k < 3; k++) { info->basis[0][k] = (float) a1[k]; info->basis[1][k] = (float) a2[k]; info->basis[2][k] = (float) a3[k]; }
}

/* Prefilter record (p1, p2, p3, -R^2) of one primitive in the scene basis; a primitive that can never be hit (and
 * the padding behind the last one) gets w = +inf: a^2 + inf > 0 never survives. */
__device__ __forceinline__ float4 filter_record(uint32_t prim, uint32_t n_prims, const rt3_bound* __restrict__ bounds, const rt3_build_info* __restrict__ info) {
    const float inf = __int_as_float(0x7f800000);
    const float4 never = make_float4(0.f, 0.f, 0.f, inf);
    if (prim >= n_prims) { return never; }
    const rt3_bound b = bounds[prim];
    if (b.R2 < 0) { return never; }
    float p[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { p[k] = (float) (b.cx * (double) info->basis[k][0] + b.cy * (double) info->basis[k][1] + b.cz * (double) info->basis[k][2]); }
    const float w = round_down_f(-fmax(b.R2, info->r2_min));
    if (!isfinite(w) || !isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) { return never; }
    return make_float4(p[0], p[1], p[2], w);
}

/* One thread per primitive pair: level-2 records (filt3), the packed level-1 pair records, and the box of the finite
 * box centres (order-independent integer min / max: deterministic). */
__global__ void build_records_kernel(uint32_t n_prims, uint32_t n_pairs, const rt3_bound* __restrict__ bounds, const float4* __restrict__ prim_lo,
                                     const float4* __restrict__ prim_hi, float4* __restrict__ filt3, float4* __restrict__ pair_xy, float2* __restrict__ pair_w,
                                     rt3_build_info* info) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pairs) { return; }
    const float4 a = filter_record(2u * j, n_prims, bounds, info), b = filter_record(2u * j + 1u, n_prims, bounds, info);
    filt3[2u * j] = a; filt3[2u * j + 1u] = b;
    pair_xy[j] = make_float4(a.x, b.x, a.y, b.y);
    pair_w[j] = make_float2(a.w, b.w);
#pragma unroll
    for (uint32_t k = 0; k < 2u; k++) {
        const uint32_t prim = 2u * j + k;
        if (prim >= n_prims) { continue; }
        const float4 lo = prim_lo[prim], hi = prim_hi[prim];
        if (!(lo.x <= hi.x)) { continue; }
        const float c[3] = { 0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z };
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            if (isfinite(c[ax])) { atomicMin(&info->cmin_bits[ax], float_to_ordered(c[ax])); atomicMax(&info->cmax_bits[ax], float_to_ordered(c[ax])); }
        }
    }
}
