/* rt3_rng.h — counter-based RNG and fixed-draw sampling primitives.
 *
 * Integer-exact on host and device: the same header is compiled by gcc (the
 * CPU oracle) and nvcc (the kernels), so both sides draw identical numbers.
 *
 * The hash and the mantissa-stuffing float construction restate
 * reference src/lib/shaders/random_v1.glsl:22-35 (_random_hash, scalar and
 * uvec4 forms) and :38-53 (_random_float_construct). The reference includes
 * that file from its render shader (raytracer_v3.glsl:24) but never calls it;
 * the counter layout below is this project's choice (DESIGN.md "RNG").
 *
 * Every floating-point operation in this header is a single IEEE-754 binary32
 * operation in a fixed order (no contraction: build hosts with
 * -ffp-contract=off, nvcc with -fmad=false), so results are bit-identical
 * across CPU and GPU.
 */
#ifndef RT3_RNG_H
#define RT3_RNG_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RT3_HD __host__ __device__ __forceinline__
#else
#define RT3_HD static inline
#endif

/* random_v1.glsl:22-29 */
RT3_HD uint32_t rt3_hash1(uint32_t x) {
    x += (x << 10u);
    x ^= (x >> 6u);
    x += (x << 3u);
    x ^= (x >> 11u);
    x += (x << 15u);
    return x;
}

/* random_v1.glsl:35 — uvec4 form: hash(x ^ hash(y) ^ hash(z) ^ hash(w)). */
RT3_HD uint32_t rt3_hash4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    return rt3_hash1(x ^ rt3_hash1(y) ^ rt3_hash1(z) ^ rt3_hash1(w));
}

/* random_v1.glsl:38-53 — 23 random mantissa bits under exponent 0 -> [1,2) -> [0,1). */
RT3_HD float rt3_float_construct(uint32_t m) {
    m &= 0x007FFFFFu;
    m |= 0x3F800000u;
    float f;
#if defined(__CUDA_ARCH__)
    f = __uint_as_float(m);
#else
    memcpy(&f, &m, sizeof f);
#endif
    return f - 1.0f;
}

/* Dimension tags. The uvec4 hash XORs the hashes of its y/z/w lanes, so it is
 * symmetric in them; tagging the dimension lane with the top bit keeps
 * (sample, dim) and (dim, sample) from colliding (samples stay < 2^31). */
#define RT3_DIM_TAG 0x80000000u
#define RT3_DIM_JITTER_X 0u
#define RT3_DIM_JITTER_Y 1u
#define RT3_DIM_LENS_R 2u
#define RT3_DIM_LENS_PHI 3u
#define RT3_DIM_BOUNCE0 4u /* bounce b uses dims RT3_DIM_BOUNCE0 + 8*b + k, k in 0..7 */
#define RT3_DIMS_PER_BOUNCE 8u

/* The per-path part of the counter: everything but the dimension. */
RT3_HD uint32_t rt3_path_key(uint32_t pixel_index, uint32_t sample, uint32_t seed) {
    return pixel_index ^ rt3_hash1(sample) ^ rt3_hash1(seed);
}

/* One uniform draw in [0,1): == float(hash4(pixel, sample, TAG|dim, seed)). */
RT3_HD float rt3_draw(uint32_t path_key, uint32_t dim) {
    return rt3_float_construct(rt3_hash1(path_key ^ rt3_hash1(RT3_DIM_TAG | dim)));
}

/* sin and cos of 2*pi*x for x in [0,1), from fixed polynomials evaluated with
 * plain multiplies and adds in a fixed order (no libm, no FMA), so host and
 * device agree bit for bit. Octant reduction keeps the polynomial argument
 * in [-pi/4, pi/4]; absolute error is below 2e-7. */
RT3_HD void rt3_sincos_2pi(float x, float* s_out, float* c_out) {
    float y = x * 8.0f;                 /* exact */
    int oct = (int) y;                  /* 0..7 */
    int q = (oct + 1) >> 1;             /* nearest multiple of pi/2: 0..4 */
    float r = (y - 2.0f * (float) q) * 0.78539816339744831f; /* (x - q/4) * 2pi, |r| <= pi/4 */
    float r2 = r * r;
    /* sin r = r + r^3 * (S1 + r2*(S2 + r2*S3)) */
    float sp = 8.3321608736e-3f + r2 * -1.9515295891e-4f;
    sp = -1.6666654611e-1f + r2 * sp;
    float sn = r + (r * r2) * sp;
    /* cos r = 1 - r2/2 + r^4 * (C1 + r2*(C2 + r2*C3)) */
    float cp = -1.388731625493765e-3f + r2 * 2.443315711809948e-5f;
    cp = 4.166664568298827e-2f + r2 * cp;
    float cs = (1.0f - 0.5f * r2) + (r2 * r2) * cp;
    float s, c;
    switch (q & 3) {
        case 0: s = sn; c = cs; break;
        case 1: s = cs; c = -sn; break;
        case 2: s = -sn; c = -cs; break;
        default: s = -cs; c = sn; break;
    }
    *s_out = s; *c_out = c;
}

#endif /* RT3_RNG_H */
