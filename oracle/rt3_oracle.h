/* rt3_oracle.h — CPU restatement of the per-pixel render path.
 *
 * TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (librt3cuda.so and the host backend) never does.
 *
 * It takes the product ABI's plain-C scene / camera / parameter structs
 * (include/rt3cuda.h) so the same inputs can be handed to both sides.
 */
#ifndef RT3_ORACLE_H
#define RT3_ORACLE_H

#include <stdint.h>
#include "rt3cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Reference-mode ray caster (PINNED: image checked bit-for-bit against the
 * compiled reference, oracle/_ref). Restates
 * reference src/lib/renderer/SequentialRenderer.cpp:47-109 (ray_color) and
 * :284-297 (primary rays, packing), additionally returning the closest-hit
 * primitive id, entity id and t. All rows y in [0, height) are rendered.
 * Analytic spheres (if any) are tested after the faces with the WIP
 * formulation of reference src/lib/shaders/raytracer/raytracer_v4.glsl:157-178. */
int orc_render_reference(const rt3_scene* scene, const rt3_camera* camera, uint32_t width, uint32_t height,
                         uint32_t* frame, uint32_t* hit_prim, uint32_t* hit_entity, float* hit_t);

/* Path tracer (PARITY UNPINNED by the reference, which has no bounce loop,
 * materials or accumulation: raytracer_v4.glsl:279 "let's not bounce just
 * yet", reduce_v1.glsl:66-76). Semantics: Ray Tracing in One Weekend book 1
 * as restated in SURVEY.md appendix C, with fixed-draw sampling from
 * rt3_rng.h. n_threads <= 0 uses all cores (OpenMP over rows).
 * accum (optional) receives the raw fixed-point sums, 3 uint64 per pixel.
 * rays_out (optional) receives the number of ray segments traced. */
int orc_render_pathtrace(const rt3_scene* scene, const rt3_camera* camera, const rt3_params* params,
                         uint32_t* frame, uint64_t* accum, uint64_t* rays_out, int n_threads);

/* RNG pass-throughs (rt3_rng.h) for the integer-parity tests. */
uint32_t orc_hash1(uint32_t x);
uint32_t orc_hash4(uint32_t x, uint32_t y, uint32_t z, uint32_t w);
float orc_float_construct(uint32_t m);
float orc_draw(uint32_t pixel_index, uint32_t sample, uint32_t seed, uint32_t dim);
void orc_sincos_2pi(float x, float* s, float* c);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif

#endif
