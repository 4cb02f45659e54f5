/* rt3_oracle.c — CPU restatement of the per-pixel render path (plain C11).
 *
 * TEST INFRASTRUCTURE, not product code: see rt3_oracle.h. Built by
 * oracle/Makefile with -O2 -ffp-contract=off -fno-fast-math: every float
 * operation below is one IEEE-754 binary32 operation, in the order written
 * (fused multiply-adds appear only where fmaf() is spelled out: the path-mode
 * sphere discriminant).
 *
 * Parity status
 *   orc_render_reference : PINNED. tests/test_oracle_pinned.py requires its
 *       packed image to equal, bit for bit, the image the compiled reference
 *       (oracle/_ref/libref_seq.so, built from /root/reference) renders for
 *       the same scenes, and checks the committed goldens under tests/golden/.
 *   orc_render_pathtrace : PARITY UNPINNED by the reference (it has no
 *       bounce loop, materials, jitter, accumulation or gamma). Its depth-1,
 *       no-jitter primary hits are cross-checked against orc_render_reference.
 */
#include "rt3_oracle.h"
#include "rt3_rng.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* Row-parallel driver: worker threads pull row indices from an atomic counter.
 * (Plain pthreads: no OpenMP runtime to clash with the one torch bundles.) */
typedef void (*row_fn)(void* ctx, uint32_t y);
typedef struct { row_fn fn; void* ctx; uint32_t height; atomic_uint next; } row_pool;

static void* row_worker(void* arg) {
    row_pool* pool = (row_pool*) arg;
    for (;;) {
        uint32_t y = atomic_fetch_add(&pool->next, 1u);
        if (y >= pool->height) { break; }
        pool->fn(pool->ctx, y);
    }
    return NULL;
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int) n : 1;
}

static void for_each_row(uint32_t height, int n_threads, row_fn fn, void* ctx) {
    if (n_threads <= 0) { n_threads = orc_max_threads(); }
    if (n_threads > 1024) { n_threads = 1024; }
    row_pool pool;
    pool.fn = fn; pool.ctx = ctx; pool.height = height;
    atomic_init(&pool.next, 0u);
    if (n_threads == 1) { row_worker(&pool); return; }
    pthread_t* threads = (pthread_t*) malloc(sizeof(pthread_t) * (size_t) n_threads);
    int started = 0;
    for (int i = 0; i < n_threads; i++) { if (pthread_create(&threads[started], NULL, row_worker, &pool) == 0) { started++; } }
    if (started == 0) { row_worker(&pool); }
    for (int i = 0; i < started; i++) { pthread_join(threads[i], NULL); }
    free(threads);
}

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 vld(const float* p) { return V(p[0], p[1], p[2]); }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* dot3 macro, SequentialRenderer.cpp:32-33; identical to glm::dot's (x+y)+z. */
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* glm::cross, glm/detail/func_geometric.inl:74-77. */
static inline v3 cross3(v3 x, v3 y) {
    return V(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
/* glm::normalize = v * inversesqrt(dot(v,v)), func_geometric.inl:88, func_exponential.inl:138. */
static inline v3 normalize3(v3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return V(a.x * inv, a.y * inv, a.z * inv); }

/* glm::packUnorm4x8(vec4(1.0, b, g, r)), SequentialRenderer.cpp:297 with
 * glm/detail/func_packing.inl:67-83: round(clamp(v,0,1)*255), std::round. */
static inline uint32_t unorm8(float c) {
    float m = (c < 0.0f) ? 0.0f : c;  /* glm::max(x, 0): (x < y) ? y : x */
    m = (1.0f < m) ? 1.0f : m;        /* glm::min(x, 1): (y < x) ? y : x */
    if (m != m) { m = 0.0f; }         /* NaN is undefined in the reference; defined as 0 here */
    return (uint32_t) (unsigned char) roundf(m * 255.0f);
}
static inline uint32_t pack_rgb(v3 c) { return (unorm8(c.x) << 24) | (unorm8(c.y) << 16) | (unorm8(c.z) << 8) | 0xFFu; }

/* Sky gradient, SequentialRenderer.cpp:105-107. `d` is the un-normalised direction. */
static inline v3 sky_reference(v3 d) {
    float len = sqrtf(dot3(d, d));
    float uy = d.y / len;
    float t = (float) (0.5 * ((double) uy + 1.0));
    float a = 1.0f - t;
    return V(a * 1.0f + t * 0.5f, a * 1.0f + t * 0.7f, a * 1.0f + t * 1.0f);
}

typedef struct { uint32_t prim; float t; } hit_rec;
#define NO_HIT 0xFFFFFFFFu

/* Closest hit over the faces: SequentialRenderer.cpp:50-96. `tmin` is 0 in
 * reference mode (the reference rejects t < 0 only). The plane offset uses
 * n.p1 - n.o, which is bit-identical to the reference's n.o + n.p1 for the
 * origin-at-zero camera it always uses. */
static hit_rec closest_face(const rt3_scene* s, v3 o, v3 d, float tmin, hit_rec best) {
    for (uint32_t i = 0; i < s->n_faces; i++) {
        const rt3_face* f = &s->faces[i];
        v3 n = vld(f->normal);
        float nd = dot3(d, n);
        if (nd == 0) { continue; }
        v3 p1 = vld(&s->vertices[f->v1].x), p2 = vld(&s->vertices[f->v2].x), p3 = vld(&s->vertices[f->v3].x);
        float pd = dot3(n, p1);
        float t = (pd - dot3(n, o)) / dot3(n, d);
        if (t < tmin || t >= best.t) { continue; }
        v3 hp = vadd(o, vscale(t, d));
        v3 a = cross3(vsub(p2, p1), vsub(hp, p1));
        v3 b = cross3(vsub(p3, p2), vsub(hp, p2));
        v3 c = cross3(vsub(p1, p3), vsub(hp, p3));
        if (-dot3(n, a) >= 0.0 && -dot3(n, b) >= 0.0 && -dot3(n, c) >= 0.0) { best.prim = i; best.t = t; }
    }
    return best;
}

/* Reference-mode analytic sphere: the WIP hit_sphere, raytracer_v4.glsl:157-178
 * (abc form, near root only, t >= 0), un-normalised direction. */
static hit_rec closest_sphere_v4(const rt3_scene* s, v3 o, v3 d, hit_rec best) {
    for (uint32_t i = 0; i < s->n_spheres; i++) {
        const rt3_sphere* sp = &s->spheres[i];
        v3 oc = vsub(o, V(sp->cx, sp->cy, sp->cz));
        float a = dot3(d, d);
        float b = 2.0f * dot3(oc, d);
        float c = dot3(oc, oc) - sp->r * sp->r;
        float D = b * b - (4.0f * a) * c;
        if (D >= 0) {
            float t = (-b - sqrtf(D)) / (2.0f * a);
            if (t >= 0.0f && t < best.t) { best.prim = s->n_faces + i; best.t = t; }
        }
    }
    return best;
}

static inline uint32_t prim_entity(const rt3_scene* s, uint32_t prim) {
    if (prim == NO_HIT) { return NO_HIT; }
    if (prim < s->n_faces) { return s->face_entity ? s->face_entity[prim] : 0u; }
    return s->sphere_entity ? s->sphere_entity[prim - s->n_faces] : 0u;
}

typedef struct {
    const rt3_scene* scene; const rt3_camera* cam; uint32_t width, height;
    uint32_t* frame; uint32_t* hit_prim; uint32_t* hit_entity; float* hit_t;
} ref_job;

static void reference_row(void* ctx, uint32_t y) {
    const ref_job* j = (const ref_job*) ctx;
    const rt3_scene* scene = j->scene; const rt3_camera* cam = j->cam;
    uint32_t width = j->width, height = j->height;
    uint32_t* frame = j->frame; uint32_t* hit_prim = j->hit_prim; uint32_t* hit_entity = j->hit_entity; float* hit_t = j->hit_t;
    v3 origin = vld(cam->origin), hor = vld(cam->horizontal), ver = vld(cam->vertical), llc = vld(cam->lower_left_corner);
    {
        for (uint32_t x = 0; x < width; x++) {
            /* SequentialRenderer.cpp:289-290 (the divide is written in double there). */
            float u = (float) ((double) (float) x / ((double) (float) width - 1.0));
            float v = (float) ((double) (float) (height - 1 - y) / ((double) (float) height - 1.0));
            /* :293 */
            v3 ray = vsub(vadd(vadd(llc, vscale(u, hor)), vscale(v, ver)), origin);
            hit_rec best = { NO_HIT, INFINITY };  /* float(1e99) == +inf, :52 */
            best = closest_face(scene, origin, ray, 0.0f, best);
            best = closest_sphere_v4(scene, origin, ray, best);
            v3 col;
            if (best.prim == NO_HIT) { col = sky_reference(ray); }
            else if (best.prim < scene->n_faces) { col = vld(scene->faces[best.prim].color); }
            else { col = vld(&scene->sphere_color[3 * (best.prim - scene->n_faces)]); }
            size_t idx = (size_t) y * width + x;
            if (frame) { frame[idx] = pack_rgb(col); }
            if (hit_prim) { hit_prim[idx] = best.prim; }
            if (hit_entity) { hit_entity[idx] = prim_entity(scene, best.prim); }
            if (hit_t) { hit_t[idx] = best.t; }
        }
    }
}

int orc_render_reference(const rt3_scene* scene, const rt3_camera* cam, uint32_t width, uint32_t height,
                         uint32_t* frame, uint32_t* hit_prim, uint32_t* hit_entity, float* hit_t) {
    if (!scene || !cam || width < 2 || height < 2) { return -1; }
    ref_job job = { scene, cam, width, height, frame, hit_prim, hit_entity, hit_t };
    for_each_row(height, 0, reference_row, &job);
    return 0;
}

/* ------------------------------------------------------------------------ *
 * Path tracer (structure of raytracer_v4.glsl:183-291)
 * ------------------------------------------------------------------------ */

#define RT3_TMIN 0.001f
#define RT3_ACC_SCALE 16777216.0f /* 2^24 fixed-point radiance */

/* Path-mode sphere: half-b form with a unit direction, near root then far
 * root, accepted iff tmin <= t < best (appendix C "Sphere hit"). The dot
 * products and the discriminant are chains of fused multiply-adds (fmaf, one
 * rounding each), written out in a fixed order -- the contraction a GLSL
 * compiler may apply to raytracer_v4.glsl:157-178, made explicit so that the
 * CUDA kernel (exact_sphere_path, __fmaf_rn in the same order) agrees bit for bit. */
/* fmaf() through libm is a slow software routine where glibc has no FMA dispatch for it; cloned for FMA hardware
 * (resolved when the library loads) so that the CPU baseline is not handicapped. The results are the same: one
 * rounding per fmaf either way. */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("fma", "default")))
#endif
static hit_rec closest_sphere_path(const rt3_scene* s, v3 o, v3 d, hit_rec best) {
    for (uint32_t i = 0; i < s->n_spheres; i++) {
        const rt3_sphere* sp = &s->spheres[i];
        v3 oc = vsub(o, V(sp->cx, sp->cy, sp->cz));
        float h = fmaf(oc.x, d.x, fmaf(oc.y, d.y, oc.z * d.z));
        float c = fmaf(oc.x, oc.x, fmaf(oc.y, oc.y, fmaf(oc.z, oc.z, -(sp->r * sp->r))));
        float disc = fmaf(h, h, -c);
        if (!(disc >= 0.0f)) { continue; }
        float sq = sqrtf(disc);
        float t = -h - sq;
        if (!(t >= RT3_TMIN && t < best.t)) {
            t = -h + sq;
            if (!(t >= RT3_TMIN && t < best.t)) { continue; }
        }
        best.prim = s->n_faces + i; best.t = t;
    }
    return best;
}

static inline void material_of(const rt3_scene* s, uint32_t prim, rt3_material* out) {
    const uint32_t* idx = (prim < s->n_faces) ? s->face_material : s->sphere_material;
    uint32_t local = (prim < s->n_faces) ? prim : prim - s->n_faces;
    if (idx) { *out = s->materials[idx[local]]; return; }
    const float* col = (prim < s->n_faces) ? s->faces[prim].color : &s->sphere_color[3 * local];
    out->kind = RT3_MAT_LAMBERTIAN; out->albedo[0] = col[0]; out->albedo[1] = col[1]; out->albedo[2] = col[2];
    out->fuzz = 0.0f; out->ior = 1.0f;
}

/* Uniform point on the unit sphere from two draws: z = 1 - 2*xi1, phi = 2*pi*xi2. */
static inline v3 unit_vector(float xi1, float xi2) {
    float z = 1.0f - 2.0f * xi1;
    float rr = 1.0f - z * z;
    rr = sqrtf(rr < 0.0f ? 0.0f : rr);
    float sn, cs;
    rt3_sincos_2pi(xi2, &sn, &cs);
    return V(rr * cs, rr * sn, z);
}

static inline float max3f(float a, float b, float c) { float m = a < b ? b : a; return m < c ? c : m; }

/* One path; returns its radiance and adds its segment count to *rays. */
static v3 trace_path(const rt3_scene* s, const rt3_camera* cam, const rt3_params* p, uint32_t x, uint32_t y,
                     uint32_t sample, uint64_t* rays) {
    uint32_t pixel_index = y * p->width + x;
    uint32_t key = rt3_path_key(pixel_index, sample, p->seed);
    float jx = 0.0f, jy = 0.0f;
    if (!(p->flags & RT3_FLAG_NO_JITTER)) { jx = rt3_draw(key, RT3_DIM_JITTER_X); jy = rt3_draw(key, RT3_DIM_JITTER_Y); }
    float u = ((float) x + jx) / ((float) p->width - 1.0f);
    float v = ((float) (p->height - 1 - y) + jy) / ((float) p->height - 1.0f);
    v3 o = vld(cam->origin);
    v3 dir = vsub(vadd(vadd(vld(cam->lower_left_corner), vscale(u, vld(cam->horizontal))), vscale(v, vld(cam->vertical))), o);
    if (cam->lens_radius > 0.0f) {
        float r = sqrtf(rt3_draw(key, RT3_DIM_LENS_R));
        float sn, cs;
        rt3_sincos_2pi(rt3_draw(key, RT3_DIM_LENS_PHI), &sn, &cs);
        float lx = cam->lens_radius * (r * cs), ly = cam->lens_radius * (r * sn);
        v3 off = vadd(vscale(lx, vld(cam->lens_u)), vscale(ly, vld(cam->lens_v)));
        o = vadd(o, off);
        dir = vsub(dir, off);
    }
    v3 d = normalize3(dir);
    v3 thr = V(1.0f, 1.0f, 1.0f);
    for (uint32_t bounce = 0; bounce < p->max_depth; bounce++) {
        hit_rec best = { NO_HIT, INFINITY };
        best = closest_face(s, o, d, RT3_TMIN, best);
        best = closest_sphere_path(s, o, d, best);
        (*rays)++;
        if (best.prim == NO_HIT) {
            /* sky, SequentialRenderer.cpp:105-107, for an already-unit direction */
            if (p->flags & RT3_FLAG_UNIFORM_SKY) { return thr; } /* white furnace */
            float t = 0.5f * (d.y + 1.0f);
            float a = 1.0f - t;
            return vmul(thr, V(a * 1.0f + t * 0.5f, a * 1.0f + t * 0.7f, a * 1.0f + t * 1.0f));
        }
        v3 hp = vadd(o, vscale(best.t, d));
        v3 outward;
        if (best.prim < s->n_faces) {
            outward = vld(s->faces[best.prim].normal);
        } else {
            const rt3_sphere* sp = &s->spheres[best.prim - s->n_faces];
            v3 pc = vsub(hp, V(sp->cx, sp->cy, sp->cz));
            outward = V(pc.x / sp->r, pc.y / sp->r, pc.z / sp->r);
        }
        int front = dot3(d, outward) < 0.0f;
        v3 n = front ? outward : vneg(outward);
        rt3_material m;
        material_of(s, best.prim, &m);
        uint32_t dim = RT3_DIM_BOUNCE0 + RT3_DIMS_PER_BOUNCE * bounce;
        v3 nd;
        if (m.kind == RT3_MAT_LAMBERTIAN) {
            v3 uv = unit_vector(rt3_draw(key, dim + 0), rt3_draw(key, dim + 1));
            nd = vadd(n, uv);
            if (fabsf(nd.x) < 1e-8f && fabsf(nd.y) < 1e-8f && fabsf(nd.z) < 1e-8f) { nd = n; }
            thr = vmul(thr, vld(m.albedo));
        } else if (m.kind == RT3_MAT_METAL) {
            float dn = dot3(d, n);
            nd = vsub(d, vscale(2.0f * dn, n));
            float fuzz = m.fuzz < 1.0f ? m.fuzz : 1.0f;
            if (fuzz > 0.0f) {
                v3 uv = unit_vector(rt3_draw(key, dim + 0), rt3_draw(key, dim + 1));
                /* radius with density 3r^2: the max of three uniforms */
                float rad = max3f(rt3_draw(key, dim + 2), rt3_draw(key, dim + 3), rt3_draw(key, dim + 4));
                nd = vadd(nd, vscale(fuzz * rad, uv));
            }
            if (!(dot3(nd, n) > 0.0f)) { return V(0.0f, 0.0f, 0.0f); }
            thr = vmul(thr, vld(m.albedo));
        } else {
            float ratio = front ? (1.0f / m.ior) : m.ior;
            float cs = -dot3(d, n);
            cs = cs < 1.0f ? cs : 1.0f;
            float s2 = 1.0f - cs * cs;
            float sn = sqrtf(s2 < 0.0f ? 0.0f : s2);
            int cannot_refract = ratio * sn > 1.0f;
            float r0 = (1.0f - ratio) / (1.0f + ratio);
            r0 = r0 * r0;
            float w = 1.0f - cs;
            float w2 = w * w;
            float schlick = r0 + (1.0f - r0) * ((w2 * w2) * w);
            if (cannot_refract || schlick > rt3_draw(key, dim + 0)) {
                float dn = dot3(d, n);
                nd = vsub(d, vscale(2.0f * dn, n));
            } else {
                v3 perp = vscale(ratio, vadd(d, vscale(cs, n)));
                float k = 1.0f - dot3(perp, perp);
                v3 par = vscale(-sqrtf(fabsf(k)), n);
                nd = vadd(perp, par);
            }
        }
        o = hp;
        d = normalize3(nd);
    }
    return V(0.0f, 0.0f, 0.0f);
}

/* Radiance -> 2^24 fixed point, saturating; NaN and negatives count as 0. */
static inline uint64_t to_fixed(float c) {
    if (!(c > 0.0f)) { return 0; }
    if (c > 1048576.0f) { c = 1048576.0f; }
    return (uint64_t) (c * RT3_ACC_SCALE + 0.5f);
}

static inline uint32_t resolve_channel(uint64_t sum, uint32_t spp, int gamma) {
    float m = (float) ((double) sum / ((double) spp * (double) RT3_ACC_SCALE));
    if (gamma) { m = sqrtf(m); }
    return unorm8(m);
}

typedef struct {
    const rt3_scene* scene; const rt3_camera* cam; const rt3_params* p;
    uint32_t* frame; uint64_t* accum; atomic_ullong rays_total;
} path_job;

static void pathtrace_row(void* ctx, uint32_t y) {
    path_job* j = (path_job*) ctx;
    const rt3_scene* scene = j->scene; const rt3_camera* cam = j->cam; const rt3_params* p = j->p;
    uint32_t* frame = j->frame; uint64_t* accum = j->accum;
    uint32_t tile_rows = p->tile_rows ? p->tile_rows : 1;
    uint32_t parts = p->part_count ? p->part_count : 1;
    int gamma = !(p->flags & RT3_FLAG_NO_GAMMA);
    uint64_t rays_total = 0;
    if ((y / tile_rows) % parts != p->part_index % parts) { return; }
    {
        for (uint32_t x = 0; x < p->width; x++) {
            uint64_t acc[3] = { 0, 0, 0 };
            uint64_t rays = 0;
            for (uint32_t sidx = 0; sidx < p->spp; sidx++) {
                v3 L = trace_path(scene, cam, p, x, y, p->first_sample + sidx, &rays); /* a one-shot render of samples [first_sample, first_sample + spp) */
                acc[0] += to_fixed(L.x); acc[1] += to_fixed(L.y); acc[2] += to_fixed(L.z);
            }
            rays_total += rays;
            size_t idx = (size_t) y * p->width + x;
            if (accum) { accum[3 * idx] = acc[0]; accum[3 * idx + 1] = acc[1]; accum[3 * idx + 2] = acc[2]; }
            if (frame) {
                frame[idx] = (resolve_channel(acc[0], p->spp, gamma) << 24) | (resolve_channel(acc[1], p->spp, gamma) << 16) |
                             (resolve_channel(acc[2], p->spp, gamma) << 8) | 0xFFu;
            }
        }
    }
    atomic_fetch_add(&j->rays_total, rays_total);
}

int orc_render_pathtrace(const rt3_scene* scene, const rt3_camera* cam, const rt3_params* p,
                         uint32_t* frame, uint64_t* accum, uint64_t* rays_out, int n_threads) {
    if (!scene || !cam || !p || p->width < 2 || p->height < 2 || p->spp < 1 || p->max_depth < 1) { return -1; }
    path_job job;
    job.scene = scene; job.cam = cam; job.p = p; job.frame = frame; job.accum = accum;
    atomic_init(&job.rays_total, 0ull);
    for_each_row(p->height, n_threads, pathtrace_row, &job);
    if (rays_out) { *rays_out = (uint64_t) atomic_load(&job.rays_total); }
    return 0;
}

uint32_t orc_hash1(uint32_t x) { return rt3_hash1(x); }
uint32_t orc_hash4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return rt3_hash4(x, y, z, w); }
float orc_float_construct(uint32_t m) { return rt3_float_construct(m); }
float orc_draw(uint32_t pixel_index, uint32_t sample, uint32_t seed, uint32_t dim) {
    return rt3_draw(rt3_path_key(pixel_index, sample, seed), dim);
}
void orc_sincos_2pi(float x, float* s, float* c) { rt3_sincos_2pi(x, s, c); }
