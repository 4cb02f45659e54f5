/* ref_driver.cpp — C-callable driver around the UNMODIFIED reference CPU backend.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md). This translation unit is
 * compiled by oracle/Makefile together with the reference's own sources, taken
 * where they lie under /root/reference, into oracle/_ref/libref_seq.so. Nothing
 * in the product path links or loads it; tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs are the only callers.
 *
 * It drives the reference exactly as reference src/Main.cpp:266-288 does:
 *   initialize_renderer() -> Camera::update() -> ECS::create_*() ->
 *   Renderer::prerender() -> Renderer::render() -> read Camera::get_frame().
 * The flattened GFace / vec4 arrays the reference renders from
 * (SequentialRenderer.hpp:28-31, private) are exported verbatim so the CUDA
 * path can be fed the identical scene.
 */
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include <CppDebugger.hpp>

#include "entities/Triangle.hpp"
#include "entities/Sphere.hpp"
#include "entities/Object.hpp"
#include "camera/Camera.hpp"
// The flattened scene is a private member of the reference class; this TU (and
// only this TU) needs to read it back out. Access specifiers do not change
// layout, so the object is the same one the reference's own code built.
#define private public
#include "renderer/SequentialRenderer.hpp"
#undef private

using namespace RayTracer;

namespace {
    struct RefScene {
        Tools::Array<ECS::RenderEntity*> entities;
        Renderer* renderer = nullptr;
        ~RefScene() {
            for (size_t i = 0; i < entities.size(); i++) { delete entities[i]; }
            delete renderer;
        }
    };

    thread_local std::string g_last_error;

    template <class F> int guarded(F&& body) {
        try { body(); return 0; }
        catch (CppDebugger::Fatal& e) { g_last_error = e.what(); return -1; }
        catch (std::exception& e) { g_last_error = e.what(); return -2; }
    }

    glm::vec3 v3(const float* p) { return glm::vec3(p[0], p[1], p[2]); }
}

extern "C" {

const char* ref_last_error() { return g_last_error.c_str(); }

void* ref_scene_create() { return new RefScene(); }
void ref_scene_destroy(void* s) { delete static_cast<RefScene*>(s); }

/* ECS::create_triangle, reference src/lib/entities/Triangle.cpp:28-53. */
int ref_scene_add_triangle(void* s, const float* p1, const float* p2, const float* p3, const float* color) {
    return guarded([&] { static_cast<RefScene*>(s)->entities.push_back(ECS::create_triangle(v3(p1), v3(p2), v3(p3), v3(color))); });
}
/* ECS::create_sphere, reference src/lib/entities/Sphere.cpp:87-115. */
int ref_scene_add_sphere(void* s, const float* center, float radius, uint32_t n_meridians, uint32_t n_parallels, const float* color) {
    return guarded([&] { static_cast<RefScene*>(s)->entities.push_back(ECS::create_sphere(v3(center), radius, n_meridians, n_parallels, v3(color))); });
}
/* ECS::create_object, reference src/lib/entities/Object.cpp:54-126. */
int ref_scene_add_object(void* s, const char* path, const float* center, float scale, const float* color) {
    return guarded([&] { static_cast<RefScene*>(s)->entities.push_back(ECS::create_object(path, v3(center), scale, v3(color))); });
}

/* Renderer::prerender, reference src/lib/renderer/SequentialRenderer.cpp:200-266. */
int ref_scene_prerender(void* s) {
    RefScene* scene = static_cast<RefScene*>(s);
    return guarded([&] {
        if (scene->renderer == nullptr) { scene->renderer = initialize_renderer(); }
        scene->renderer->prerender(scene->entities);
    });
}

uint32_t ref_scene_n_entities(void* s) { return (uint32_t) static_cast<RefScene*>(s)->entities.size(); }
uint32_t ref_scene_n_faces(void* s) {
    RefScene* scene = static_cast<RefScene*>(s);
    return scene->renderer ? (uint32_t) static_cast<SequentialRenderer*>(scene->renderer)->entity_faces.size() : 0;
}
uint32_t ref_scene_n_vertices(void* s) {
    RefScene* scene = static_cast<RefScene*>(s);
    return scene->renderer ? (uint32_t) static_cast<SequentialRenderer*>(scene->renderer)->entity_vertices.size() : 0;
}

/* Copies out the flattened arrays: faces as 48-byte GFace records
 * (reference src/lib/renderer/Vertex.hpp:39-51), vertices as 16-byte vec4. */
int ref_scene_export(void* s, void* faces_out, void* vertices_out) {
    RefScene* scene = static_cast<RefScene*>(s);
    if (!scene->renderer) { g_last_error = "prerender first"; return -1; }
    SequentialRenderer* r = static_cast<SequentialRenderer*>(scene->renderer);
    static_assert(sizeof(GFace) == 48, "GFace layout changed");
    static_assert(sizeof(glm::vec4) == 16, "vec4 layout changed");
    for (size_t i = 0; i < r->entity_faces.size(); i++) { memcpy((char*) faces_out + 48 * i, &r->entity_faces[i], 48); }
    for (size_t i = 0; i < r->entity_vertices.size(); i++) { memcpy((char*) vertices_out + 16 * i, &r->entity_vertices[i], 16); }
    return 0;
}

/* Per-entity (faces, vertices) counts in entity order: the face -> entity map
 * is the prefix sum of these (SequentialRenderer.cpp:174-195). */
int ref_scene_entity_counts(void* s, uint32_t* faces_per_entity, uint32_t* vertices_per_entity) {
    RefScene* scene = static_cast<RefScene*>(s);
    for (size_t i = 0; i < scene->entities.size(); i++) {
        faces_per_entity[i] = scene->entities[i]->pre_render_faces;
        vertices_per_entity[i] = scene->entities[i]->pre_render_vertices;
    }
    return 0;
}

/* Camera::update + Renderer::render (reference Camera.cpp:77-96,
 * SequentialRenderer.cpp:269-308). If cam_override != NULL it holds
 * origin, horizontal, vertical, lower_left_corner (12 floats) written over the
 * camera's public fields after update(). The frame is zero-filled first, so
 * the row the reference never writes (SequentialRenderer.cpp:286) reads 0.
 * seconds_out receives the steady_clock time of render() alone. */
int ref_scene_render(void* s, uint32_t width, uint32_t height, float focal_length, float viewport_width, float viewport_height,
                     const float* cam_override, uint32_t* frame_out, double* seconds_out) {
    RefScene* scene = static_cast<RefScene*>(s);
    if (!scene->renderer) { g_last_error = "prerender first"; return -1; }
    return guarded([&] {
        Camera cam;
        cam.update(width, height, focal_length, viewport_width, viewport_height);
        if (cam_override) {
            cam.origin = v3(cam_override); cam.horizontal = v3(cam_override + 3);
            cam.vertical = v3(cam_override + 6); cam.lower_left_corner = v3(cam_override + 9);
        }
        memset(cam.get_frame().d(), 0, sizeof(uint32_t) * (size_t) width * height);
        auto t0 = std::chrono::steady_clock::now();
        scene->renderer->render(cam);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds_out) { *seconds_out = std::chrono::duration<double>(t1 - t0).count(); }
        memcpy(frame_out, cam.get_frame().d(), sizeof(uint32_t) * (size_t) width * height);
    });
}

/* Reads back the four camera vectors Camera::update produces (12 floats). */
int ref_camera_vectors(uint32_t width, uint32_t height, float focal_length, float viewport_width, float viewport_height, float* out12) {
    return guarded([&] {
        Camera cam;
        cam.update(width, height, focal_length, viewport_width, viewport_height);
        const glm::vec3* v[4] = { &cam.origin, &cam.horizontal, &cam.vertical, &cam.lower_left_corner };
        for (int i = 0; i < 4; i++) { out12[3 * i] = v[i]->x; out12[3 * i + 1] = v[i]->y; out12[3 * i + 2] = v[i]->z; }
    });
}

}  // extern "C"
