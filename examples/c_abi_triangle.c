/* c_abi_triangle.c — the C ABI of include/rt3cuda.h from plain C, no C++ and no Python in between.
 *
 * Renders the reference's single-triangle scene (reference src/Main.cpp:279,
 * create_triangle({1,0,-3}, {-1,0,-3}, {0,1,-3}, {1,0,0})) with the camera of Main.cpp:272 in
 * RT3_MODE_REFERENCE and prints the 64-bit hash of rows 0..H-2 (the rows the reference's CPU loop
 * writes, SequentialRenderer.cpp:286), which tests/golden/goldens.json holds for the compiled reference.
 *
 *   gcc -std=c99 -I include examples/c_abi_triangle.c -L raytracer-3_b200/csrc -lrt3cuda \
 *       -Wl,-rpath,$PWD/raytracer-3_b200/csrc -o c_abi_triangle && ./c_abi_triangle
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rt3cuda.h"

static int check(int status, const char* what) {
    if (status != RT3_OK) { fprintf(stderr, "%s: %s\n", what, rt3_last_error()); }
    return status;
}

int main(void) {
    enum { W = 400, H = 225 };
    rt3_vertex vertices[3] = { { 1.f, 0.f, -3.f, 0.f }, { -1.f, 0.f, -3.f, 0.f }, { 0.f, 1.f, -3.f, 0.f } };
    rt3_face face;
    rt3_scene scene;
    rt3_camera cam;
    rt3_params params;
    rt3_stats stats;
    rt3_ctx* ctx = NULL;
    uint32_t* frame = (uint32_t*) malloc(sizeof(uint32_t) * W * H);
    unsigned long long h = 0xcbf29ce484222325ull;
    const float vw = ((float) W / (float) H) * 2.0f, vh = 2.0f, focal = 2.0f;
    size_t i;

    memset(&face, 0, sizeof face);
    face.v1 = 0; face.v2 = 1; face.v3 = 2;
    face.normal[2] = 1.0f;   /* normalize(cross(p3 - p1, p2 - p1)) = (0, 0, 1), Triangle.cpp:48 */
    face.color[0] = 1.0f;
    memset(&scene, 0, sizeof scene);
    scene.n_faces = 1; scene.n_vertices = 3; scene.faces = &face; scene.vertices = vertices;

    memset(&cam, 0, sizeof cam);   /* Camera::update(W, H, 2, (W/H)*2, 2), Camera.cpp:77-96 */
    cam.horizontal[0] = vw;
    cam.vertical[1] = vh;
    cam.lower_left_corner[0] = 0.0f - vw / 2.0f;
    cam.lower_left_corner[1] = 0.0f - vh / 2.0f;
    cam.lower_left_corner[2] = 0.0f - focal;

    memset(&params, 0, sizeof params);
    params.width = W; params.height = H; params.mode = RT3_MODE_REFERENCE;

    if (!frame) { return 2; }
    if (check(rt3_create(&ctx, 0), "rt3_create") != RT3_OK) { return 1; }
    if (check(rt3_scene_upload(ctx, &scene), "rt3_scene_upload") != RT3_OK) { return 1; }
    if (check(rt3_render(ctx, &cam, &params, frame), "rt3_render") != RT3_OK) { return 1; }
    if (check(rt3_get_stats(ctx, &stats), "rt3_get_stats") != RT3_OK) { return 1; }
    for (i = 0; i < (size_t) W * (H - 1); i++) { h = (h ^ frame[i]) * 0x100000001b3ull; }
    printf("hash %016llx centre %08x rays %llu device_ms %.3f\n", h, (unsigned) frame[(H / 2) * W + W / 2], (unsigned long long) stats.rays,
           stats.device_ms);
    rt3_destroy(ctx);
    free(frame);
    return 0;
}
