/* entities/Triangle.cpp — single-triangle entity.
 * Behaviour of reference src/lib/entities/Triangle.cpp:28-76: the normal is
 * normalize(cross(p3 - p1, p2 - p1)) (the negative of the CCW normal, which is
 * why the hit test negates, SequentialRenderer.cpp:87-89) and the colour is
 * stored unshaded. */
#include "Triangle.hpp"

using namespace RayTracer;

ECS::Triangle* ECS::create_triangle(const glm::vec3& p1, const glm::vec3& p2, const glm::vec3& p3, const glm::vec3& color) {
    Triangle* t = new Triangle;
    t->type = et_triangle;
    t->pre_render_mode = eprmf_cpu;
    t->pre_render_operation = epro_generate_triangle;
    t->pre_render_faces = 1;
    t->pre_render_vertices = 3;
    t->points[0] = p1; t->points[1] = p2; t->points[2] = p3;
    t->normal = glm::normalize(glm::cross(p3 - p1, p2 - p1));
    t->color = color;
    return t;
}

void ECS::cpu_pre_render_triangle(Tools::Array<GFace>& faces_buffer, Tools::Array<glm::vec4>& vertex_buffer, Triangle* triangle) {
    faces_buffer.resize(1);
    vertex_buffer.resize(3);
    for (int i = 0; i < 3; i++) { vertex_buffer[i] = glm::vec4(triangle->points[i], 0.0f); }
    GFace& f = faces_buffer[0];
    f.v1 = 0; f.v2 = 1; f.v3 = 2;
    f.normal = triangle->normal;
    f.color = triangle->color;
}
