/* entities/Triangle.hpp — ECS triangle (reference src/lib/entities/Triangle.hpp:28-43). */
#ifndef RT3_HOST_ENTITIES_TRIANGLE_HPP
#define RT3_HOST_ENTITIES_TRIANGLE_HPP

#include "glm/glm.hpp"
#include "RenderEntity.hpp"
#include "renderer/Vertex.hpp"
#include "tools/Array.hpp"

namespace RayTracer::ECS {
    struct Triangle : public RenderEntity {
        glm::vec3 points[3];
        glm::vec3 normal;
        glm::vec3 color;
    };

    Triangle* create_triangle(const glm::vec3& p1, const glm::vec3& p2, const glm::vec3& p3, const glm::vec3& color);
    void cpu_pre_render_triangle(Tools::Array<GFace>& faces_buffer, Tools::Array<glm::vec4>& vertex_buffer, Triangle* triangle);
}

#endif
