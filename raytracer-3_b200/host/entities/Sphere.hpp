/* entities/Sphere.hpp — ECS sphere (reference src/lib/entities/Sphere.hpp:34-51). */
#ifndef RT3_HOST_ENTITIES_SPHERE_HPP
#define RT3_HOST_ENTITIES_SPHERE_HPP

#include "glm/glm.hpp"
#include "RenderEntity.hpp"
#include "renderer/Vertex.hpp"
#include "tools/Array.hpp"

namespace RayTracer::ECS {
    struct Sphere : public RenderEntity {
        glm::vec3 center;
        float radius;
        uint32_t n_meridians;
        uint32_t n_parallels;
        glm::vec3 color;
    };

    Sphere* create_sphere(const glm::vec3& center, float radius, uint32_t n_meridians, uint32_t n_parallels, const glm::vec3& color);
    /* UV-sphere tessellation into pre-sized buffers (pre_render_faces / pre_render_vertices entries). */
    void cpu_pre_render_sphere(Tools::Array<GFace>& faces_buffer, Tools::Array<glm::vec4>& vertex_buffer, Sphere* sphere);
}

#endif
