/* entities/Object.hpp — ECS Wavefront-OBJ mesh (reference src/lib/entities/Object.hpp:29-46). */
#ifndef RT3_HOST_ENTITIES_OBJECT_HPP
#define RT3_HOST_ENTITIES_OBJECT_HPP

#include <string>

#include "glm/glm.hpp"
#include "RenderEntity.hpp"
#include "renderer/Vertex.hpp"
#include "tools/Array.hpp"

namespace RayTracer::ECS {
    struct Object : public RenderEntity {
        std::string file_path;
        glm::vec3 center;
        float scale;
        glm::vec3 color;
    };

    /* Counts the file's v / f records; the geometry itself is read during pre-rendering. */
    Object* create_object(const std::string& file_path, const glm::vec3& center, float scale, const glm::vec3& color);
    void cpu_pre_render_object(Tools::Array<GFace>& faces_buffer, Tools::Array<glm::vec4>& vertex_buffer, Object* obj);
}

#endif
