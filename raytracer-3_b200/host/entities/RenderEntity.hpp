/* entities/RenderEntity.hpp — the ECS entity header (reference src/lib/entities/RenderEntity.hpp:24-89). */
#ifndef RT3_HOST_ENTITIES_RENDER_ENTITY_HPP
#define RT3_HOST_ENTITIES_RENDER_ENTITY_HPP

#include <cstdint>
#include <string>

namespace RayTracer::ECS {
    enum EntityType { et_none = 0, et_triangle = 1, et_sphere = 2, et_object = 3 };
    static const std::string entity_type_names[] = { "none", "triangle", "sphere", "object" };

    enum EntityPreRenderModeFlags { eprmf_none = 0x0, eprmf_cpu = 0x1, eprmf_gpu = 0x2 };

    enum EntityPreRenderOperation { epro_none = 0, epro_generate_triangle = 1, epro_generate_sphere = 2, epro_load_object_file = 3 };
    static const std::string entity_pre_render_operation_names[] = { "none", "generate_triangle", "generate_sphere", "load_object_file" };

    /* Plain header every entity starts with; face / vertex counts are known before pre-rendering. */
    struct RenderEntity {
        EntityType type;
        unsigned int pre_render_mode;
        EntityPreRenderOperation pre_render_operation;
        uint32_t pre_render_faces;
        uint32_t pre_render_vertices;
    };
}

#endif
