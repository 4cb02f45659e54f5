/* renderer/Vertex.hpp — the flattened face record (reference src/lib/renderer/Vertex.hpp:39-51).
 * Same 48-byte layout as the reference's GFace and as rt3_face in include/rt3cuda.h, so the
 * flattened array is handed to rt3_scene_upload without conversion. */
#ifndef RT3_HOST_RENDERER_VERTEX_HPP
#define RT3_HOST_RENDERER_VERTEX_HPP

#include <cstdint>

#include "glm/glm.hpp"

namespace RayTracer {
    struct GFace {
        alignas(4) uint32_t v1;
        alignas(4) uint32_t v2;
        alignas(4) uint32_t v3;
        alignas(16) glm::vec3 normal;
        alignas(16) glm::vec3 color;
    };
    static_assert(sizeof(GFace) == 48, "GFace must match the reference layout");
}

#endif
