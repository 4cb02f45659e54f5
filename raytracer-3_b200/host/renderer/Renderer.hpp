/* renderer/Renderer.hpp — the renderer plug-in interface (reference src/lib/renderer/Renderer.hpp:34-63).
 * Exactly one backend defines initialize_renderer(); here that is CudaRenderer.cpp. */
#ifndef RT3_HOST_RENDERER_RENDERER_HPP
#define RT3_HOST_RENDERER_RENDERER_HPP

#include "glm/glm.hpp"
#include "camera/Camera.hpp"
#include "entities/RenderEntity.hpp"
#include "tools/Array.hpp"
#include "Vertex.hpp"

namespace RayTracer {
    class Renderer {
    protected:
        Renderer() {}

    public:
        Renderer(const Renderer&) {}
        virtual ~Renderer() {}

        /* Flattens the entities for rendering. The caller keeps ownership and may delete them afterwards. */
        virtual void prerender(const Tools::Array<ECS::RenderEntity*>& entities) = 0;
        /* Renders into camera.get_frame(). Blocking. */
        virtual void render(Camera& camera) const = 0;
    };

    /* Factory implemented by the linked backend. */
    extern Renderer* initialize_renderer();
}

#endif
