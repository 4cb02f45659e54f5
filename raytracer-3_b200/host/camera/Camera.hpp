/* camera/Camera.hpp — the camera at the plug-in boundary (reference src/lib/camera/Camera.hpp:25-67):
 * four public vectors and an owned Frame. look_at() is an addition: it fills the same four vectors
 * (plus an optional thin lens) for an arbitrary viewpoint. */
#ifndef RT3_HOST_CAMERA_CAMERA_HPP
#define RT3_HOST_CAMERA_CAMERA_HPP

#include "glm/glm.hpp"
#include "Frame.hpp"

namespace RayTracer {
    class Camera {
    public:
        glm::vec3 origin;
        glm::vec3 horizontal;
        glm::vec3 vertical;
        glm::vec3 lower_left_corner;
        /* Thin-lens extension; lens_radius == 0 is the reference's pinhole. */
        float lens_radius = 0.0f;
        glm::vec3 lens_u, lens_v;

    private:
        Frame* frame;

    public:
        Camera();
        Camera(const Camera& other);
        ~Camera();
        Camera& operator=(const Camera& other) = delete;

        /* Pinhole at the origin looking down -Z (reference camera/Camera.cpp:77-96). Re-allocates the frame. */
        void update(uint32_t width, uint32_t height, float focal_length, float viewport_width, float viewport_height);
        /* RTIOW look-at camera expressed in the same four vectors. Re-allocates the frame. */
        void look_at(uint32_t width, uint32_t height, const glm::vec3& from, const glm::vec3& at, const glm::vec3& up, float vfov_degrees,
                     float aperture, float focus_distance);

        inline uint32_t w() const { return this->frame->w(); }
        inline uint32_t h() const { return this->frame->h(); }
        inline const Frame& get_frame() const { return *this->frame; }
    };
}

#endif
