/* camera/Frame.hpp — the host framebuffer (reference src/lib/camera/Frame.hpp:41-82).
 * uint32 per pixel, r<<24 | g<<16 | b<<8 | a, index y*w + x, row 0 at the top. */
#ifndef RT3_HOST_CAMERA_FRAME_HPP
#define RT3_HOST_CAMERA_FRAME_HPP

#include <cstdint>
#include <string>

namespace RayTracer {
    class Frame {
        uint32_t* data;
        uint32_t width;
        uint32_t height;

    public:
        Frame(uint32_t width, uint32_t height);
        Frame(const Frame& other);
        Frame(Frame&& other);
        ~Frame();
        Frame& operator=(const Frame& other) = delete;

        /* Binary PPM (P6), reference Frame.cpp:110-148 (without its per-pixel log line, Frame.cpp:137). */
        void to_ppm(const std::string& path) const;
        /* 8-bit RGBA PNG like reference Frame.cpp:82-106, written by a small encoder of its own (zlib "stored"
         * blocks: valid for every PNG reader, no compression) instead of the reference's vendored LodePNG. */
        void to_png(const std::string& path) const;

        inline uint32_t w() const { return this->width; }
        inline uint32_t h() const { return this->height; }
        /* Mutable access from a const Frame, as in the reference (Frame.hpp:70): renderers write through it. */
        inline uint32_t* d() const { return this->data; }
    };
}

#endif
