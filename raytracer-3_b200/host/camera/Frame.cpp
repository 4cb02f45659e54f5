/* camera/Frame.cpp — see Frame.hpp (reference src/lib/camera/Frame.cpp:35-148). */
#include <cstdio>
#include <cstring>
#include <vector>

#include <CppDebugger.hpp>

#include "Frame.hpp"

using namespace RayTracer;
using namespace CppDebugger::SeverityValues;

Frame::Frame(uint32_t w, uint32_t h) : data(new uint32_t[(size_t) w * h]()), width(w), height(h) {}

Frame::Frame(const Frame& other) : data(new uint32_t[(size_t) other.width * other.height]), width(other.width), height(other.height) {
    std::memcpy(this->data, other.data, sizeof(uint32_t) * (size_t) this->width * this->height);
}

Frame::Frame(Frame&& other) : data(other.data), width(other.width), height(other.height) { other.data = nullptr; }

Frame::~Frame() { delete[] this->data; }

void Frame::to_ppm(const std::string& path) const {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (f == nullptr) { DLOG(fatal, "Could not open output file '" + path + "'"); }
    std::fprintf(f, "P6\n%u %u\n255\n", this->width, this->height);
    std::vector<unsigned char> row((size_t) this->width * 3);
    for (uint32_t y = 0; y < this->height; y++) {
        for (uint32_t x = 0; x < this->width; x++) {
            const uint32_t px = this->data[(size_t) y * this->width + x];
            row[3 * x] = (unsigned char) (px >> 24); row[3 * x + 1] = (unsigned char) (px >> 16); row[3 * x + 2] = (unsigned char) (px >> 8);
        }
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
}
