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

namespace {
    /* CRC-32 (PNG chunks) and Adler-32 (zlib stream), bit by bit / byte by byte: the frame is written once. */
    uint32_t crc32_update(uint32_t crc, const unsigned char* p, size_t n) {
        static uint32_t table[256];
        static bool ready = false;
        if (!ready) {
            for (uint32_t i = 0; i < 256; i++) {
                uint32_t c = i;
                for (int k = 0; k < 8; k++) { c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; }
                table[i] = c;
            }
            ready = true;
        }
        for (size_t i = 0; i < n; i++) { crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8); }
        return crc;
    }
    void put_be32(std::vector<unsigned char>& v, uint32_t x) {
        v.push_back((unsigned char) (x >> 24)); v.push_back((unsigned char) (x >> 16)); v.push_back((unsigned char) (x >> 8)); v.push_back((unsigned char) x);
    }
    void write_chunk(std::FILE* f, const char type[4], const std::vector<unsigned char>& body) {
        std::vector<unsigned char> head;
        put_be32(head, (uint32_t) body.size());
        head.insert(head.end(), type, type + 4);
        uint32_t crc = crc32_update(0xFFFFFFFFu, head.data() + 4, 4);
        crc = crc32_update(crc, body.data(), body.size()) ^ 0xFFFFFFFFu;
        std::vector<unsigned char> tail;
        put_be32(tail, crc);
        std::fwrite(head.data(), 1, head.size(), f);
        std::fwrite(body.data(), 1, body.size(), f);
        std::fwrite(tail.data(), 1, tail.size(), f);
    }
}

void Frame::to_png(const std::string& path) const {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (f == nullptr) { DLOG(fatal, "Could not open output file '" + path + "'"); }
    /* raw scanlines: filter byte 0 + RGBA, alpha 255 as the reference writes it (Frame.cpp:93-96) */
    const size_t stride = 1 + (size_t) this->width * 4;
    std::vector<unsigned char> raw(stride * this->height);
    for (uint32_t y = 0; y < this->height; y++) {
        unsigned char* row = raw.data() + stride * y;
        row[0] = 0;
        for (uint32_t x = 0; x < this->width; x++) {
            const uint32_t px = this->data[(size_t) y * this->width + x];
            row[1 + 4 * x] = (unsigned char) (px >> 24); row[2 + 4 * x] = (unsigned char) (px >> 16); row[3 + 4 * x] = (unsigned char) (px >> 8);
            row[4 + 4 * x] = 255;
        }
    }
    /* zlib stream of stored deflate blocks (at most 65535 bytes each) */
    std::vector<unsigned char> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    size_t pos = 0;
    do {
        const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((unsigned char) (n & 0xFF)); z.push_back((unsigned char) (n >> 8));
        z.push_back((unsigned char) (~n & 0xFF)); z.push_back((unsigned char) ((~n >> 8) & 0xFF));
        for (size_t i = 0; i < n; i++) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
        z.insert(z.end(), raw.begin() + (long) pos, raw.begin() + (long) (pos + n));
        pos += n;
    } while (pos < raw.size());
    put_be32(z, (b << 16) | a);
    static const unsigned char signature[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    std::fwrite(signature, 1, 8, f);
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, this->width); put_be32(ihdr, this->height);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); /* 8-bit RGBA, no interlace */
    write_chunk(f, "IHDR", ihdr);
    write_chunk(f, "IDAT", z);
    write_chunk(f, "IEND", std::vector<unsigned char>());
    std::fclose(f);
}

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
