/* entities/Object.cpp — Wavefront-OBJ mesh entity.
 *
 * For the v / f triplet subset the reference accepts (reference
 * src/lib/entities/Object.cpp:54-126,131-199) this yields identical arrays:
 * vertex = centre + scale * v (fp32), face indices re-based by the smallest
 * index in the file, normal = normalize(cross(p3 - p1, p2 - p1)),
 * colour = colour * |n . (0,0,-1)|. Unlike the reference's loader it also
 * tolerates comments, blank lines, other record types and `f a/b/c` tokens
 * (the reference aborts on those, Object.cpp:100-102), and reads the file
 * ONCE per entity: create_object parses it to announce the sizes (reference
 * Object.cpp:81-122) and keeps the records for the pre-render that follows,
 * instead of parsing the file a second time there (reference Object.cpp:136-178).
 */
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <CppDebugger.hpp>

#include "Object.hpp"

using namespace RayTracer;
using namespace CppDebugger::SeverityValues;

namespace {
    struct ObjRecord { char kind; float a, b, c; };

    /* Parses "<kind> <n> <n> <n>"; numbers are read with strtof (what operator>> uses), a trailing "/..." is skipped. */
    bool parse_record(const std::string& line, ObjRecord& rec) {
        const char* s = line.c_str();
        while (*s == ' ' || *s == '\t') { s++; }
        if ((*s != 'v' && *s != 'f') || (s[1] != ' ' && s[1] != '\t')) { return false; }
        rec.kind = *s++;
        float* out[3] = { &rec.a, &rec.b, &rec.c };
        for (int i = 0; i < 3; i++) {
            char* end = nullptr;
            *out[i] = std::strtof(s, &end);
            if (end == s) { return false; }
            s = end;
            while (*s != '\0' && *s != ' ' && *s != '\t' && *s != '\r') { s++; }
        }
        return true;
    }

    /* Records parsed by create_object, waiting for the entity's pre-render (entities are plain structs of the reference's layout, so
     * the hand-over lives beside them). An entity that is deleted without ever being pre-rendered leaves its records here until
     * the address is reused by the next create_object, which replaces them. */
    std::mutex g_parsed_mutex;
    std::unordered_map<const void*, std::vector<ObjRecord>> g_parsed;

    std::vector<ObjRecord> read_records(const std::string& path) {
        std::ifstream in(path);
        if (!in.is_open()) { DLOG(fatal, "Could not open file: " + std::string(std::strerror(errno))); }
        std::vector<ObjRecord> records;
        std::string line;
        ObjRecord rec;
        while (std::getline(in, line)) { if (parse_record(line, rec)) { records.push_back(rec); } }
        return records;
    }
}

ECS::Object* ECS::create_object(const std::string& file_path, const glm::vec3& center, float scale, const glm::vec3& color) {
    Object* obj = new Object;
    obj->type = et_object;
    obj->pre_render_mode = eprmf_cpu;
    obj->pre_render_operation = epro_load_object_file;
    obj->pre_render_faces = 0;
    obj->pre_render_vertices = 0;
    obj->file_path = file_path;
    obj->center = center;
    obj->scale = scale;
    obj->color = color;
    try {
        std::vector<ObjRecord> records = read_records(file_path);
        for (const ObjRecord& rec : records) {
            if (rec.kind == 'f') { ++obj->pre_render_faces; } else { ++obj->pre_render_vertices; }
        }
        std::lock_guard<std::mutex> lock(g_parsed_mutex);
        g_parsed[obj] = std::move(records);
    } catch (...) { delete obj; throw; }
    return obj;
}

void ECS::cpu_pre_render_object(Tools::Array<GFace>& faces_buffer, Tools::Array<glm::vec4>& vertex_buffer, Object* obj) {
    std::vector<ObjRecord> records;
    bool parsed = false;
    {
        std::lock_guard<std::mutex> lock(g_parsed_mutex);
        std::unordered_map<const void*, std::vector<ObjRecord>>::iterator it = g_parsed.find(obj);
        if (it != g_parsed.end()) { records = std::move(it->second); g_parsed.erase(it); parsed = true; }
    }
    if (!parsed) { records = read_records(obj->file_path); } /* a second pre-render of the same entity */
    faces_buffer.resize(obj->pre_render_faces);
    vertex_buffer.resize(obj->pre_render_vertices);
    size_t n_faces = 0, n_vertices = 0;
    uint32_t lowest = std::numeric_limits<uint32_t>::max();
    for (const ObjRecord& rec : records) {
        if (rec.kind == 'v') {
            if (n_vertices >= vertex_buffer.size()) { DLOG(fatal, "Object file changed since create_object: too many vertices"); }
            vertex_buffer[n_vertices++] = glm::vec4(obj->center + obj->scale * glm::vec3(rec.a, rec.b, rec.c), 0.0f);
        } else {
            if (n_faces >= faces_buffer.size()) { DLOG(fatal, "Object file changed since create_object: too many faces"); }
            GFace& f = faces_buffer[n_faces++];
            f.v1 = (uint32_t) rec.a; f.v2 = (uint32_t) rec.b; f.v3 = (uint32_t) rec.c;
            lowest = std::min(lowest, std::min(f.v1, std::min(f.v2, f.v3)));
        }
    }
    for (size_t i = 0; i < n_faces; i++) {
        GFace& f = faces_buffer[i];
        f.v1 -= lowest; f.v2 -= lowest; f.v3 -= lowest;
        if (f.v1 >= n_vertices || f.v2 >= n_vertices || f.v3 >= n_vertices) { DLOG(fatal, "Face " + std::to_string(i) + " references a vertex that does not exist"); }
        const glm::vec4 &a = vertex_buffer[f.v1], &b = vertex_buffer[f.v2], &c = vertex_buffer[f.v3];
        const glm::vec3 p1(a.x, a.y, a.z), p2(b.x, b.y, b.z), p3(c.x, c.y, c.z);
        f.normal = glm::normalize(glm::cross(p3 - p1, p2 - p1));
        f.color = obj->color * glm::abs(glm::dot(f.normal, glm::vec3(0.0f, 0.0f, -1.0f)));
    }
}
