/* sceneparser/SceneParser.hpp — SceneLang front end: a .scene file -> the entity list main() hands to
 * Renderer::prerender.
 *
 * The reference specifies the language (src/lib/sceneparser/SceneLang.md:1-171, fixture tests/test.scene) but its
 * parser is a stub (SceneParser.cpp); scenes there are code in main (Main.cpp:276-283). This parser implements the
 * specification's core so that the benchmark scenes can be data files:
 *   sections   data { } / entities { } / global { }, any number of times, in order (SceneLang.md:11-23)
 *   data       .obj <id> { inline text }  |  extern .obj <id>: "path";   (SceneLang.md:46-58)
 *   entities   triangle | sphere | object <id> { [<type>] <key>: <expr>;  data[ <key>]: .obj <id>; }  (SceneLang.md:63-83)
 *   expr       C precedence over + - * / % comparisons && || & | ^ << >>, unary - + ! ~, casts, parentheses,
 *              <entity>.<key> / global.<key> references, vec3 as three juxtaposed values (SceneLang.md:90-141)
 *   statements @warning / @error / @ignore (and the fixture's @suppress), #include "file"   (SceneLang.md:29-43,151-155)
 * Reserved keys (the specification's appendix B is empty; these are ECS::create_*'s arguments,
 * reference Triangle.hpp:45, Sphere.hpp:58, Object.hpp:46): triangle p1 p2 p3 color; sphere center radius n_meridians
 * n_parallels color; object center scale data color. Errors are DLOG(fatal, "<file>:<line>: ...").
 */
#ifndef RT3_HOST_SCENEPARSER_SCENE_PARSER_HPP
#define RT3_HOST_SCENEPARSER_SCENE_PARSER_HPP

#include <string>
#include <vector>

#include "entities/RenderEntity.hpp"
#include "tools/Array.hpp"

namespace RayTracer {
    struct ParsedScene {
        Tools::Array<ECS::RenderEntity*> entities; /* in file order; owned by the caller (delete by type, like Main.cpp:286-288) */
        std::vector<std::string> names;            /* entity identifiers, parallel to entities */
        std::vector<std::string> warnings;         /* @warning statements and unused-parameter notes */
    };

    namespace SceneParser {
        /* Parses SceneLang text; relative extern / #include paths are resolved against base_dir. */
        void parse_string(const std::string& text, const std::string& base_dir, ParsedScene& out, const std::string& name = "<string>");
        void parse_file(const std::string& path, ParsedScene& out);
    }
}

#endif
