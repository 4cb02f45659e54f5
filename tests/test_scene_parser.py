"""SceneLang front end (raytracer-3_b200/host/sceneparser): a .scene text yields the entity list that the same
ECS::create_* calls yield (reference src/lib/sceneparser/SceneLang.md; the reference's own parser is a stub,
so the checker is the API-built scene, flattened by the same host code). CPU only."""
import numpy as np
import pytest

import hostlib
from rt3_b200 import abi

SCENE = r'''
/* data first: one inline mesh, one external */
data {
    .obj quad {
        v 0 0 -3
        v 1 0 -3
        v 1 1 -3
        v 0 1 -3
        f 1 2 3
        f 1 3 4
    }
    @suppress unused-data
    extern .obj outside: "meshes/tri.obj";
}
global {
    float lift: 0.25;
    vec3 red: 1.0 0.0 0.0;
    uint rings: 4 + 2 * 2;          // 8
}
entities {
    triangle tri_1 {                 // the fixture's untyped form
        p1: -1.0 0.0 -3.0;
        p2: 1.0 0.0 -3.0;
        p3: 0.0 1.0 -3.0;
        color: global.red;
    }
    sphere ball {
        vec3 center: 0.0 global.lift -3.0;
        float radius: (float) 3 / 2.0 - 0.5;       /* 1.0 */
        uint n_meridians: global.rings;
        uint n_parallels: (uint) 8.9;
        vec3 color: tri_1.color * 0.5 + vec3(0.0, 0.5, 0.0);
        float unused_note: 1e3;
    }
    object flat_quad {
        center: ball.center - (0.0 0.25 0.0);
        scale: 2.0;
        data: .obj quad;
        color: 0.0 0.0 1.0;
    }
}
entities {
    @warning "second entities section"
    object far_tri { center: 0.0 0.0 -6.0; scale: 1.0; data mesh: .obj outside; color: 1.0 1.0 1.0; }
}
'''


def same(a, b):
    assert a.n_faces == b.n_faces and len(a.vertices) == len(b.vertices)
    assert np.array_equal(a.faces.view(np.uint8), b.faces.view(np.uint8))
    assert np.array_equal(a.vertices.view(np.uint8), b.vertices.view(np.uint8))
    assert np.array_equal(a.face_entity, b.face_entity)


def test_scene_text_equals_api_calls(built, tmp_path):
    (tmp_path / "meshes").mkdir()
    (tmp_path / "meshes" / "tri.obj").write_text("v -1 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    (tmp_path / "quad.obj").write_text("v 0 0 -3\nv 1 0 -3\nv 1 1 -3\nv 0 1 -3\nf 1 2 3\nf 1 3 4\n")
    parsed = hostlib.HostScene()
    n, warnings = parsed.add_scene_text(SCENE, tmp_path)
    assert n == 4 and len(warnings) == 1 and "second entities section" in warnings[0]
    api = hostlib.HostScene()
    api.add_triangle((-1, 0, -3), (1, 0, -3), (0, 1, -3), (1, 0, 0))
    api.add_sphere((0, 0.25, -3), 1.0, 8, 8, (0.5, 0.5, 0.0))
    api.add_object(str(tmp_path / "quad.obj"), (0, 0, -3), 2.0, (0, 0, 1))
    api.add_object(str(tmp_path / "meshes" / "tri.obj"), (0, 0, -6), 1.0, (1, 1, 1))
    same(parsed.flatten(), api.flatten())


def test_include_and_expressions(built, tmp_path):
    (tmp_path / "common.scene").write_text("global { int n: 7 % 4 << 1; bool big: 3 > 2 && !(1 == 2); vec3 up: 0.0 1.0 0.0; }\n")
    text = '''#include "common.scene"
    entities { sphere s { center: -global.up * (float) global.n; radius: 0.5; n_meridians: global.n; n_parallels: global.n - (int) global.big * 3;
                          color: 0.25 0.5 1.0; } }'''
    parsed = hostlib.HostScene()
    assert parsed.add_scene_text(text, tmp_path)[0] == 1
    api = hostlib.HostScene()
    api.add_sphere((0, -6, 0), 0.5, 6, 3, (0.25, 0.5, 1.0))     # n = (7 % 4) << 1 = 6; parallels = 6 - 1 * 3
    same(parsed.flatten(), api.flatten())


def test_builtin_functions(built):
    """Scalar and vector built-ins (SceneLang.md section 3 names them, its appendix C lists none): evaluated in float like the API caller would."""
    text = '''global { vec3 a: 1.0 2.0 2.0; vec3 b: 0.0 0.0 -1.0; float third: 1.0 / length(global.a); }
    entities { sphere s { center: cross(global.a, global.b) + normalize(global.a) * 3.0;
                          radius: pow(2.0, -1.0) + dot(global.a, global.b) * 0.0 + tan(0.0);
                          n_meridians: (uint) ceil(5.5); n_parallels: (uint) floor(4.9) + (uint) (radians(180.0) > 3.14);
                          color: global.third global.third max(0.25, min(global.third, 1.0)); } }'''
    parsed = hostlib.HostScene()
    assert parsed.add_scene_text(text)[0] == 1
    f = np.float32
    third = f(1) / f(3)
    n = np.array([1, 2, 2], f) * (f(1) / np.sqrt(f(9)))
    centre = np.array([-2, 1, 0], f) + n * f(3)          # a x b = (2*-1 - 0*2, 2*0 - (-1)*1, 0) = (-2, 1, 0)
    api = hostlib.HostScene()
    api.add_sphere(tuple(float(x) for x in centre), 0.5, 6, 5, (float(third), float(third), float(third)))
    same(parsed.flatten(), api.flatten())
    for bad, message in (("global { float x: dot(1.0, 2.0); }", "needs vec3"), ("global { float x: length(); }", "takes 1 argument"),
                         ("global { float x: frobnicate(1.0); }", "unknown function")):
        with pytest.raises(hostlib.HostError, match=message):
            hostlib.HostScene().add_scene_text(bad)


@pytest.mark.parametrize("text,message", [
    ("lights { }", "unknown section 'lights'"),
    ("entities { cube c { } }", "unknown entity type 'cube'"),
    ("entities { sphere s { center: 0.0 0.0 0.0; radius: 1.0; n_meridians: 8; color: 1.0 1.0 1.0; } }", "missing parameter 'n_parallels'"),
    ("entities { sphere s { center: 0.0 0.0 0.0; radius: 1.0; n_meridians: 8; n_parallels: 2; color: 1.0 1.0 1.0; } }", "n_parallels >= 3"),
    ("entities { triangle t { p1: 0.0 0.0 0.0; p2: other.p1; p3: 0.0 1.0 0.0; color: 1.0 1.0 1.0; } }", "unknown entity 'other'"),
    ("entities { object o { center: 0.0 0.0 0.0; scale: 1.0; data: .obj nothing; color: 1.0 1.0 1.0; } }", "undefined data 'nothing'"),
    ("data { .obj a { v 0 0 0 } .obj a { v 1 1 1 } }", "defined twice"),
    ("@error \"stop here\"\nentities { }", "stop here"),
    ("entities { triangle t { p1: 0.0 0.0 0.0; p2: 1.0 0.0 0.0; p3: 0.0 1.0 0.0; color: 1.0 1.0 1.0 }", "expected ';'"),
    ("global { float x: 1.0 / ; }", "unexpected ';'"),
    ("global { int x: 1 / 0; }", "division by zero"),
    ("entities { sphere global { } }", "'global' cannot be used"),
    ("data { extern .obj gone: \"no/such/file.obj\"; } entities { object o { center: 0.0 0.0 0.0; scale: 1.0; data: .obj gone; color: 1.0 1.0 1.0; } }",
     "Could not open file"),
])
def test_errors_name_the_line(built, text, message):
    hs = hostlib.HostScene()
    with pytest.raises(hostlib.HostError, match=message):
        hs.add_scene_text(text)
    assert hs.flatten().n_faces == 0, "a failed parse must not leave entities behind"
