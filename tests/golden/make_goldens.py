"""Regenerates tests/golden/* from the COMPILED REFERENCE (oracle/_ref/libref_seq.so).

Run in the build container, where /root/reference exists:
    make -C oracle ref port && python tests/golden/make_goldens.py

For each scene the reference's own Sequential renderer (driven exactly as
reference src/Main.cpp:266-288) produces the packed frame; the flattened
GFace/vec4 arrays it rendered from are exported next to it, so GPU-box tests can
feed the identical scene to the CUDA path without /root/reference. The id / t
hashes come from the restatement (oracle/rt3_oracle.c) and are only recorded
when its image equals the reference's bit for bit.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oraclelib as ol  # noqa: E402
from rt3_b200 import abi  # noqa: E402

CASES = {
    # name: (builder, W, H, store_scene, store_frame)   -- scenes of reference src/Main.cpp:278-283
    "default_400x225": ("default", 400, 225, True, True),
    "default_800x600": ("default", 800, 600, False, False),
    "triangle_400x225": ("triangle", 400, 225, True, True),
    "sphere8_400x225": ("sphere8", 400, 225, True, True),
    "sphere225_64x36": ("sphere225", 64, 36, False, True),
}


def build(kind):
    s = ol.RefScene()
    if kind == "default":
        s.add_object(ol.REF_TEDDY, (0, 0, -3), 1.0 / 17.0, (1, 0, 0))
        s.add_sphere((-2, 0, -5), 1.0, 8, 8, (0, 0, 1))
    elif kind == "triangle":
        s.add_triangle((1, 0, -3), (-1, 0, -3), (0, 1, -3), (1, 0, 0))
    elif kind == "sphere8":
        s.add_sphere((0, 0, -3), 1.0, 8, 8, (1, 0, 0))
    elif kind == "sphere225":
        s.add_sphere((0, 0, -3), 1.0, 225, 225, (1, 0, 0))
    s.prerender()
    return s


def main():
    meta = {}
    for name, (kind, w, h, store_scene, store_frame) in CASES.items():
        rs = build(kind)
        scene = rs.export()
        frame, seconds = rs.render(w, h)
        cam = abi.reference_camera(w, h)
        oframe, oprim, oent, ot = ol.oracle_reference(scene, cam, w, h)
        mism = int((frame[:h - 1] != oframe[:h - 1]).sum())
        assert mism == 0, f"{name}: restatement differs from the compiled reference in {mism} pixels"
        ent_ids, ent_counts = np.unique(oent[:h - 1], return_counts=True)
        meta[name] = {
            "scene": kind, "width": w, "height": h,
            "n_faces": int(scene.n_faces), "n_vertices": int(len(scene.vertices)),
            "reference_render_seconds_1core": round(seconds, 3),
            "frame_fnv64_rows_0_to_Hm2": f"{ol.fnv64(frame[:h - 1]):016x}",
            "pixel_0": f"{int(frame[0, 0]):08x}", "pixel_centre": f"{int(frame[h // 2, w // 2]):08x}",
            "restatement_prim_fnv64": f"{ol.fnv64(oprim[:h - 1]):016x}",
            "restatement_entity_fnv64": f"{ol.fnv64(oent[:h - 1]):016x}",
            "restatement_t_bits_fnv64": f"{ol.fnv64(ot[:h - 1].view(np.uint32)):016x}",
            "entity_histogram": {f"{int(e):x}": int(c) for e, c in zip(ent_ids, ent_counts)},
        }
        arrays = {}
        if store_scene:
            arrays.update(faces=scene.faces.view(np.uint8).reshape(-1, 48), vertices=scene.vertices.view(np.uint8).reshape(-1, 16),
                          face_entity=scene.face_entity)
        if store_frame:
            arrays.update(frame=frame[:h - 1], prim=oprim[:h - 1], t_bits=ot[:h - 1].view(np.uint32))
        if arrays:
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, meta[name]["frame_fnv64_rows_0_to_Hm2"], f"{seconds:.2f}s")
        rs.close()
    with open(os.path.join(HERE, "goldens.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
