import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/profiles")
import importlib.util
spec = importlib.util.spec_from_file_location("cfg", "/root/repo/profiles/configs.py")
import numpy as np
import rt3_b200
from rt3_b200 import abi, scenes
import hostlib
hs = hostlib.HostScene(); hs.add_sphere((0, 0, -3), 1.0, 225, 225, (0.8, 0.3, 0.3)); mesh = hs.flatten()
mats = np.zeros(3, abi.MATERIAL_DTYPE); mats["kind"] = [0, 1, 2]; mats["albedo"] = [(0.8, 0.8, 0.0), (0.8, 0.6, 0.2), (1, 1, 1)]; mats["fuzz"] = [0, 0.1, 0]; mats["ior"] = [1, 1, 1.5]
spheres = np.array([(0, -101, -3, 100), (2.2, 0, -3, 1), (-2.2, 0, -3, 1)], np.float32)
scene = abi.SceneArrays(faces=mesh.faces, vertices=mesh.vertices, face_entity=mesh.face_entity, spheres=spheres, sphere_material=np.arange(3, dtype=np.uint32), sphere_entity=np.arange(1, 4, dtype=np.uint32), materials=mats)
w, h = 1920, 1080
ctx = abi.Context(0); ctx.upload(scene)
p = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=32, max_depth=50, seed=1, flags=abi.FLAG_BVH)
for _ in range(2):
    ctx.render(abi.reference_camera(w, h), p); st = ctx.stats()
    print("ms", st.trace_kernel_ms, "rays", st.rays, "visits/ray", st.accel_node_visits / st.rays, "tests/ray", st.accel_prim_tests / st.rays)
