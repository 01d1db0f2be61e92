"""ORACLE (test infrastructure only) -- the course's text scene format -> FlatScene.

[OWN SPEC] There is NO parser at reference HEAD: `src/main.rs:48` keeps only the comment
`// let scene = parse_file_content(file_lines);` and `gltf::import` is called unconditionally (`main.rs:45`), so
a `.txt` input panics.  The grammar below is recovered from the scene files the reference still ships
(`scenes/practice3_1.txt:1-25`, `practice3_5.txt:1-52`, `practice3_2.txt:26` ROTATION, `practice3_3.txt:47` METALLIC,
`practice3_4.txt:47-48` DIELECTRIC/IOR, `working.txt` TRIANGLE) -- SURVEY.md A.4 -- and the mapping onto the
reference's HEAD data model (`scene.rs:6-39`, `geometry.rs:27-46`) is this repository's own specification,
DESIGN.md section 12.  The product parser is the C++ one behind rt_scene_load_text (csrc/text_loader.cpp); the two
are compared value by value in tests/test_text_scene.py.  Parity status: UNPINNED (nothing to pin against).

Grammar: line oriented, whitespace separated, blank lines ignored, unknown keywords ignored.
  header      DIMENSIONS w h | RAY_DEPTH n | SAMPLES n | BG_COLOR r g b | CAMERA_POSITION x y z |
              CAMERA_RIGHT|CAMERA_UP|CAMERA_FORWARD x y z | CAMERA_FOV_X radians
  primitive   NEW_PRIMITIVE, then any of:  PLANE nx ny nz | ELLIPSOID rx ry rz | BOX sx sy sz |
              TRIANGLE ax ay az bx by bz cx cy cz | POSITION x y z | ROTATION x y z w | COLOR r g b | EMISSION r g b |
              METALLIC | DIELECTRIC | IOR x
Mapping (own spec):
  * fov_y from the aspect ratio: tan(fov_y/2) = tan(fov_x/2) * h / w.
  * ROTATION x y z w -> UnitQuaternion (i, j, k, w), normalised (new_normalize).
  * PLANE normals are normalised; planes go to Scene::infinite_primitives (scene.rs:37), everything else to the BVH.
  * TRIANGLE vertex normals = the face normal normalize((b-a) x (c-a)) at all three vertices.
  * material (scene.rs:6-11): COLOR -> base_color_factor (default 0,0,0 -- pure emitters carry no COLOR);
      default (diffuse)  -> metallic_factor 0, metallic_roughness 1            (through the reference's brdf, rendering.rs:133-155)
      METALLIC           -> metallic_factor 1, metallic_roughness 0.03         (the loader's roughness floor, gltf_to_scene.rs:221)
      DIELECTRIC + IOR   -> mat_kind 1, ior (scene.rs:18): smooth dielectric, see oracle.cpp get_ray_color
  * light iff |EMISSION| > EPS and the primitive is finite (gltf_to_scene.rs:240).
"""
from __future__ import annotations

import math

import numpy as np

from .gltf_ref import FlatScene

SHAPE = {"TRIANGLE": 0, "BOX": 1, "ELLIPSOID": 2, "PLANE": 3}


def parse_text_scene(path: str, width: int = 0, height: int = 0, samples: int = 0) -> FlatScene:
    """width / height / samples > 0 override DIMENSIONS / SAMPLES (the reference CLI passes them, main.rs:39-41)."""
    hdr = {"DIMENSIONS": [0, 0], "RAY_DEPTH": [1], "SAMPLES": [1], "BG_COLOR": [0.0, 0.0, 0.0], "CAMERA_POSITION": [0.0, 0.0, 0.0],
           "CAMERA_RIGHT": [1.0, 0.0, 0.0], "CAMERA_UP": [0.0, 1.0, 0.0], "CAMERA_FORWARD": [0.0, 0.0, -1.0], "CAMERA_FOV_X": [math.pi / 2]}
    prims = []
    cur = None
    with open(path, "r") as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            key, args = tok[0], tok[1:]
            if key == "NEW_PRIMITIVE":
                cur = {"kind": None, "shape": np.zeros(9), "position": np.zeros(3), "rotation": np.array([0.0, 0.0, 0.0, 1.0]), "color": np.zeros(3),
                       "emission": np.zeros(3), "metallic": False, "dielectric": False, "ior": 1.0}
                prims.append(cur)
            elif key in hdr:
                hdr[key] = [float(x) for x in args]
            elif cur is None:
                continue
            elif key in SHAPE:
                cur["kind"] = SHAPE[key]
                vals = [float(x) for x in args]
                cur["shape"][: len(vals)] = vals
            elif key == "POSITION":
                cur["position"] = np.array([float(x) for x in args[:3]])
            elif key == "ROTATION":
                cur["rotation"] = np.array([float(x) for x in args[:4]])
            elif key == "COLOR":
                cur["color"] = np.array([float(x) for x in args[:3]])
            elif key == "EMISSION":
                cur["emission"] = np.array([float(x) for x in args[:3]])
            elif key == "METALLIC":
                cur["metallic"] = True
            elif key == "DIELECTRIC":
                cur["dielectric"] = True
            elif key == "IOR":
                cur["ior"] = float(args[0])
    prims = [p for p in prims if p["kind"] is not None]
    n = len(prims)
    W = int(width) if width > 0 else int(hdr["DIMENSIONS"][0])
    H = int(height) if height > 0 else int(hdr["DIMENSIONS"][1])
    S = int(samples) if samples > 0 else int(hdr["SAMPLES"][0])
    fov_x = float(hdr["CAMERA_FOV_X"][0])
    fov_y = 2.0 * math.atan(math.tan(fov_x * 0.5) * float(H) / float(W))
    fl = FlatScene(width=W, height=H, samples=S, ray_depth=int(hdr["RAY_DEPTH"][0]), bg_color=np.array(hdr["BG_COLOR"][:3]),
                   camera_position=np.array(hdr["CAMERA_POSITION"][:3]), camera_forward=np.array(hdr["CAMERA_FORWARD"][:3]),
                   camera_right=np.array(hdr["CAMERA_RIGHT"][:3]), camera_up=np.array(hdr["CAMERA_UP"][:3]), camera_fov_x=fov_x, camera_fov_y=fov_y)
    fl.tri_v = np.zeros((n, 9)); fl.tri_n = np.zeros((n, 9)); fl.tri_material = np.zeros((n, 5)); fl.tri_emission = np.zeros((n, 3))
    fl.kind = np.zeros(n, dtype=np.int32); fl.position = np.zeros((n, 3)); fl.rotation = np.zeros((n, 4)); fl.ior = np.ones(n)
    fl.mat_kind = np.zeros(n, dtype=np.int32)
    for i, p in enumerate(prims):
        fl.kind[i] = p["kind"]
        shape = p["shape"].copy()
        if p["kind"] == SHAPE["PLANE"]:
            nrm = shape[:3]
            shape[:3] = nrm / math.sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2])
        fl.tri_v[i] = shape
        if p["kind"] == SHAPE["TRIANGLE"]:
            a, b, c = shape[0:3], shape[3:6], shape[6:9]
            e1, e2 = b - a, c - a
            ng = np.array([e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]])
            ng = ng / math.sqrt(ng[0] * ng[0] + ng[1] * ng[1] + ng[2] * ng[2])
            fl.tri_n[i] = np.concatenate([ng, ng, ng])
        q = p["rotation"]
        fl.rotation[i] = q / math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
        fl.position[i] = p["position"]
        fl.tri_material[i, :3] = p["color"]
        fl.tri_material[i, 3] = 1.0 if p["metallic"] else 0.0
        fl.tri_material[i, 4] = 0.03 if p["metallic"] else 1.0
        fl.tri_emission[i] = p["emission"]
        if p["dielectric"]:
            fl.mat_kind[i] = 1
            fl.ior[i] = p["ior"]
    return fl
