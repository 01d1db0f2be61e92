"""ORACLE (test infrastructure only) -- glTF 2.0 -> flat scene, a numpy restatement of
/root/reference/src/gltf_to_scene.rs:21-256 (convert_gltf_to_scene, read_primitives).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product loader is the C++ one behind rt_scene_load_gltf (csrc/gltf_loader.cpp).

Parity status: UNPINNED -- the reference ships no fixtures for its loader; this restates the cited
lines plus the published behaviour of the `gltf` crate 1.x (Cargo.toml:22-24, exact version unpinned: no
Cargo.lock) that those lines call:
  * JSON numbers are deserialised as f32 (serde, gltf-json).
  * Node::transform().matrix() of a TRS node composes T*R*S in f32 (gltf crate `math.rs`), the reference
    widens each entry to f64 afterwards (gltf_to_scene.rs:107-122).
  * material getters fall back to the glTF defaults: baseColorFactor (1,1,1,1), metallicFactor 1,
    roughnessFactor 1, emissiveFactor 0, emissive_strength None (gltf_to_scene.rs:215-231).
Quirks restated on purpose:
  * every node of `gltf.nodes()` is visited with the identity parent transform, and children are then
    visited again through the recursion (gltf_to_scene.rs:42-52, 245-255);
  * only the first primitive of a mesh is read (gltf_to_scene.rs:148);
  * normals are rotated by the accumulated unit quaternion local*parent, scale ignored (:112-117, 192-194);
  * camera basis = columns of the node matrix, fov_x = aspect*yfov computed in f32 (:134-143);
  * roughness clamped to >= 0.03 (:221); emission = factor*strength (:223-231); light iff |E| > 1e-5 (:240);
  * ray_depth 6, background black (:65, :73).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np

EPS = 1e-5  # geometry.rs:49
F32 = np.float32

_COMPONENT_DTYPE = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_TYPE_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


@dataclass
class FlatScene:
    """What Scene (scene.rs:22-39) holds after the loader ran, before the BVH build, in load order."""
    width: int
    height: int
    samples: int
    ray_depth: int = 6
    bg_color: np.ndarray = field(default_factory=lambda: np.zeros(3))
    camera_position: np.ndarray = field(default_factory=lambda: np.zeros(3))
    camera_forward: np.ndarray = field(default_factory=lambda: np.zeros(3))
    camera_right: np.ndarray = field(default_factory=lambda: np.zeros(3))
    camera_up: np.ndarray = field(default_factory=lambda: np.zeros(3))
    camera_fov_x: float = 0.0
    camera_fov_y: float = 0.0
    tri_v: np.ndarray = field(default_factory=lambda: np.zeros((0, 9)))       # a, b, c world space, f64
    tri_n: np.ndarray = field(default_factory=lambda: np.zeros((0, 9)))       # a_norm, b_norm, c_norm
    tri_material: np.ndarray = field(default_factory=lambda: np.zeros((0, 5)))  # base rgb, metallic, roughness
    tri_emission: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    # general primitives (text scenes, oracle/text_ref.py); None = triangles with identity transforms (all the glTF loader emits).
    # For kind != 0 the first three entries of a tri_v row are the box half sizes / ellipsoid radii / plane normal.
    kind: np.ndarray | None = None         # (n,) int32: 0 triangle, 1 box, 2 ellipsoid, 3 plane   (geometry.rs:27-39 + own spec)
    position: np.ndarray | None = None     # (n, 3)  Object3D.position (geometry.rs:44)
    rotation: np.ndarray | None = None     # (n, 4)  Object3D.rotation as (i, j, k, w)  (geometry.rs:45)
    ior: np.ndarray | None = None          # (n,)    Primitive.ior (scene.rs:18)
    mat_kind: np.ndarray | None = None     # (n,) int32: 0 metallic-roughness (scene.rs:6-11), 1 dielectric (own spec)

    @property
    def n_tris(self) -> int:
        return int(self.tri_v.shape[0])

    @property
    def light_ids(self) -> np.ndarray:
        # gltf_to_scene.rs:240  `if emission.norm() > EPS`; planes (infinite_primitives) are never lights
        lit = np.linalg.norm(self.tri_emission, axis=1) > EPS
        if self.kind is not None:
            lit &= np.asarray(self.kind) != 3
        return np.nonzero(lit)[0].astype(np.int32)


def _quat_mul(a, b):
    """Hamilton product, (w, x, y, z) order -- nalgebra Quaternion * Quaternion (gltf_to_scene.rs:112-117)."""
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ], dtype=np.float64)


def _quat_rotate(q, v):
    """UnitQuaternion::transform_vector (nalgebra): t = (q.v x v) * 2; (t*w + q.v x t) + v  (rows of v)."""
    w, qv = q[0], q[1:4]
    t = np.cross(np.broadcast_to(qv, v.shape), v) * 2.0
    return (t * w + np.cross(np.broadcast_to(qv, v.shape), t)) + v


def _trs_matrix_f32(t, r, s):
    """gltf crate Transform::Decomposed -> matrix(): T * R * S evaluated in f32 (column-major result).

    R from the quaternion (x, y, z, w) with the doubled-products form; T and S only contribute exact zeros and
    ones to the products, so M[:, j] = R[:, j] * s[j] (one f32 rounding) and M[:, 3] = t.
    Returned as a row-major 4x4 f64 array == the Matrix4<f64> the reference builds at :119-122.
    """
    x, y, z, w = (F32(r[0]), F32(r[1]), F32(r[2]), F32(r[3]))
    x2, y2, z2 = x + x, y + y, z + z
    xx2, xy2, xz2 = x2 * x, x2 * y, x2 * z
    yy2, yz2, zz2 = y2 * y, y2 * z, z2 * z
    sx2, sy2, sz2 = x2 * w, y2 * w, z2 * w
    one = F32(1.0)
    col0 = np.array([one - yy2 - zz2, xy2 + sz2, xz2 - sy2], dtype=F32)
    col1 = np.array([xy2 - sz2, one - xx2 - zz2, yz2 + sx2], dtype=F32)
    col2 = np.array([xz2 + sy2, yz2 - sx2, one - xx2 - yy2], dtype=F32)
    m = np.zeros((4, 4), dtype=F32)
    m[:3, 0] = col0 * F32(s[0])
    m[:3, 1] = col1 * F32(s[1])
    m[:3, 2] = col2 * F32(s[2])
    m[:3, 3] = np.array(t, dtype=F32)
    m[3, 3] = one
    return m.astype(np.float64)


def _node_transform(node):
    """(matrix f64 4x4 row-major, rotation (x,y,z,w) as f32 values) of a node, like node.transform()."""
    if "matrix" in node:
        m = np.array(node["matrix"], dtype=F32).reshape(4, 4).T.astype(np.float64)
        # gltf crate 1.x `Transform::Matrix { matrix }.decomposed()` (scene/mod.rs; restated from its published source -- the crate is
        # not vendored and there is no Cargo.lock): ALL in f32.  i = upper 3x3 by columns; sx = |i.x|, sy = |i.y|,
        # sz = signum(det i) * |i.z| (a mirror goes to the z scale only); columns divided by their scale (multiply by 1/s);
        # rotation = Quaternion::from_matrix(i) (the cgmath branch order: trace >= 0, then the largest diagonal element).
        c = np.array(node["matrix"], dtype=F32).reshape(4, 4)[:3, :3].copy()          # c[k] = column k (x, y, z components)
        mag = lambda v: F32(np.sqrt(F32(F32(F32(v[0] * v[0]) + F32(v[1] * v[1])) + F32(v[2] * v[2]))))   # noqa: E731
        x, y, z = c[0], c[1], c[2]
        det = F32(F32(F32(x[0] * F32(F32(y[1] * z[2]) - F32(z[1] * y[2]))) - F32(y[0] * F32(F32(x[1] * z[2]) - F32(z[1] * x[2])))) + F32(z[0] * F32(F32(x[1] * y[2]) - F32(y[1] * x[2]))))
        sx, sy = mag(x), mag(y)
        sz = F32(F32(np.copysign(F32(1.0), det)) * mag(z))                               # f32::signum: +-1 (also for +-0)
        x = np.array([F32(v * F32(F32(1.0) / sx)) for v in x], dtype=F32)
        y = np.array([F32(v * F32(F32(1.0) / sy)) for v in y], dtype=F32)
        z = np.array([F32(v * F32(F32(1.0) / sz)) for v in z], dtype=F32)
        trace = F32(F32(x[0] + y[1]) + z[2])
        h = F32(0.5)
        if trace >= 0:
            s_ = F32(np.sqrt(F32(F32(1.0) + trace)))
            w = F32(h * s_); s_ = F32(h / s_)
            q = [F32(F32(y[2] - z[1]) * s_), F32(F32(z[0] - x[2]) * s_), F32(F32(x[1] - y[0]) * s_), w]
        elif x[0] > y[1] and x[0] > z[2]:
            s_ = F32(np.sqrt(F32(F32(F32(x[0] - y[1]) - z[2]) + F32(1.0))))
            qx = F32(h * s_); s_ = F32(h / s_)
            q = [qx, F32(F32(y[0] + x[1]) * s_), F32(F32(x[2] + z[0]) * s_), F32(F32(y[2] - z[1]) * s_)]
        elif y[1] > z[2]:
            s_ = F32(np.sqrt(F32(F32(F32(y[1] - x[0]) - z[2]) + F32(1.0))))
            qy = F32(h * s_); s_ = F32(h / s_)
            q = [F32(F32(y[0] + x[1]) * s_), qy, F32(F32(z[1] + y[2]) * s_), F32(F32(z[0] - x[2]) * s_)]
        else:
            s_ = F32(np.sqrt(F32(F32(F32(z[2] - x[0]) - y[1]) + F32(1.0))))
            qz = F32(h * s_); s_ = F32(h / s_)
            q = [F32(F32(x[2] + z[0]) * s_), F32(F32(z[1] + y[2]) * s_), qz, F32(F32(x[1] - y[0]) * s_)]
        return m, [float(c_) for c_ in q]
    t = node.get("translation", [0.0, 0.0, 0.0])
    r = node.get("rotation", [0.0, 0.0, 0.0, 1.0])
    s = node.get("scale", [1.0, 1.0, 1.0])
    return _trs_matrix_f32(t, r, s), [float(F32(c)) for c in r]


class _Doc:
    def __init__(self, path):
        self.dir = os.path.dirname(os.path.abspath(path))
        with open(path, "r") as f:
            self.j = json.load(f)
        self.buffers = []
        for b in self.j.get("buffers", []):
            uri = b["uri"]
            if uri.startswith("data:"):                      # gltf::import decodes base64 data URIs
                import base64
                self.buffers.append(base64.b64decode(uri.split(",", 1)[1]))
                continue
            with open(os.path.join(self.dir, uri), "rb") as f:
                self.buffers.append(f.read())

    def accessor(self, idx):
        acc = self.j["accessors"][idx]
        bv = self.j["bufferViews"][acc["bufferView"]]
        dt = np.dtype(_COMPONENT_DTYPE[acc["componentType"]])
        ncomp = _TYPE_NCOMP[acc["type"]]
        off = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * ncomp
        raw = np.frombuffer(self.buffers[bv["buffer"]], dtype=np.uint8)
        n = acc["count"]
        idxs = off + stride * np.arange(n)[:, None] + np.arange(dt.itemsize * ncomp)[None, :]
        return raw[idxs].copy().view(dt).reshape(n, ncomp)


def convert_gltf_to_scene(path: str, width: int, height: int, samples: int) -> FlatScene:
    """main.rs:45-47: gltf::import(path) then convert_gltf_to_scene(&gltf, &buffers, width, height, samples)."""
    doc = _Doc(path)
    j = doc.j
    sc = FlatScene(width=width, height=height, samples=samples)
    tv, tn, tm, te = [], [], [], []

    def material_of(prim):
        mat = j["materials"][prim["material"]] if "material" in prim else {}
        pbr = mat.get("pbrMetallicRoughness", {})
        base = [float(F32(c)) for c in pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0])[:3]]
        metallic = float(F32(pbr.get("metallicFactor", 1.0)))
        rough = max(float(F32(pbr.get("roughnessFactor", 1.0))), 0.03)
        ef = [float(F32(c)) for c in mat.get("emissiveFactor", [0.0, 0.0, 0.0])]
        st = mat.get("extensions", {}).get("KHR_materials_emissive_strength", None)
        # gltf crate: emissive_strength() is Some(..) iff the extension object exists; emissiveStrength defaults to 1.0
        strength = float(F32(st.get("emissiveStrength", 1.0))) if st is not None else 1.0
        return base + [metallic, rough], [e * strength for e in ef]

    def read_primitives(node_idx, transformation, rotation):
        node = j["nodes"][node_idx]
        local_m, r = _node_transform(node)
        # Quaternion::new(w, i, j, k) * rotation, then normalised (UnitQuaternion::from_quaternion)  :112-117
        q = _quat_mul(np.array([r[3], r[0], r[1], r[2]], dtype=np.float64), rotation)
        # Unit::new_normalize: q / sqrt(i^2 + j^2 + k^2 + w^2) (nalgebra stores quaternion coords as [i, j, k, w])
        current_rotation = q / np.sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0])
        m = transformation @ local_m                                                    # :127
        if "camera" in node:
            cam = j["cameras"][node["camera"]]
            if cam.get("type") != "perspective":
                raise NotImplementedError("non-perspective camera (todo!() at gltf_to_scene.rs:131-133)")
            persp = cam["perspective"]
            yfov = F32(persp["yfov"])
            aspect = F32(persp["aspectRatio"]) if "aspectRatio" in persp else F32(1.0)
            sc.camera_fov_y = float(yfov)
            sc.camera_fov_x = float(F32(aspect * yfov))                                 # f32 multiply, then widened
            pos = m @ np.array([0.0, 0.0, 0.0, 1.0])
            sc.camera_position = pos[:3] / pos[3]
            sc.camera_forward = (-m @ np.array([0.0, 0.0, 1.0, 0.0]))[:3]               # :136,142
            sc.camera_up = (m @ np.array([0.0, 1.0, 0.0, 0.0]))[:3]                      # :137,140
            sc.camera_right = (m @ np.array([1.0, 0.0, 0.0, 0.0]))[:3]                   # :138,141
        if "mesh" in node:
            prim = j["meshes"][node["mesh"]]["primitives"][0]                            # :148
            if "indices" not in prim:
                raise NotImplementedError("non-indexed mesh (todo!() at gltf_to_scene.rs:151-153)")
            idx = doc.accessor(prim["indices"]).reshape(-1).astype(np.int64)
            pos = doc.accessor(prim["attributes"]["POSITION"]).astype(np.float64)
            ntri = idx.shape[0] // 3
            idx = idx[: ntri * 3].reshape(ntri, 3)
            # m * (x, y, z, 1) accumulated column by column (nalgebra gemv), then divided by w   :172-180
            p4 = (m[:, 0][None, :] * pos[:, 0:1])
            p4 = p4 + m[:, 1][None, :] * pos[:, 1:2]
            p4 = p4 + m[:, 2][None, :] * pos[:, 2:3]
            p4 = p4 + m[:, 3][None, :]
            wpos = p4[:, :3] / p4[:, 3:4]
            a, b, c = wpos[idx[:, 0]], wpos[idx[:, 1]], wpos[idx[:, 2]]
            if "NORMAL" in prim["attributes"]:
                nrm = doc.accessor(prim["attributes"]["NORMAL"]).astype(np.float64)
                wn = _quat_rotate(current_rotation, nrm)                                 # :192-194
                an, bn, cn = wn[idx[:, 0]], wn[idx[:, 1]], wn[idx[:, 2]]
            else:
                dn = np.cross(b - a, c - a)
                dn = dn / np.linalg.norm(dn, axis=1, keepdims=True)                      # :184
                an = bn = cn = dn
            mat, emi = material_of(prim)
            tv.append(np.concatenate([a, b, c], axis=1))
            tn.append(np.concatenate([an, bn, cn], axis=1))
            tm.append(np.tile(np.array(mat, dtype=np.float64), (ntri, 1)))
            te.append(np.tile(np.array(emi, dtype=np.float64), (ntri, 1)))
        for child in node.get("children", []):                                           # :245-255
            read_primitives(child, m, current_rotation)

    ident_q = np.array([1.0, 0.0, 0.0, 0.0])
    for node_idx in range(len(j.get("nodes", []))):                                      # :42-52 (all nodes)
        read_primitives(node_idx, np.eye(4), ident_q)

    if tv:
        sc.tri_v = np.ascontiguousarray(np.concatenate(tv, axis=0))
        sc.tri_n = np.ascontiguousarray(np.concatenate(tn, axis=0))
        sc.tri_material = np.ascontiguousarray(np.concatenate(tm, axis=0))
        sc.tri_emission = np.ascontiguousarray(np.concatenate(te, axis=0))
    return sc
