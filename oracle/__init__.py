"""ORACLE -- test infrastructure only (see oracle/oracle.cpp header).  ctypes front-end to liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
Parity status: UNPINNED (no reference golden vectors exist for this path).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .gltf_ref import FlatScene, convert_gltf_to_scene  # noqa: F401
from .text_ref import parse_text_scene  # noqa: F401

SHAPE_TRIANGLE, SHAPE_BOX, SHAPE_ELLIPSOID, SHAPE_PLANE = range(4)
MAT_PBR, MAT_DIELECTRIC = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)


class OrSceneDesc(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("samples", C.c_int32), ("ray_depth", C.c_int32),
        ("bg_color", C.c_double * 3), ("camera_position", C.c_double * 3), ("camera_forward", C.c_double * 3),
        ("camera_right", C.c_double * 3), ("camera_up", C.c_double * 3),
        ("camera_fov_x", C.c_double), ("camera_fov_y", C.c_double),
        ("n_tris", C.c_int32), ("_pad", C.c_int32),
        ("tri_v", _dp), ("tri_n", _dp), ("tri_material", _dp), ("tri_emission", _dp),
        ("kind", C.POINTER(C.c_int32)), ("position", _dp), ("rotation", _dp), ("ior", _dp), ("mat_kind", C.POINTER(C.c_int32)),
        ("max_attempts", C.c_int32), ("_pad2", C.c_int32),
    ]


class OrInfo(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_nodes", "n_leaves", "depth", "n_lights", "n_light_nodes", "validate_failures")]


class OrStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("node_tests", "tri_tests", "segments", "vertices", "attempts", "light_node_tests",
                                          "light_tri_tests", "vndf_assert_fail", "nan_pixels", "samples")] + [("seconds", C.c_double), ("attempt_cap_hits", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force: bool = False) -> str:
    """Compile oracle/oracle.cpp -> oracle/_build/liboracle.so (g++, generic x86-64)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.or_scene_create.restype = C.c_void_p
        _lib.or_scene_create.argtypes = [C.POINTER(OrSceneDesc)]
        _lib.or_scene_destroy.argtypes = [C.c_void_p]
        _lib.or_scene_info.argtypes = [C.c_void_p, C.POINTER(OrInfo)]
        _lib.or_render.restype = C.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class OracleScene:
    """Scene (scene.rs:22-39) built from a FlatScene: both BVHs are built by the restated bvh.rs builder."""

    def __init__(self, flat: FlatScene, max_attempts: int = 0):
        """max_attempts = 0: the reference's unbounded rejection loop; > 0: the path ends after that many failed attempts."""
        self.flat = flat
        L = lib()
        d = OrSceneDesc()
        d.width, d.height, d.samples, d.ray_depth = flat.width, flat.height, flat.samples, flat.ray_depth
        for name in ("bg_color", "camera_position", "camera_forward", "camera_right", "camera_up"):
            getattr(d, name)[:] = [float(x) for x in getattr(flat, name)]
        d.camera_fov_x, d.camera_fov_y = flat.camera_fov_x, flat.camera_fov_y
        d.n_tris = flat.n_tris
        self._keep = [_d(flat.tri_v), _d(flat.tri_n), _d(flat.tri_material), _d(flat.tri_emission)]
        d.tri_v, d.tri_n, d.tri_material, d.tri_emission = (_p(a) for a in self._keep)
        if flat.kind is not None:
            n = flat.n_tris
            ex = [np.ascontiguousarray(flat.kind, dtype=np.int32), _d(flat.position if flat.position is not None else np.zeros((n, 3))),
                  _d(flat.rotation if flat.rotation is not None else np.tile([0.0, 0.0, 0.0, 1.0], (n, 1))),
                  _d(flat.ior if flat.ior is not None else np.ones(n)),
                  np.ascontiguousarray(flat.mat_kind if flat.mat_kind is not None else np.zeros(n), dtype=np.int32)]
            self._keep += ex
            d.kind, d.position, d.rotation, d.ior, d.mat_kind = _p(ex[0], C.c_int32), _p(ex[1]), _p(ex[2]), _p(ex[3]), _p(ex[4], C.c_int32)
        d.max_attempts = int(max_attempts)
        self._h = C.c_void_p(L.or_scene_create(C.byref(d)))
        self.width, self.height, self.samples = flat.width, flat.height, flat.samples

    def close(self):
        if self._h:
            lib().or_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        o = OrInfo()
        lib().or_scene_info(self._h, C.byref(o))
        return {k: getattr(o, k) for k, _ in o._fields_}

    def bvh_order(self) -> np.ndarray:
        ids = np.zeros(self.flat.n_tris, dtype=np.int32)
        lib().or_scene_bvh_order(self._h, _p(ids, C.c_int32))
        return ids

    def bvh_nodes(self) -> np.ndarray:
        n = self.info()["n_nodes"]
        out = np.zeros((n, 10), dtype=np.float64)
        lib().or_scene_bvh_nodes(self._h, _p(out))
        return out

    def render(self, seed: int = 0, n_threads: int = 0, rows=None, want_rgb=True, want_mean=True, want_var=False):
        """render_scene (rendering.rs:21-69).  rows = (y0, y1, step) restricts to a row subset (bounded CPU samples)."""
        W, H = self.width, self.height
        y0, y1, step = rows if rows is not None else (0, H, 1)
        rgb = np.zeros((H, W, 3), dtype=np.uint8) if want_rgb else None
        mean = np.zeros((H, W, 3), dtype=np.float64) if want_mean else None
        var = np.zeros((H, W, 3), dtype=np.float64) if want_var else None
        st = OrStats()
        rc = lib().or_render(self._h, C.c_uint64(seed), C.c_int(n_threads), C.c_int(y0), C.c_int(y1), C.c_int(step),
                             _p(rgb, C.c_uint8), _p(mean), _p(var), C.byref(st))
        if rc != 0:
            raise RuntimeError("or_render failed (samples/width/height must be > 0)")
        return {"rgb": rgb, "mean": mean, "var": var, "stats": st.as_dict()}

    def primary_rays(self, xy, xi) -> np.ndarray:
        xy = np.ascontiguousarray(xy, dtype=np.int32)
        xi = _d(xi)
        out = np.zeros((xy.shape[0], 6), dtype=np.float64)
        lib().or_primary_rays(self._h, _p(xy, C.c_int32), _p(xi), C.c_int64(xy.shape[0]), _p(out))
        return out

    def trace_primary(self, rays, want_second=True):
        rays = _d(rays)
        n = rays.shape[0]
        tid = np.zeros(n, dtype=np.int32)
        t, u, v = np.zeros(n), np.zeros(n), np.zeros(n)
        second = np.zeros(n) if want_second else None
        st = OrStats()
        lib().or_trace_primary(self._h, _p(rays), C.c_int64(n), _p(tid, C.c_int32), _p(t), _p(u), _p(v), _p(second), C.byref(st))
        return {"tri_id": tid, "t": t, "u": u, "v": v, "second_t": second, "stats": st.as_dict()}

    def trace_hits(self, rays) -> np.ndarray:
        """(n, 9): t, normal_geometry xyz, normal_shading xyz, original primitive id, is_outer_to_inner."""
        rays = _d(rays)
        out = np.zeros((rays.shape[0], 9), dtype=np.float64)
        lib().or_trace_hits(self._h, _p(rays), C.c_int64(rays.shape[0]), _p(out))
        return out

    def pdf_light(self, point, l):
        point, l = _d(point), _d(l)
        out = np.zeros(point.shape[0])
        lib().or_pdf_light(self._h, _p(point), _p(l), C.c_int64(point.shape[0]), _p(out))
        return out

    def pdf_mix(self, point, n, l, v, mat):
        point, n, l, v, mat = _d(point), _d(n), _d(l), _d(v), _d(mat)
        out = np.zeros(point.shape[0])
        lib().or_pdf_mix(self._h, _p(point), _p(n), _p(l), _p(v), _p(mat), C.c_int64(point.shape[0]), _p(out))
        return out

    def sample_light(self, light_idx, point, draws):
        """draws (n, 2..4): triangle (u, v); box (x01, sign, c1, c2 in [-1, 1)); ellipsoid (unit sphere xyz) -- see or_sample_light."""
        light_idx = np.ascontiguousarray(light_idx, dtype=np.int32)
        point, draws = _d(point), _d(draws)
        if draws.shape[1] < 4:
            draws = _d(np.concatenate([draws, np.zeros((draws.shape[0], 4 - draws.shape[1]))], axis=1))
        out = np.zeros((point.shape[0], 3))
        lib().or_sample_light(self._h, _p(light_idx, C.c_int32), _p(point), _p(draws), C.c_int64(point.shape[0]), _p(out))
        return out


# ---- scene-independent unit functions -------------------------------------------------------------------------
def brdf(l, n, v, mat):
    l, n, v, mat = _d(l), _d(n), _d(v), _d(mat)
    out = np.zeros((l.shape[0], 3))
    lib().or_brdf(_p(l), _p(n), _p(v), _p(mat), C.c_int64(l.shape[0]), _p(out))
    return out


def specular_brdf(l, n, v, h, rough):
    l, n, v, h, rough = _d(l), _d(n), _d(v), _d(h), _d(rough)
    out = np.zeros(l.shape[0])
    lib().or_specular_brdf(_p(l), _p(n), _p(v), _p(h), _p(rough), C.c_int64(l.shape[0]), _p(out))
    return out


def pdf_cosine(n, l):
    n, l = _d(n), _d(l)
    out = np.zeros(n.shape[0])
    lib().or_pdf_cosine(_p(n), _p(l), C.c_int64(n.shape[0]), _p(out))
    return out


def pdf_vndf(n, l, v, rough):
    n, l, v, rough = _d(n), _d(l), _d(v), _d(rough)
    out = np.zeros(n.shape[0])
    lib().or_pdf_vndf(_p(n), _p(l), _p(v), _p(rough), C.c_int64(n.shape[0]), _p(out))
    return out


def sample_cosine(n, sphere_unit):
    n, s = _d(n), _d(sphere_unit)
    out = np.zeros((n.shape[0], 3))
    lib().or_sample_cosine(_p(n), _p(s), C.c_int64(n.shape[0]), _p(out))
    return out


def sample_vndf(n, v, rough, u12):
    n, v, rough, u12 = _d(n), _d(v), _d(rough), _d(u12)
    out = np.zeros((n.shape[0], 3))
    lib().or_sample_vndf(_p(n), _p(v), _p(rough), _p(u12), C.c_int64(n.shape[0]), _p(out))
    return out


def sample_vndf_cap(n, v, rough, u12):
    """The DEVICE's VNDF sampler (spherical caps), f64 restatement -- not a reference function; see oracle.cpp."""
    n, v, rough, u12 = _d(n), _d(v), _d(rough), _d(u12)
    out = np.zeros((n.shape[0], 3))
    lib().or_sample_vndf_cap(_p(n), _p(v), _p(rough), _p(u12), C.c_int64(n.shape[0]), _p(out))
    return out


def color_to_pixel(rgb):
    rgb = _d(rgb)
    out = np.zeros((rgb.shape[0], 3), dtype=np.uint8)
    lib().or_color_to_pixel(_p(rgb), C.c_int64(rgb.shape[0]), _p(out, C.c_uint8))
    return out


def intersect_triangle(o, d, abc):
    o, d, abc = _d(o), _d(d), _d(abc)
    tuv = np.zeros(3)
    hit = lib().or_intersect_triangle(_p(o), _p(d), _p(abc), _p(tuv))
    return (bool(hit), tuv)


def aabb_first_hit(o, d, mn, mx):
    o, d, mn, mx = _d(o), _d(d), _d(mn), _d(mx)
    t = C.c_double(0.0)
    outer = C.c_int32(0)
    hit = lib().or_aabb_first_hit(_p(o), _p(d), _p(mn), _p(mx), C.byref(t), C.byref(outer))
    return (bool(hit), t.value, bool(outer.value))


def intersect_shape(kind, params, o, d, upper=np.inf):
    """Object-space hits of one shape (geometry.rs:79-91 + own-spec arms): list of (t, normal xyz, is_outer_to_inner)."""
    params = _d(np.resize(np.asarray(params, dtype=np.float64), 9) if np.size(params) < 9 else params)
    o, d = _d(o), _d(d)
    out = np.zeros(10)
    lib().or_intersect_shape.restype = C.c_int
    n = lib().or_intersect_shape(C.c_int32(kind), _p(params), _p(o), _p(d), C.c_double(upper), _p(out))
    return [(out[5 * k], out[5 * k + 1:5 * k + 4].copy(), bool(out[5 * k + 4])) for k in range(n)]


def dielectric(n, v, ior_outer_u):
    """[OWN SPEC] smooth dielectric direction choice: (cnt, 4) = direction xyz, refracted flag."""
    n, v, iou = _d(n), _d(v), _d(ior_outer_u)
    out = np.zeros((n.shape[0], 4))
    lib().or_dielectric(_p(n), _p(v), _p(iou), C.c_int64(n.shape[0]), _p(out))
    return out


def quat_transform(q_ijkw, v, conjugate=False):
    q, v = _d(q_ijkw), _d(v)
    out = np.zeros(3)
    lib().or_quat_transform(_p(q), _p(v), C.c_int32(1 if conjugate else 0), _p(out))
    return out


def object_aabb(kind, params, position, q_ijkw):
    params = _d(np.resize(np.asarray(params, dtype=np.float64), 9) if np.size(params) < 9 else params)
    out = np.zeros(6)
    lib().or_object_aabb(C.c_int32(kind), _p(params), _p(_d(position)), _p(_d(q_ijkw)), _p(out))
    return out[:3], out[3:]


def rng_u64(seed: int, cnt: int) -> np.ndarray:
    out = np.zeros(cnt, dtype=np.uint64)
    lib().or_rng_u64(C.c_uint64(seed), C.c_int64(cnt), _p(out, C.c_uint64))
    return out
