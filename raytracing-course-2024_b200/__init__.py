"""raytracing-course-2024_b200 -- Python front-end (ctypes) of librt_b200.so, the B200-native replacement of the
reference's per-pixel path-tracing loop.  The product is the C ABI in include/rt_api.h (C++/CUDA, sm_100a);
this module only mirrors the reference's call sequence for tests and benchmarks:

    main.rs:45-47   gltf::import + convert_gltf_to_scene   ->  convert_gltf_to_scene(path, w, h, samples)
    main.rs:55      render_scene(&scene) -> Vec<u8>         ->  render_scene(scene) -> bytes-like (H, W, 3) uint8
    main.rs:88-95   dump_rendered_to_ppm                    ->  dump_rendered_to_ppm(scene, rendered, path)

There is NO CPU fallback and no pure-Python path: if the shared library is missing or no CUDA device is present
the compute calls raise RtError.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(_HERE, "_build", "librt_b200.so")   # env override: A/B builds of the same ABI
CLI_PATH = os.path.join(_HERE, "_build", "raytracing-engine")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "rt_api.h")

RT_OK, RT_ERR_INVALID, RT_ERR_IO, RT_ERR_FORMAT, RT_ERR_CUDA, RT_ERR_LIMIT = range(6)
FN_BRDF, FN_PDF_COSINE, FN_PDF_VNDF, FN_PDF_LIGHT, FN_PDF_MIX, FN_SAMPLE_COSINE, FN_SAMPLE_VNDF, FN_SAMPLE_LIGHT, FN_PHILOX, FN_SAMPLE_LIGHT_GEN, FN_DIELECTRIC = range(1, 12)
_FN_WIDTHS = {FN_BRDF: (14, 3), FN_PDF_COSINE: (6, 1), FN_PDF_VNDF: (10, 1), FN_PDF_LIGHT: (6, 1), FN_PDF_MIX: (13, 1),
              FN_SAMPLE_COSINE: (5, 6), FN_SAMPLE_VNDF: (9, 3), FN_SAMPLE_LIGHT: (6, 3), FN_PHILOX: (4, 4), FN_SAMPLE_LIGHT_GEN: (8, 3),
              FN_DIELECTRIC: (9, 4)}
SHAPE_TRIANGLE, SHAPE_BOX, SHAPE_ELLIPSOID, SHAPE_PLANE = range(4)
MATERIAL_PBR, MATERIAL_DIELECTRIC = 0, 1

_dp = C.POINTER(C.c_double)


class RtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"rt error {code}: {message}")
        self.code = code


class RtSceneDesc(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("samples", C.c_int32), ("ray_depth", C.c_int32),
        ("bg_color", C.c_double * 3), ("camera_position", C.c_double * 3), ("camera_forward", C.c_double * 3),
        ("camera_right", C.c_double * 3), ("camera_up", C.c_double * 3),
        ("camera_fov_x", C.c_double), ("camera_fov_y", C.c_double),
        ("n_tris", C.c_int32), ("reserved0", C.c_int32),
        ("tri_v", _dp), ("tri_n", _dp), ("tri_material", _dp), ("tri_emission", _dp),
    ]


_ip = C.POINTER(C.c_int32)


class RtSceneDesc2(C.Structure):
    _fields_ = [("base", RtSceneDesc), ("shape_kind", _ip), ("position", _dp), ("rotation", _dp), ("ior", _dp), ("material_kind", _ip)]


class RtSceneInfo(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n_tris", "n_lights", "n_materials", "n_nodes", "n_leaves", "bvh_depth", "max_leaf_size",
                                         "bvh_validate_failures", "scene_in_shared_memory", "device")] + [("device_bytes", C.c_int64)] + \
               [("bvh_builder", C.c_int32), ("reserved1", C.c_int32), ("bvh_build_ms", C.c_double), ("n_infinite", C.c_int32), ("general_primitives", C.c_int32)]


class RtRenderParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("sample_begin", C.c_int32), ("sample_end", C.c_int32), ("max_attempts", C.c_int32),
                ("collect_stats", C.c_int32), ("kernel_variant", C.c_int32), ("tile_shard_index", C.c_int32), ("tile_shard_count", C.c_int32),
                ("reserved", C.c_int32)]


class RtStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("samples", "segments", "vertices", "attempts", "node_tests", "tri_tests", "light_tri_tests",
                                          "attempt_cap_hits", "nonfinite_samples", "kernel_launches")] + [("kernel_ms", C.c_double), ("total_ms", C.c_double)] + \
               [(k, C.c_int32) for k in ("kernel", "block_threads", "blocks_per_sm", "grid_blocks", "regs_per_thread", "smem_bytes_per_block",
                                         "scene_in_shared_memory", "n_chunks")] + [("render_ms", C.c_double), ("reduce_ms", C.c_double), ("resolve_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU) -> _build/librt_b200.so + the CLI."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building librt_b200.so failed")
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Fails loudly if it was not built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RtError(RT_ERR_CUDA, f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        L.rt_last_error.restype = C.c_char_p
        vp = C.c_void_p
        L.rt_scene_load_gltf.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
        L.rt_scene_load_text.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
        L.rt_scene_load.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
        L.rt_scene_create.argtypes = [C.POINTER(RtSceneDesc), C.c_int32, C.POINTER(vp)]
        L.rt_scene_create2.argtypes = [C.POINTER(RtSceneDesc2), C.c_int32, C.POINTER(vp)]
        L.rt_scene_get_desc2.argtypes = [vp, C.POINTER(RtSceneDesc2)]
        L.rt_trace_hits.argtypes = [vp, _dp, C.c_int64, _dp]
        L.rt_scene_destroy.argtypes = [vp]
        L.rt_scene_destroy.restype = None
        L.rt_scene_get_desc.argtypes = [vp, C.POINTER(RtSceneDesc)]
        L.rt_scene_info.argtypes = [vp, C.POINTER(RtSceneInfo)]
        L.rt_scene_set_frame.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
        L.rt_scene_get_bvh.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        L.rt_scene_get_quantised_bvh.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        L.rt_render.argtypes = [vp, C.POINTER(RtRenderParams), C.POINTER(C.c_uint8), C.POINTER(RtStats)]
        L.rt_render_linear.argtypes = [vp, C.POINTER(RtRenderParams), C.POINTER(C.c_float), C.POINTER(RtStats)]
        L.rt_render_accumulate_device.argtypes = [vp, C.POINTER(RtRenderParams), vp, vp, C.POINTER(RtStats)]
        L.rt_resolve_device.argtypes = [vp, C.c_int32, C.c_int32, vp, vp]
        L.rt_multi_init.argtypes = [C.POINTER(vp), C.c_int32]
        L.rt_render_multi.argtypes = [C.POINTER(vp), C.c_int32, C.POINTER(RtRenderParams), C.POINTER(C.c_uint8), C.POINTER(RtStats)]
        L.rt_trace_primary.argtypes = [vp, _dp, C.c_int64, C.c_int32, C.POINTER(C.c_int32), _dp]
        L.rt_primary_rays.argtypes = [vp, C.POINTER(C.c_int32), _dp, C.c_int64, _dp]
        L.rt_eval.argtypes = [vp, C.c_int32, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_float)]
        L.rt_write_ppm.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint8), C.c_int32]
        L.rt_write_png.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
        L.rt_measure_fp32_peak.argtypes = [C.c_int32, _dp, _dp]
        L.rt_device_count.argtypes = [C.POINTER(C.c_int32)]
        _lib = L
    return _lib


def _check(rc):
    if rc != RT_OK:
        raise RtError(rc, lib().rt_last_error().decode("utf-8", "replace"))


def debug_bounds_violations(device=0, reset=False):
    """Counters of the -DRT_DEBUG_BOUNDS build (RT_B200_LIB=.../_build_dbg/librt_b200.so); RtError in a regular build."""
    out = (C.c_uint64 * 7)()
    lib().rt_debug_bounds_violations.argtypes = [C.c_int32, C.POINTER(C.c_uint64), C.c_int32]
    _check(lib().rt_debug_bounds_violations(device, out, 1 if reset else 0))
    return dict(zip(("pool", "queue", "stack", "blob", "prim", "layer", "chunk"), [int(x) for x in out]))


def release_device_cache():
    """Frees the frame buffers rt_scene_destroy parked for reuse (include/rt_api.h)."""
    _check(lib().rt_release_device_cache())


def device_count() -> int:
    n = C.c_int32(0)
    rc = lib().rt_device_count(C.byref(n))
    return int(n.value) if rc == RT_OK else 0


def _params(seed=0, sample_begin=0, sample_end=0, max_attempts=0, collect_stats=False, kernel_variant=0, tile_shard=(0, 0)):
    p = RtRenderParams()
    p.seed, p.sample_begin, p.sample_end = seed, sample_begin, sample_end
    p.max_attempts, p.collect_stats, p.kernel_variant = max_attempts, 1 if collect_stats else 0, kernel_variant
    p.tile_shard_index, p.tile_shard_count = tile_shard
    return p


class Scene:
    """Owner of an RtScene* (the reference's `Scene`, scene.rs:22-39, resident on one GPU)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    # -- construction -------------------------------------------------------------------------------------------
    @classmethod
    def from_gltf(cls, path, width, height, samples, device=0):
        h = C.c_void_p()
        _check(lib().rt_scene_load_gltf(os.fsencode(path), width, height, samples, device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_text(cls, path, width=0, height=0, samples=0, device=0):
        """The course's text scene format (own spec: no parser at reference HEAD); 0 keeps the file's DIMENSIONS / SAMPLES."""
        h = C.c_void_p()
        _check(lib().rt_scene_load_text(os.fsencode(path), width, height, samples, device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_file(cls, path, width=0, height=0, samples=0, device=0):
        h = C.c_void_p()
        _check(lib().rt_scene_load(os.fsencode(path), width, height, samples, device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_arrays(cls, *, width, height, samples, ray_depth, bg_color, camera_position, camera_forward, camera_right, camera_up,
                    camera_fov_x, camera_fov_y, tri_v, tri_n, tri_material, tri_emission, device=0,
                    shape_kind=None, position=None, rotation=None, ior=None, material_kind=None):
        """rt_scene_create, or rt_scene_create2 when any general-primitive array (shape_kind, position, rotation (i,j,k,w), ior,
        material_kind) is given."""
        d2 = RtSceneDesc2()
        d = d2.base
        d.width, d.height, d.samples, d.ray_depth = width, height, samples, ray_depth
        for name, val in (("bg_color", bg_color), ("camera_position", camera_position), ("camera_forward", camera_forward),
                          ("camera_right", camera_right), ("camera_up", camera_up)):
            getattr(d, name)[:] = [float(x) for x in val]
        d.camera_fov_x, d.camera_fov_y = camera_fov_x, camera_fov_y
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (tri_v, tri_n, tri_material, tri_emission)]
        d.n_tris = int(arrs[0].size // 9)
        d.tri_v, d.tri_n, d.tri_material, d.tri_emission = (a.ctypes.data_as(_dp) for a in arrs)
        h = C.c_void_p()
        if all(a is None for a in (shape_kind, position, rotation, ior, material_kind)):
            _check(lib().rt_scene_create(C.byref(d), device, C.byref(h)))
            return cls(h.value)
        keep = []
        for name, val, dt, pt in (("shape_kind", shape_kind, np.int32, _ip), ("position", position, np.float64, _dp), ("rotation", rotation, np.float64, _dp),
                                  ("ior", ior, np.float64, _dp), ("material_kind", material_kind, np.int32, _ip)):
            if val is not None:
                a = np.ascontiguousarray(val, dtype=dt)
                keep.append(a)
                setattr(d2, name, a.ctypes.data_as(pt))
        _check(lib().rt_scene_create2(C.byref(d2), device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_flat(cls, fl, device=0, **over):
        """From an object with the FlatScene attribute names (oracle.FlatScene; tests only -- this module never imports oracle)."""
        kw = dict(width=fl.width, height=fl.height, samples=fl.samples, ray_depth=fl.ray_depth, bg_color=fl.bg_color, camera_position=fl.camera_position,
                  camera_forward=fl.camera_forward, camera_right=fl.camera_right, camera_up=fl.camera_up, camera_fov_x=fl.camera_fov_x,
                  camera_fov_y=fl.camera_fov_y, tri_v=fl.tri_v, tri_n=fl.tri_n, tri_material=fl.tri_material, tri_emission=fl.tri_emission,
                  shape_kind=getattr(fl, "kind", None), position=getattr(fl, "position", None), rotation=getattr(fl, "rotation", None),
                  ior=getattr(fl, "ior", None), material_kind=getattr(fl, "mat_kind", None), device=device)
        kw.update(over)
        return cls.from_arrays(**kw)

    def close(self):
        if self._h:
            lib().rt_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inspection ---------------------------------------------------------------------------------------------
    def desc(self) -> dict:
        d2 = RtSceneDesc2()
        _check(lib().rt_scene_get_desc2(self._h, C.byref(d2)))
        d = d2.base
        n = d.n_tris

        def arr(p, w):
            return np.ctypeslib.as_array(p, shape=(n, w)).copy() if n else np.zeros((0, w))
        out = {k: getattr(d, k) for k in ("width", "height", "samples", "ray_depth", "camera_fov_x", "camera_fov_y", "n_tris")}
        for k in ("bg_color", "camera_position", "camera_forward", "camera_right", "camera_up"):
            out[k] = np.array(list(getattr(d, k)))
        out.update(tri_v=arr(d.tri_v, 9), tri_n=arr(d.tri_n, 9), tri_material=arr(d.tri_material, 5), tri_emission=arr(d.tri_emission, 3))
        if d2.shape_kind:                                   # general-primitive scene
            out.update(shape_kind=np.ctypeslib.as_array(d2.shape_kind, shape=(n,)).copy(), position=arr(d2.position, 3), rotation=arr(d2.rotation, 4),
                       ior=np.ctypeslib.as_array(d2.ior, shape=(n,)).copy(), material_kind=np.ctypeslib.as_array(d2.material_kind, shape=(n,)).copy())
        return out

    def info(self) -> dict:
        i = RtSceneInfo()
        _check(lib().rt_scene_info(self._h, C.byref(i)))
        return {k: getattr(i, k) for k, _ in i._fields_}

    @property
    def width(self):
        return self.desc_scalar("width")

    def desc_scalar(self, key):
        d = RtSceneDesc()
        _check(lib().rt_scene_get_desc(self._h, C.byref(d)))
        return getattr(d, key)

    def set_frame(self, width, height, samples):
        _check(lib().rt_scene_set_frame(self._h, width, height, samples))

    def bvh(self):
        i = self.info()
        nodes = np.zeros((i["n_nodes"], 14), dtype=np.float32)
        order = np.zeros(max(i["n_tris"], 1), dtype=np.int32)
        _check(lib().rt_scene_get_bvh(self._h, nodes.ctypes.data_as(C.POINTER(C.c_float)), order.ctypes.data_as(C.POINTER(C.c_int32))))
        return nodes, order[: i["n_tris"]]

    def quantised_bvh(self):
        """(words (n_nodes, 8) uint32, grid origin (3,), grid cell (3,)) of the quantised node array, or None when the scene has none."""
        n = self.info()["n_nodes"]
        words, grid, present = np.zeros((n, 8), dtype=np.uint32), np.zeros(6, dtype=np.float32), C.c_int32(0)
        _check(lib().rt_scene_get_quantised_bvh(self._h, words.ctypes.data_as(C.POINTER(C.c_uint32)), grid.ctypes.data_as(C.POINTER(C.c_float)), C.byref(present)))
        return (words, grid[:3].copy(), grid[3:].copy()) if present.value else None

    # -- the hot path -------------------------------------------------------------------------------------------
    def render(self, **kw):
        """render_scene (rendering.rs:21-69): returns ((H, W, 3) uint8, stats dict)."""
        W, H = self.desc_scalar("width"), self.desc_scalar("height")
        out = np.zeros((H, W, 3), dtype=np.uint8)
        st = RtStats()
        p = _params(**kw)
        _check(lib().rt_render(self._h, C.byref(p), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st)))
        return out, st.as_dict()

    def render_into(self, out: np.ndarray, **kw):
        """rt_render into a caller-owned host buffer (e.g. pinned memory); returns the stats dict."""
        st = RtStats()
        p = _params(**kw)
        _check(lib().rt_render(self._h, C.byref(p), C.cast(out.ctypes.data, C.POINTER(C.c_uint8)), C.byref(st)))
        return st.as_dict()

    def render_linear(self, **kw):
        W, H = self.desc_scalar("width"), self.desc_scalar("height")
        out = np.zeros((H, W, 3), dtype=np.float32)
        st = RtStats()
        p = _params(**kw)
        _check(lib().rt_render_linear(self._h, C.byref(p), out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st)))
        return out, st.as_dict()

    def render_accumulate_device(self, accum_ptr: int, stream_ptr: int = 0, want_stats=False, **kw):
        """Adds the radiance sums of a sample shard into a device accumulator (W*H*4 floats); plain pointers."""
        st = RtStats()
        p = _params(**kw)
        _check(lib().rt_render_accumulate_device(self._h, C.byref(p), C.c_void_p(accum_ptr), C.c_void_p(stream_ptr), C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def trace_primary(self, rays, precision=32):
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        n = rays.shape[0]
        tid = np.zeros(n, dtype=np.int32)
        t = np.zeros(n, dtype=np.float64)
        _check(lib().rt_trace_primary(self._h, rays.ctypes.data_as(_dp), n, precision, tid.ctypes.data_as(C.POINTER(C.c_int32)), t.ctypes.data_as(_dp)))
        return tid, t

    def trace_hits(self, rays):
        """(n, 9): t, normal_geometry xyz, normal_shading xyz (normalised, object space), original primitive id, is_outer_to_inner."""
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        out = np.zeros((rays.shape[0], 9), dtype=np.float64)
        _check(lib().rt_trace_hits(self._h, rays.ctypes.data_as(_dp), rays.shape[0], out.ctypes.data_as(_dp)))
        return out

    def primary_rays(self, xy, xi):
        xy = np.ascontiguousarray(xy, dtype=np.int32)
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        out = np.zeros((xy.shape[0], 6), dtype=np.float64)
        _check(lib().rt_primary_rays(self._h, xy.ctypes.data_as(C.POINTER(C.c_int32)), xi.ctypes.data_as(_dp), xy.shape[0], out.ctypes.data_as(_dp)))
        return out

    def eval(self, fn, inputs):
        return eval_fn(fn, inputs, scene=self)


def eval_fn(fn, inputs, scene: Scene | None = None):
    """Runs one of the device unit functions (RT_FN_*) on (n, width) float32 inputs."""
    win, wout = _FN_WIDTHS[fn]
    x = np.ascontiguousarray(inputs, dtype=np.float32).reshape(-1, win)
    out = np.zeros((x.shape[0], wout), dtype=np.float32)
    h = scene._h if scene is not None else None
    _check(lib().rt_eval(h, fn, x.ctypes.data_as(C.POINTER(C.c_float)), x.shape[0], out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


def resolve_device(accum_ptr: int, width: int, height: int, rgb_ptr: int, stream_ptr: int = 0):
    _check(lib().rt_resolve_device(C.c_void_p(accum_ptr), width, height, C.c_void_p(rgb_ptr), C.c_void_p(stream_ptr)))


def measure_fp32_peak(device=0):
    t, m = C.c_double(0), C.c_double(0)
    _check(lib().rt_measure_fp32_peak(device, C.byref(t), C.byref(m)))
    return t.value, m.value


# ---- the reference's call sequence (main.rs:45-67) ------------------------------------------------------------
def convert_gltf_to_scene(path, width, height, samples, device=0) -> Scene:
    return Scene.from_gltf(path, width, height, samples, device)


def parse_file_content(path, width=0, height=0, samples=0, device=0) -> Scene:
    """main.rs:48 `// let scene = parse_file_content(file_lines);` -- the text-scene entry the reference dropped before HEAD."""
    return Scene.from_text(path, width, height, samples, device)


def render_scene(scene: Scene, seed: int = 0) -> np.ndarray:
    return scene.render(seed=seed)[0]


def dump_rendered_to_ppm(scene: Scene, rendered: np.ndarray, path, append=False):
    rendered = np.ascontiguousarray(rendered, dtype=np.uint8)
    H, W = rendered.shape[0], rendered.shape[1]
    _check(lib().rt_write_ppm(os.fsencode(path), W, H, rendered.ctypes.data_as(C.POINTER(C.c_uint8)), 1 if append else 0))


def dump_rendered_to_png(scene: Scene, rendered: np.ndarray, path):
    """main.rs:75-86: the rendered RGB8 frame as a PNG."""
    rendered = np.ascontiguousarray(rendered, dtype=np.uint8)
    H, W = rendered.shape[0], rendered.shape[1]
    _check(lib().rt_write_png(os.fsencode(path), W, H, rendered.ctypes.data_as(C.POINTER(C.c_uint8))))


def render_multi(scenes, **kw):
    """rt_render_multi: one frame sharded by samples over the devices the scenes live on (one process, one NCCL reduce)."""
    W, H = scenes[0].desc_scalar("width"), scenes[0].desc_scalar("height")
    out = np.zeros((H, W, 3), dtype=np.uint8)
    st = RtStats()
    arr = (C.c_void_p * len(scenes))(*[s._h.value for s in scenes])
    _check(lib().rt_render_multi(arr, len(scenes), C.byref(_params(**kw)), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st)))
    return out, st.as_dict()
