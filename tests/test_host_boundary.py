"""CPU tests of the product's host side and of the C-ABI boundary (no compute calls: there is no GPU here):
the library loads and exports every symbol include/rt_api.h declares, the C++ glTF loader agrees bit-for-bit with
the oracle's restatement of gltf_to_scene.rs, the flattened device BVH satisfies the reference's validate_bvh
invariants (bvh.rs:299-322) and yields the oracle's nearest hits when walked on the host, error codes replace the
reference's panics, and the PPM writer reproduces main.rs:88-95 byte for byte."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, scene_path

SCENES4 = ["practice7_1", "practice7_4", "practice7_2", "practice7_3"]


def test_library_exports_every_declared_symbol(rt):
    hdr = open(os.path.join(ROOT, "include", "rt_api.h")).read()
    declared = sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 18
    L = rt.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in rt_api.h but not exported by librt_b200.so"
    assert L.rt_api_version() == 2


def test_header_is_plain_c(tmp_path):
    """The boundary must be bindable from C/Rust: compile the header as C99."""
    src = tmp_path / "t.c"
    src.write_text('#include "rt_api.h"\nint main(void){ RtSceneDesc d; RtRenderParams p; RtStats s; (void)d; (void)p; (void)s; return RT_OK; }\n')
    rc = os.system(f"gcc -std=c99 -Wall -Werror -pedantic -I{ROOT}/include -c {src} -o {tmp_path}/t.o")
    assert rc == 0


@pytest.mark.parametrize("name", SCENES4)
def test_loader_matches_oracle_loader_bit_exact(rt, oracle, name):
    """rt_scene_load_gltf (C++) vs oracle/gltf_ref.py (numpy), both restating gltf_to_scene.rs:21-256."""
    sc = rt.Scene.from_gltf(scene_path(name), 40, 30, 7, device=-1)
    d = sc.desc()
    fl = oracle.convert_gltf_to_scene(scene_path(name), 40, 30, 7)
    assert (d["width"], d["height"], d["samples"], d["ray_depth"]) == (40, 30, 7, 6)
    for k in ("tri_v", "tri_n", "tri_material", "tri_emission", "camera_position", "camera_forward", "camera_right", "camera_up", "bg_color"):
        assert np.array_equal(d[k], getattr(fl, k)), k
    assert d["camera_fov_x"] == fl.camera_fov_x and d["camera_fov_y"] == fl.camera_fov_y
    info = sc.info()
    assert info["n_tris"] == fl.n_tris and info["n_lights"] == len(fl.light_ids)
    assert info["bvh_validate_failures"] == 0
    assert info["n_leaves"] == info["n_nodes"] + 1 and info["max_leaf_size"] <= 8
    sc.close()


def _walk_flat_bvh(nodes, order, tri_v, o, d):
    """Host emulation of the device traversal (rt_device.cuh trace_nearest) in f64: returns (orig id, t)."""
    inv = 1.0 / np.where(np.abs(d) < 1e-20, 1e-20, d)
    best_t, best = np.inf, -1
    stack, cur = [], 0
    while True:
        while cur >= 0:
            nd = nodes[cur]
            hits = []
            for s in range(2):
                mn, mx = nd[6 * s: 6 * s + 3].astype(np.float64), nd[6 * s + 3: 6 * s + 6].astype(np.float64)
                t0, t1 = (mn - o) * inv, (mx - o) * inv
                tmin = max(np.minimum(t0, t1).max(), 0.0)
                tmax = min(np.maximum(t0, t1).min(), best_t)
                hits.append((tmin <= tmax, tmin, int(nd[12 + s])))
            (h0, t0_, c0), (h1, t1_, c1) = hits
            if h0 and h1:
                if t1_ < t0_:
                    stack.append(c0); cur = c1
                else:
                    stack.append(c1); cur = c0
            elif h0 or h1:
                cur = c0 if h0 else c1
            else:
                if not stack:
                    return best, best_t
                cur = stack.pop()
        code = ~cur
        first, cnt = code >> 3, (code & 7) + 1
        for k in range(first, first + cnt):
            v = tri_v[order[k]]
            a, e1, e2 = v[0:3], v[3:6] - v[0:3], v[6:9] - v[0:3]
            p = np.cross(d, e2); det = e1 @ p
            if det == 0:
                continue
            tv = o - a; u = (tv @ p) / det; q = np.cross(tv, e1); vv = (d @ q) / det; t = (e2 @ q) / det
            if u >= 0 and vv >= 0 and u + vv <= 1 and t > 0 and t < best_t:
                best_t, best = t, int(order[k])
        if not stack:
            return best, best_t
        cur = stack.pop()


@pytest.mark.parametrize("name", ["practice7_1", "practice7_4", "practice7_3"])
def test_flat_bvh_walk_matches_oracle_hits(rt, oracle, name):
    W = 24
    sc = rt.Scene.from_gltf(scene_path(name), W, W, 1, device=-1)
    nodes, order = sc.bvh()
    assert sorted(order.tolist()) == list(range(sc.info()["n_tris"]))
    # child refs are exact in float32 only below 2^24: check the encoding survived
    assert np.all(np.abs(nodes[:, 12:14]) < 2 ** 24)
    fl = oracle.convert_gltf_to_scene(scene_path(name), W, W, 1)
    osc = oracle.OracleScene(fl)
    xy = np.array([[x, y] for y in range(W) for x in range(W)], dtype=np.int32)
    rays = osc.primary_rays(xy, np.full((len(xy), 2), 0.5))
    # add bounce-like rays from inside the box
    rng = np.random.default_rng(1)
    dirs = rng.normal(size=(200, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    inner = np.concatenate([np.tile([0.1, 0.2, 0.3], (200, 1)), dirs], axis=1)
    rays = np.concatenate([rays[::3], inner])
    ref = osc.trace_primary(rays, want_second=False)
    tv = fl.tri_v
    for i, r in enumerate(rays):
        tid, t = _walk_flat_bvh(nodes, order, tv, r[:3], r[3:])
        if tid != ref["tri_id"][i]:
            # only exact ties (shared edges) may resolve differently: same distance
            assert t == pytest.approx(ref["t"][i], rel=1e-12), (i, tid, ref["tri_id"][i])
        else:
            assert (t == ref["t"][i]) or t == pytest.approx(ref["t"][i], rel=1e-9)
    sc.close()


@pytest.mark.parametrize("name", ["practice7_2", "practice7_3"])
def test_quantised_nodes_enclose_the_full_precision_nodes(rt, name):
    """Triangle scenes too large for shared memory walk 32-byte nodes whose planes are 16-bit coordinates on a global grid
    (rt_api.cu quant_nodes, rt_device.cuh pair_step_quant).  Containment check in the spirit of validate_bvh (bvh.rs:299-322):
    every grid box, reconstructed with the float constants the device uses, encloses the f32 box of the same child with at least
    1.5 cells of slack (what absorbs the FP32 plane evaluation) and at most 5 (tightness); the references describe the same tree."""
    sc = rt.Scene.from_gltf(scene_path(name), 8, 8, 1, device=-1)
    nodes, _ = sc.bvh()
    q = sc.quantised_bvh()
    assert q is not None, "large meshes carry quantised nodes"
    words, org, cell = q
    assert words.shape == (nodes.shape[0], 8) and (cell > 0).all()
    org, cell = org.astype(np.float64), cell.astype(np.float64)
    for child in (0, 1):
        lo = nodes[:, child * 6:child * 6 + 3].astype(np.float64)
        hi = nodes[:, child * 6 + 3:child * 6 + 6].astype(np.float64)
        shift = 16 * child
        qlo = np.stack([(words[:, 2 * a] >> shift) & 0xffff for a in range(3)], axis=1).astype(np.float64)
        qhi = np.stack([(words[:, 2 * a + 1] >> shift) & 0xffff for a in range(3)], axis=1).astype(np.float64)
        plo, phi = org + qlo * cell, org + qhi * cell
        assert (plo <= lo - 1.5 * cell + 1e-9 * np.abs(lo)).all() and (phi >= hi + 1.5 * cell - 1e-9 * np.abs(hi)).all()
        assert (plo >= lo - 5.0 * cell).all() and (phi <= hi + 5.0 * cell).all()
        ref_f = nodes[:, 12 + child].astype(np.int64)
        ref_q = words[:, 6 + child].astype(np.uint32).view(np.int32).astype(np.int64)
        inner = ref_f >= 0
        assert ((ref_q >= 0) == inner).all()
        assert (ref_q[inner] == ref_f[inner] * 32).all() and (ref_q[~inner] == ref_f[~inner]).all()
    small = rt.Scene.from_gltf(scene_path("practice7_4"), 8, 8, 1, device=-1)
    assert small.quantised_bvh() is None                     # shared-memory scenes keep the packed full-precision nodes only
    sc.close(); small.close()


def _tree_stats(sc):
    """(sum of the half-areas of the inner-node boxes, sorted leaf codes, set of triangles in the leaves) of a scene's flattened tree."""
    nodes, order = sc.bvh()
    lo = nodes[:, [0, 1, 2, 6, 7, 8]].reshape(-1, 2, 3).astype(np.float64)
    hi = nodes[:, [3, 4, 5, 9, 10, 11]].reshape(-1, 2, 3).astype(np.float64)
    ext = hi - lo
    area = ext[..., 0] * ext[..., 1] + ext[..., 1] * ext[..., 2] + ext[..., 2] * ext[..., 0]
    refs = nodes[:, 12:14].astype(np.int64)
    inner = refs >= 0
    leaves = np.sort(refs[~inner])
    tris = []
    for code in (~leaves).tolist():
        tris += list(range(code >> 3, (code >> 3) + (code & 7) + 1))
    return float(area[inner].sum()), float(area.sum()), leaves, sorted(order[tris].tolist())


@pytest.mark.parametrize("name", ["practice7_1", "practice7_4"])
def test_bottom_up_builder_gives_a_valid_tree_of_lower_cost(rt, name, monkeypatch):
    """Sets of <= 512 primitives are built bottom-up (greedy clustering by the area of the union + re-insertion, bvh_builder.cpp)
    instead of by the top-down SAH sweep (RT_BVH_AGGLO=0).  Both trees must hold every triangle exactly once and pass the
    containment check of bvh.rs:299-322; the bottom-up one must be cheaper under the surface-area cost both builders minimise."""
    monkeypatch.setenv("RT_BVH_AGGLO", "0")
    sweep = rt.Scene.from_gltf(scene_path(name), 8, 8, 1, device=-1)
    monkeypatch.delenv("RT_BVH_AGGLO")
    bottom_up = rt.Scene.from_gltf(scene_path(name), 8, 8, 1, device=-1)
    monkeypatch.setenv("RT_BVH_REINSERT", "0")
    no_reinsert = rt.Scene.from_gltf(scene_path(name), 8, 8, 1, device=-1)
    monkeypatch.delenv("RT_BVH_REINSERT")
    n = sweep.info()["n_tris"]
    stats = {}
    for label, sc in (("sweep", sweep), ("bottom_up", bottom_up), ("no_reinsert", no_reinsert)):
        info = sc.info()
        assert info["bvh_validate_failures"] == 0 and info["max_leaf_size"] <= 2 and info["n_tris"] == n
        inner_area, all_area, _, tris = _tree_stats(sc)
        assert tris == list(range(n)), label                       # every triangle in exactly one leaf
        stats[label] = all_area
    assert stats["bottom_up"] <= stats["no_reinsert"] * (1 + 1e-9)     # re-insertion never makes the tree worse
    assert stats["bottom_up"] < (0.95 if name == "practice7_4" else 1.0) * stats["sweep"], stats   # less box area than the sweep
    for sc in (sweep, bottom_up, no_reinsert):
        sc.close()


def test_bottom_up_top_of_a_large_tree_keeps_the_leaves(rt, monkeypatch):
    """regraft_top_sah (bvh_builder.cpp): the part of a large triangle mesh's tree above its 128 largest subtrees is rebuilt bottom-up
    and the subtrees are grafted back.  Same node count, same leaves, valid containment, less box area near the root."""
    monkeypatch.setenv("RT_BVH_TOP_AGGLO", "0")
    plain = rt.Scene.from_gltf(scene_path("practice7_3"), 8, 8, 1, device=-1)
    monkeypatch.delenv("RT_BVH_TOP_AGGLO")
    grafted = rt.Scene.from_gltf(scene_path("practice7_3"), 8, 8, 1, device=-1)
    ip, ig = plain.info(), grafted.info()
    assert ip["bvh_validate_failures"] == 0 and ig["bvh_validate_failures"] == 0
    assert ip["n_nodes"] == ig["n_nodes"] and ip["n_leaves"] == ig["n_leaves"] and ip["bvh_builder"] == 0 and ig["bvh_builder"] == 0
    _, area_p, leaves_p, tris_p = _tree_stats(plain)
    _, area_g, leaves_g, tris_g = _tree_stats(grafted)
    assert np.array_equal(leaves_p, leaves_g) and tris_p == tris_g == list(range(ip["n_tris"]))
    assert area_g < area_p                                             # the rebuilt top is what every ray pays for
    assert ig["bvh_depth"] <= ip["bvh_depth"] + 16
    plain.close(); grafted.close()


@pytest.mark.parametrize("n,poison", [(300, "nan"), (300, "huge"), (300, "same"), (3000, "nan"), (3000, "halfsame"), (3, "nan"), (513, "none")])
def test_builders_survive_degenerate_geometry(rt, n, poison):
    """Non-finite or astronomically large vertices must not hang or crash the builders (their boxes are parked at the origin: such a
    triangle cannot be hit anyway), and hundreds of coincident triangles -- every split costs the same, the cost-driven trees become
    chains deeper than the traversal stack -- fall back to median splits instead of failing with RT_ERR_LIMIT."""
    rng = np.random.default_rng(0)
    c = rng.uniform(-5, 5, (n, 1, 3))
    v = (c + rng.normal(0, 0.2, (n, 3, 3))).reshape(n, 9)
    if poison == "nan":
        v[min(5, n - 1), 2] = np.nan; v[min(100, n - 1), 7] = np.inf
    if poison == "huge":
        v[7] *= 1e35
    if poison == "same":
        v[:] = v[0]
    if poison == "halfsame":
        v[: n // 2] = v[0]
    sc = rt.Scene.from_arrays(width=16, height=16, samples=1, ray_depth=3, bg_color=[0, 0, 0], camera_position=[0, 0, 20], camera_forward=[0, 0, -1],
                              camera_right=[1, 0, 0], camera_up=[0, 1, 0], camera_fov_x=1.0, camera_fov_y=1.0, tri_v=v, tri_n=np.tile([0, 0, 1.0], (n, 3)).reshape(n, 9),
                              tri_material=np.tile([0.8, 0.8, 0.8, 0, 1.0], (n, 1)), tri_emission=np.zeros((n, 3)), device=-1)
    info = sc.info()
    assert info["n_tris"] == n and info["max_leaf_size"] <= 2
    assert info["bvh_depth"] <= 40, info["bvh_depth"]                    # (coincident triangles: median fallback, depth ~ log2 n)
    _, order = sc.bvh()
    assert sorted(order.tolist()) == list(range(n))
    if poison in ("none", "same", "halfsame"):
        assert info["bvh_validate_failures"] == 0
    else:
        assert info["bvh_validate_failures"] <= 2                          # only the poisoned triangles are outside their (parked) boxes
    sc.close()


def test_host_only_scene_refuses_compute_and_reports_errors(rt, tmp_path):
    sc = rt.Scene.from_gltf(scene_path("practice7_1"), 8, 8, 1, device=-1)
    with pytest.raises(rt.RtError) as e:
        sc.render()
    assert e.value.code == rt.RT_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(rt.RtError) as e:
        sc.trace_primary(np.zeros((1, 6)))
    assert e.value.code == rt.RT_ERR_CUDA
    sc.close()
    with pytest.raises(rt.RtError) as e:                       # the reference: unwrap() panic (main.rs:45)
        rt.Scene.from_gltf(str(tmp_path / "missing.gltf"), 8, 8, 1, device=-1)
    assert e.value.code == rt.RT_ERR_IO
    bad = tmp_path / "bad.gltf"
    bad.write_text('{"nodes": [ {"mesh": 0 ]')
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_gltf(str(bad), 8, 8, 1, device=-1)
    assert e.value.code == rt.RT_ERR_FORMAT
    ortho = tmp_path / "ortho.gltf"                             # the reference: todo!() (gltf_to_scene.rs:131-133)
    ortho.write_text('{"asset":{"version":"2.0"},"nodes":[{"camera":0}],"cameras":[{"type":"orthographic","orthographic":{"xmag":1,"ymag":1,"zfar":10,"znear":0.1}}]}')
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_gltf(str(ortho), 8, 8, 1, device=-1)
    assert e.value.code == rt.RT_ERR_FORMAT


def test_embedded_buffer_children_and_defaults(rt, oracle, tmp_path):
    """Loader rules the shipped scenes never exercise: data: URI buffers, u8 indices, a mesh without NORMAL and
    without a material (glTF defaults), and the reference's double visit of child nodes
    (gltf_to_scene.rs:42-52 + 245-255)."""
    import base64
    import json
    import struct
    pos = struct.pack("<9f", 0, 0, 0, 1, 0, 0, 0, 1, 0)
    idx = struct.pack("<3B", 0, 1, 2) + b"\0"
    blob = pos + idx
    g = {
        "asset": {"version": "2.0"},
        "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 3}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                      {"bufferView": 1, "componentType": 5121, "count": 3, "type": "SCALAR"}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}],
        "cameras": [{"type": "perspective", "perspective": {"yfov": 0.5, "znear": 0.1}}],
        "nodes": [{"children": [1], "translation": [0, 0, -3], "rotation": [0, 0.38268343, 0, 0.92387953]},
                  {"mesh": 0, "scale": [2, 2, 2], "translation": [1, 0, 0]},
                  {"camera": 0, "translation": [0, 0, 5]}],
    }
    path = tmp_path / "tiny.gltf"
    path.write_text(json.dumps(g))
    sc = rt.Scene.from_gltf(str(path), 8, 8, 1, device=-1)
    d = sc.desc()
    fl = oracle.convert_gltf_to_scene(str(path), 8, 8, 1)
    assert d["n_tris"] == 2 == fl.n_tris            # node 1 visited as a child of node 0 AND as a top-level node
    for k in ("tri_v", "tri_n", "tri_material", "tri_emission"):
        assert np.allclose(d[k], getattr(fl, k), rtol=0, atol=1e-15), k
    assert np.allclose(d["tri_material"][0], [1, 1, 1, 1, 1])           # glTF default material
    assert not np.allclose(d["tri_v"][0], d["tri_v"][1])                # parent transform applied only on one visit
    assert d["camera_fov_x"] == d["camera_fov_y"] == np.float32(0.5)    # aspect defaults to 1.0
    sc.close()


def test_scene_from_arrays_round_trip(rt, oracle):
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_4"), 16, 16, 2)
    sc = rt.Scene.from_arrays(width=16, height=16, samples=2, ray_depth=6, bg_color=fl.bg_color, camera_position=fl.camera_position,
                              camera_forward=fl.camera_forward, camera_right=fl.camera_right, camera_up=fl.camera_up,
                              camera_fov_x=fl.camera_fov_x, camera_fov_y=fl.camera_fov_y, tri_v=fl.tri_v, tri_n=fl.tri_n,
                              tri_material=fl.tri_material, tri_emission=fl.tri_emission, device=-1)
    d = sc.desc()
    assert np.array_equal(d["tri_v"], fl.tri_v) and sc.info()["n_lights"] == 2 and sc.info()["scene_in_shared_memory"] == 1
    sc.set_frame(3840, 2160, 1024)
    assert sc.desc_scalar("width") == 3840 and sc.desc_scalar("samples") == 1024
    sc.close()
    # empty scene is legal (everything misses)
    e = rt.Scene.from_arrays(width=4, height=4, samples=1, ray_depth=6, bg_color=[0, 0, 0], camera_position=[0, 0, 0], camera_forward=[0, 0, -1],
                             camera_right=[1, 0, 0], camera_up=[0, 1, 0], camera_fov_x=1.0, camera_fov_y=1.0, tri_v=np.zeros((0, 9)),
                             tri_n=np.zeros((0, 9)), tri_material=np.zeros((0, 5)), tri_emission=np.zeros((0, 3)), device=-1)
    assert e.info()["n_tris"] == 0 and e.info()["n_lights"] == 0
    e.close()


def test_write_ppm_bytes_and_append_quirk(rt, tmp_path):
    """main.rs:88-95: 'P6\\n{W} {H}\\n255\\n' + RGB bytes; main.rs:62-66 opens in append mode."""
    img = (np.arange(2 * 3 * 3) % 251).astype(np.uint8).reshape(2, 3, 3)
    p = tmp_path / "o.ppm"
    rt.dump_rendered_to_ppm(None, img, str(p))
    assert p.read_bytes() == b"P6\n3 2\n255\n" + img.tobytes()
    rt.dump_rendered_to_ppm(None, img, str(p))                       # default: truncate (documented deviation)
    assert p.read_bytes() == b"P6\n3 2\n255\n" + img.tobytes()
    rt.dump_rendered_to_ppm(None, img, str(p), append=True)          # reference behaviour on request
    assert p.read_bytes() == (b"P6\n3 2\n255\n" + img.tobytes()) * 2


def test_cli_reports_errors_instead_of_panicking(rt, tmp_path):
    import subprocess
    r = subprocess.run([rt.CLI_PATH], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([rt.CLI_PATH, str(tmp_path / "nope.gltf"), "8", "8", "1", str(tmp_path / "o.ppm")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot read" in r.stderr


def test_write_png_decodes_to_the_same_pixels(rt, tmp_path):
    """dump_rendered_to_png (main.rs:75-86): any PNG reader must give back the rendered bytes; sizes that need several
    stored-deflate blocks (> 65535 bytes) and rows that are not multiples of 4."""
    from PIL import Image
    rng = np.random.default_rng(5)
    for H, W in ((3, 5), (97, 301), (1, 1)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        p = tmp_path / f"o_{H}x{W}.png"
        rt.dump_rendered_to_png(None, img, str(p))
        back = np.asarray(Image.open(p).convert("RGB"))
        assert back.shape == img.shape and np.array_equal(back, img)


def test_malformed_gltf_is_rejected_not_read_out_of_bounds(rt, tmp_path):
    """accessor / bufferView sizes are untrusted.  Negative, fractional or huge `count` / `byteOffset` /
    `byteStride` / `byteLength`, accessors that leave their bufferView, non-integer child indices and absurdly nested JSON must
    produce RT_ERR_FORMAT (the gltf crate rejects such files; the reference then panics in `unwrap`, main.rs:45) -- never an
    out-of-bounds read."""
    import base64
    import json
    import struct
    blob = struct.pack("<9f", 0, 0, 0, 1, 0, 0, 0, 1, 0) + struct.pack("<3B", 0, 1, 2) + b"\0"

    def doc(**patch):
        g = {
            "asset": {"version": "2.0"},
            "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
            "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 3}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5121, "count": 3, "type": "SCALAR"}],
            "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}],
            "cameras": [{"type": "perspective", "perspective": {"yfov": 0.5, "znear": 0.1}}],
            "nodes": [{"mesh": 0}, {"camera": 0, "translation": [0, 0, 5]}],
        }
        for path, val in patch.items():
            keys = path.split("__")
            cur = g
            for k in keys[:-1]:
                cur = cur[int(k)] if k.isdigit() else cur[k]
            cur[int(keys[-1]) if keys[-1].isdigit() else keys[-1]] = val
        return g

    def load(g, name):
        p = tmp_path / name
        p.write_text(json.dumps(g) if not isinstance(g, str) else g)
        return rt.Scene.from_gltf(str(p), 8, 8, 1, device=-1)

    load(doc(), "ok.gltf").close()                                         # the unpatched document loads
    bad = {
        "count_negative": {"accessors__1__count": -1},
        "count_huge": {"accessors__0__count": 2 ** 62},
        "count_past_view": {"accessors__0__count": 4},
        "count_fraction": {"accessors__1__count": 2.5},
        "offset_negative": {"bufferViews__0__byteOffset": -4},
        "offset_huge": {"accessors__0__byteOffset": 1e300},
        "stride_wraps": {"bufferViews__0__byteStride": 2 ** 63, "accessors__0__count": 3},
        "stride_small": {"bufferViews__0__byteStride": 4},
        "view_past_buffer": {"bufferViews__1__byteLength": 400},
        "view_length_negative": {"bufferViews__1__byteLength": -3},
        "accessor_leaves_view": {"bufferViews__0__byteLength": 24},
        "buffer_index": {"bufferViews__0__buffer": 3},
        "view_index_negative": {"accessors__0__bufferView": -1},
        "child_negative": {"nodes__0__children": [-1]},
        "child_fraction": {"nodes__0__children": [0.5]},
        "child_out_of_range": {"nodes__0__children": [99]},
        "mesh_index": {"nodes__0__mesh": 1e12},
        "vertex_index": {"accessors__0__count": 2},                         # index 2 now points past POSITION
    }
    for name, patch in bad.items():
        with pytest.raises(rt.RtError) as e:
            load(doc(**patch), name + ".gltf")
        assert e.value.code == rt.RT_ERR_FORMAT, (name, str(e.value))
    with pytest.raises(rt.RtError) as e:                                    # 100 000 nested arrays: bounded recursion, clean error
        load("[" * 100000, "deep.gltf")
    assert e.value.code == rt.RT_ERR_FORMAT


def test_matrix_nodes_follow_the_gltf_crate_decomposition(rt, oracle, tmp_path):
    """`Transform::Matrix.decomposed()` of the gltf crate works in f32, gives a mirror to the z scale only
    (sz = signum(det) |z|) and converts with the cgmath branch order of Quaternion::from_matrix.  Pure rotations about every axis by
    angles that reach all four branches, a scaled rotation and MIRRORED matrices: the product loader equals the oracle loader bit for
    bit, and for pure rotations the normals come out rotated by that rotation."""
    import base64
    import json
    import struct
    pos = struct.pack("<9f", 0, 0, 0, 1, 0, 0, 0, 1, 0)
    nrm = struct.pack("<9f", 0, 0, 1, 0, 0, 1, 0, 0, 1)
    idx = struct.pack("<3B", 0, 1, 2) + b"\0"
    blob = pos + nrm + idx

    def rot(axis, ang):
        c, s = np.cos(ang), np.sin(ang)
        x, y, z = axis
        return np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                         [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                         [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])
    cases = [("rx_30", rot((1, 0, 0), 0.5), 1), ("rx_170", rot((1, 0, 0), 2.97), 1), ("ry_170", rot((0, 1, 0), 2.97), 1), ("rz_170", rot((0, 0, 1), 2.97), 1),
             ("scaled", rot((0.6, 0.0, 0.8), 1.1) @ np.diag([2.0, 3.0, 0.5]), 0), ("mirror", rot((0, 1, 0), 0.7) @ np.diag([1.0, 1.0, -1.0]), 0),
             ("mirror_x", rot((0, 0, 1), 0.3) @ np.diag([-2.0, 1.0, 1.0]), 0)]
    for name, R, pure in cases:
        M = np.eye(4); M[:3, :3] = R; M[:3, 3] = [0.5, -1.0, 2.0]
        g = {
            "asset": {"version": "2.0"},
            "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
            "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 36}, {"buffer": 0, "byteOffset": 72, "byteLength": 3}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC3"},
                          {"bufferView": 2, "componentType": 5121, "count": 3, "type": "SCALAR"}],
            "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1}, "indices": 2}]}],
            "cameras": [{"type": "perspective", "perspective": {"yfov": 0.5, "znear": 0.1}}],
            "nodes": [{"mesh": 0, "matrix": [float(v) for v in M.T.reshape(-1)]}, {"camera": 0, "translation": [0, 0, 5]}],
        }
        path = tmp_path / f"{name}.gltf"
        path.write_text(json.dumps(g))
        sc = rt.Scene.from_gltf(str(path), 8, 8, 1, device=-1)
        d = sc.desc()
        fl = oracle.convert_gltf_to_scene(str(path), 8, 8, 1)
        assert np.array_equal(d["tri_v"], fl.tri_v) and np.array_equal(d["tri_n"], fl.tri_n), name
        n = d["tri_n"][0, :3]
        assert abs(np.linalg.norm(n) - 1) < 1e-6, name
        if pure:
            assert np.allclose(n, R @ [0, 0, 1], atol=2e-6), (name, n, R @ [0, 0, 1])
        assert np.allclose(d["tri_v"][0, 3:6], R @ [1, 0, 0] + [0.5, -1.0, 2.0], atol=1e-6), name     # positions always go through the full matrix
        sc.close()
