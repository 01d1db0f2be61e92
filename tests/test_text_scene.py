"""CPU tests of the general-primitive / text-scene row: the product's C++ text loader against the oracle's
parser, and the ORACLE's restatement of the parts the reference still has -- Shape3D::Box (geometry.rs:140-194), the Object3D
transform (geometry.rs:196-251), object AABBs (aabb.rs:53-94), the box arm of the light sampler / pdf (distributions.rs:70-148)
-- pinned by hand-derived known answers and by the pdf-normalisation methodology of the reference's own (commented-out) tests
`test_light_box_distribution_norotation` / `test_light_box_distribution` (tests.rs:56-85), made two-sided and seeded.
PLANE / ELLIPSOID / DIELECTRIC are this repository's own specification: their tests pin the spec."""
import math
import os

import numpy as np
import pytest

from conftest import SCENES

TEXT_SCENES = ["practice3_1", "practice3_2", "practice3_3", "practice3_4", "practice3_5", "working"]
EPS = 1e-5


def text_path(name):
    return os.path.join(SCENES, name + ".txt")


# ------------------------------------------------------------------------------------------------ loader
@pytest.mark.parametrize("name", TEXT_SCENES)
def test_cpp_text_loader_matches_the_oracle_parser(rt, oracle, name):
    """rt_scene_load_text (csrc/text_loader.cpp) == oracle/text_ref.py value by value (both parse with correctly rounded
    strtod / float() and apply the same arithmetic), through a HOST-ONLY scene: no GPU needed, no compute call made."""
    fl = oracle.parse_text_scene(text_path(name))
    sc = rt.Scene.from_text(text_path(name), device=-1)
    d = sc.desc()
    for k in ("width", "height", "samples", "ray_depth", "camera_fov_x", "camera_fov_y"):
        assert d[k] == getattr(fl, k), k
    for k in ("bg_color", "camera_position", "camera_forward", "camera_right", "camera_up"):
        assert np.array_equal(d[k], getattr(fl, k)), k
    for k, a in (("tri_v", "tri_v"), ("tri_n", "tri_n"), ("tri_material", "tri_material"), ("tri_emission", "tri_emission"), ("shape_kind", "kind"),
                 ("position", "position"), ("rotation", "rotation"), ("ior", "ior"), ("material_kind", "mat_kind")):
        assert np.array_equal(d[k], getattr(fl, a)), k
    info = sc.info()
    n_planes = int((fl.kind == 3).sum())
    assert info["general_primitives"] == 1 and info["n_infinite"] == n_planes and info["n_tris"] == fl.n_tris - n_planes
    assert info["n_lights"] == len(fl.light_ids) and info["bvh_validate_failures"] == 0
    sc.close()


def test_text_scene_facts(oracle):
    """What the shipped files contain."""
    a = oracle.parse_text_scene(text_path("practice3_1"))
    assert (a.width, a.height, a.samples, a.ray_depth) == (640, 480, 64, 6) and a.kind.tolist() == [3, 2, 1] and np.allclose(a.bg_color, 1)
    assert len(a.light_ids) == 0
    assert a.camera_fov_y == pytest.approx(2 * math.atan(math.tan(1.54857776 / 2) * 480 / 640), abs=1e-15)
    b = oracle.parse_text_scene(text_path("practice3_5"))
    assert (b.width, b.height, b.samples) == (512, 512, 64) and b.kind.tolist() == [3, 3, 3, 3, 3, 1, 1, 2] and b.light_ids.tolist() == [5]
    assert np.allclose(b.tri_emission[5], 2) and np.allclose(b.tri_v[5, :3], [2, 0.1, 2])
    assert np.allclose(b.rotation[6], [0, math.sin(math.pi / 8), 0, math.cos(math.pi / 8)], atol=1e-7)       # 45 degrees about y
    assert abs(np.linalg.norm(b.rotation[6]) - 1) < 1e-15                                                       # normalised on load
    c = oracle.parse_text_scene(text_path("practice3_3"))
    assert c.tri_material[6].tolist() == [0.75, 0.75, 0.75, 1.0, 0.03] and c.kind[5] == 2 and c.light_ids.tolist() == [5]   # METALLIC; ellipsoid light
    d = oracle.parse_text_scene(text_path("practice3_4"))
    assert d.mat_kind.tolist() == [0, 0, 0, 0, 0, 0, 1] and d.ior[6] == 1.5
    w = oracle.parse_text_scene(text_path("working"))
    assert w.n_tris == 1379 and np.bincount(w.kind).tolist() == [505, 446, 423, 5] and int(w.mat_kind.sum()) == 469
    assert int((w.tri_material[:, 3] == 1).sum()) == 431
    tri = np.nonzero(w.kind == 0)[0][0]
    v = w.tri_v[tri].reshape(3, 3)
    ng = np.cross(v[1] - v[0], v[2] - v[0]); ng /= np.linalg.norm(ng)
    assert np.allclose(w.tri_n[tri].reshape(3, 3), ng)                                                           # face normal at the three vertices
    # CLI-style overrides (main.rs:39-41)
    o = oracle.parse_text_scene(text_path("practice3_1"), 100, 50, 7)
    assert (o.width, o.height, o.samples) == (100, 50, 7) and o.camera_fov_y == pytest.approx(2 * math.atan(math.tan(1.54857776 / 2) * 0.5))


def test_text_loader_errors(rt, tmp_path):
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_text(str(tmp_path / "missing.txt"), device=-1)
    assert e.value.code == rt.RT_ERR_IO
    bad = tmp_path / "bad.txt"
    bad.write_text("DIMENSIONS 8 8\nNEW_PRIMITIVE\nBOX 1 2\n")
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_text(str(bad), device=-1)
    assert e.value.code == rt.RT_ERR_FORMAT and "bad.txt:3" in str(e.value)
    nodim = tmp_path / "nodim.txt"
    nodim.write_text("SAMPLES 4\n")
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_text(str(nodim), device=-1)
    assert e.value.code == rt.RT_ERR_FORMAT
    ok = tmp_path / "ok.txt"
    ok.write_text("DIMENSIONS 8 4\nUNKNOWN_KEY 1 2 3\nNEW_PRIMITIVE\nCOLOR 1 0 0\n\nNEW_PRIMITIVE\nELLIPSOID 1 1 1\n")   # shapeless primitive dropped
    sc = rt.Scene.from_file(str(ok), device=-1)
    assert sc.info()["n_tris"] == 1 and sc.desc()["samples"] == 1
    sc.close()


def test_create2_validation(rt):
    kw = dict(width=8, height=8, samples=1, ray_depth=6, bg_color=[0, 0, 0], camera_position=[0, 0, 5], camera_forward=[0, 0, -1], camera_right=[1, 0, 0],
              camera_up=[0, 1, 0], camera_fov_x=1.0, camera_fov_y=1.0, tri_v=np.array([[1.0, 1, 1, 0, 0, 0, 0, 0, 0]]), tri_n=np.zeros((1, 9)),
              tri_material=np.array([[1.0, 1, 1, 0, 1]]), tri_emission=np.zeros((1, 3)), device=-1)
    for bad in (dict(shape_kind=[7]), dict(shape_kind=[1], rotation=[[0, 0, 0, 2.0]]), dict(shape_kind=[1], material_kind=[3]),
                dict(shape_kind=[1], material_kind=[1], ior=[0.0])):
        with pytest.raises(rt.RtError) as e:
            rt.Scene.from_arrays(**kw, **bad)
        assert e.value.code == rt.RT_ERR_INVALID
    sc = rt.Scene.from_arrays(**kw, shape_kind=[1], position=[[0, 0, 0]], rotation=[[0, 0, 0, 1.0]])
    assert sc.info()["general_primitives"] == 1 and sc.desc()["shape_kind"].tolist() == [1]
    sc.close()


# ------------------------------------------------------------------------------------------------ Shape3D::Box, Object3D (reference)
def test_box_intersection_known_answers(oracle):
    """geometry.rs:140-194 by hand: slabs with `d + 0.001*EPS` denominators, entry then exit, both normals facing the ray."""
    s = [1.0, 2.0, 3.0]
    hits = oracle.intersect_shape(oracle.SHAPE_BOX, s, [5, 0.5, 0.25], [-1, 0, 0])
    assert len(hits) == 2
    (t0, n0, o0), (t1, n1, o1) = hits
    assert t0 == (1.0 - 5.0) / (-1.0 + 0.001 * EPS) and t1 == (-1.0 - 5.0) / (-1.0 + 0.001 * EPS)                # 4.00000004, 6.00000006
    assert n0.tolist() == [1, 0, 0] and o0 and n1.tolist() == [1, 0, 0] and not o1                                # exit: -(signum(-1), 0, 0)
    # origin inside: only the exit hit, inner -> outer, normal pointing back inside
    (t, n, o), = oracle.intersect_shape(oracle.SHAPE_BOX, s, [0, 0, 0], [0, 0, 1])
    assert t == 3.0 / (1.0 + 0.001 * EPS) and n.tolist() == [0, 0, -1] and not o
    # upper bound cuts the list like geometry.rs:170,180 (0 < t < upper)
    assert len(oracle.intersect_shape(oracle.SHAPE_BOX, s, [5, 0.5, 0.25], [-1, 0, 0], upper=5.0)) == 1
    assert oracle.intersect_shape(oracle.SHAPE_BOX, s, [5, 0.5, 0.25], [-1, 0, 0], upper=3.0) == []
    # miss, and a box behind the origin
    assert oracle.intersect_shape(oracle.SHAPE_BOX, s, [5, 5, 0], [-1, 0, 0]) == []
    assert oracle.intersect_shape(oracle.SHAPE_BOX, s, [5, 0, 0], [1, 0, 0]) == []
    # face choice at an edge: the x, y, z cascade of :161-169 (a point within EPS of both the x and the y face reports x)
    d = np.array([-1.0, -1.0, 0.0]) / math.sqrt(2)
    (t, n, o), _ = oracle.intersect_shape(oracle.SHAPE_BOX, s, [1 + 3, 2 + 3, 0], d)
    assert n.tolist() == [1, 0, 0]
    # axis-parallel ray: the 1e-8 perturbation keeps the other slabs finite
    (t, n, o), _ = oracle.intersect_shape(oracle.SHAPE_BOX, s, [0.5, 0.5, 10], [0, 0, -1])
    assert t == pytest.approx(7.0, rel=1e-7) and n.tolist() == [0, 0, 1]


def test_quaternion_transform_and_object_aabb(oracle):
    """nalgebra UnitQuaternion * Vector3 and the 8-corner object AABB of aabb.rs:75-94."""
    c, s = math.cos(math.pi / 4), math.sin(math.pi / 4)
    q = [0, s, 0, c]                                                    # 90 degrees about +y
    assert np.allclose(oracle.quat_transform(q, [1, 0, 0]), [0, 0, -1], atol=1e-15)
    assert np.allclose(oracle.quat_transform(q, [0, 0, 1]), [1, 0, 0], atol=1e-15)
    assert np.allclose(oracle.quat_transform(q, [1, 0, 0], conjugate=True), [0, 0, 1], atol=1e-15)
    assert oracle.quat_transform([0, 0, 0, 1], [0.1, 0.2, 0.3]).tolist() == [0.1, 0.2, 0.3]      # identity: exact
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        v = rng.normal(size=3)
        i, j, k, w = q
        R = np.array([[1 - 2 * (j * j + k * k), 2 * (i * j - k * w), 2 * (i * k + j * w)], [2 * (i * j + k * w), 1 - 2 * (i * i + k * k), 2 * (j * k - i * w)],
                      [2 * (i * k - j * w), 2 * (j * k + i * w), 1 - 2 * (i * i + j * j)]])
        assert np.allclose(oracle.quat_transform(q, v), R @ v, atol=1e-14)
        assert np.allclose(oracle.quat_transform(q, oracle.quat_transform(q, v), conjugate=True), v, atol=1e-14)
    # practice3_5's rotated box: 45 degrees about y, s = (1.5, 3, 1.5), position (-2, -2, -1)
    q = np.array([0, 0.3826834, 0, 0.9238795]); q /= np.linalg.norm(q)
    mn, mx = oracle.object_aabb(oracle.SHAPE_BOX, [1.5, 3, 1.5], [-2, -2, -1], q)
    r = (1.5 + EPS) * math.sqrt(2)
    assert np.allclose(mn, [-2 - r, -2 - 3 - EPS, -1 - r], atol=1e-6) and np.allclose(mx, [-2 + r, -2 + 3 + EPS, -1 + r], atol=1e-6)
    # a triangle: padded bounds of the vertices when the object is not rotated
    mn, mx = oracle.object_aabb(oracle.SHAPE_TRIANGLE, [0, 0, 0, 1, 0, 0, 0, 2, 0], [1, 1, 1], [0, 0, 0, 1])
    assert np.allclose(mn, [1 - EPS, 1 - EPS, 1 - EPS], atol=1e-15) and np.allclose(mx, [2 + EPS, 3 + EPS, 1 + EPS], atol=1e-15)


def _one_light_scene(oracle, kind, s, position, rotation):
    """A scene whose only primitive is an emissive box / ellipsoid: MultipleLightSamplingDistribution over one light ==
    DirectLightSamplingDistribution (distributions.rs:150-185 with len 1)."""
    fl = oracle.FlatScene(width=4, height=4, samples=1)
    fl.camera_forward = np.array([0.0, 0, -1]); fl.camera_right = np.array([1.0, 0, 0]); fl.camera_up = np.array([0.0, 1, 0])
    fl.camera_fov_x = fl.camera_fov_y = 1.0
    fl.tri_v = np.array([list(s) + [0.0] * 6]); fl.tri_n = np.zeros((1, 9)); fl.tri_material = np.array([[1.0, 1, 1, 0, 1]]); fl.tri_emission = np.array([[1.0, 1, 1]])
    fl.kind = np.array([kind], dtype=np.int32); fl.position = np.array([position], dtype=float); fl.rotation = np.array([rotation], dtype=float)
    return fl


def _sphere(n, rng):
    x = rng.normal(size=(n, 3))
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("rotated", [False, True])
def test_light_box_pdf_normalisation(oracle, rotated):
    """tests.rs:56-85 (`test_light_box_distribution[_norotation]`, disabled in the reference): the box s = (1, 2, 3) at (0, 0, 4)
    seen from the origin; mean pdf over uniform sphere directions x 4 pi must be 1 -- here two-sided and with a fixed seed.
    The pdf counts BOTH pierced faces (entry and exit, geometry.rs:170-189), which is what makes a sampler that draws points on
    all six faces integrate to 1."""
    rng = np.random.default_rng(11)
    if rotated:
        q = rng.random(4); q /= np.linalg.norm(q)                       # tests.rs:71-76: Quaternion::new(rng.gen() x 4), normalised
    else:
        q = np.array([0.0, 0, 0, 1])
    osc = oracle.OracleScene(_one_light_scene(oracle, oracle.SHAPE_BOX, [1, 2, 3], [0, 0, 4], q))
    n = 2_000_000
    l = _sphere(n, rng)
    pdf = osc.pdf_light(np.zeros((n, 3)), l)
    est = pdf.mean() * 4 * math.pi
    err = pdf.std() / math.sqrt(n) * 4 * math.pi
    assert abs(est - 1.0) < max(5 * err, 0.01), (est, err)
    # sampler <-> pdf consistency: E_sampler[1 / pdf] = solid angle of the box = 4 pi x P(uniform direction hits it)
    m = 200_000
    draws = np.stack([rng.random(m), rng.choice([-1.0, 1.0], m), rng.random(m) * 2 - 1, rng.random(m) * 2 - 1], axis=1)
    ls = osc.sample_light(np.zeros(m, dtype=np.int32), np.zeros((m, 3)), draws)
    ps = osc.pdf_light(np.zeros((m, 3)), ls)
    assert (ps > 0).all()
    solid = (pdf > 0).mean() * 4 * math.pi
    assert (1.0 / ps).mean() == pytest.approx(solid, rel=0.02)


def test_box_light_sampler_draws_every_face_by_area(oracle):
    """distributions.rs:86-110: face pair chosen proportionally to its area, random sign, uniform inside the face."""
    osc = oracle.OracleScene(_one_light_scene(oracle, oracle.SHAPE_BOX, [1, 2, 3], [0, 0, 0], [0, 0, 0, 1]))
    rng = np.random.default_rng(5)
    m = 120_000
    draws = np.stack([rng.random(m), rng.choice([-1.0, 1.0], m), rng.random(m) * 2 - 1, rng.random(m) * 2 - 1], axis=1)
    far = np.tile([0.0, 0.0, 0.0], (m, 1))                              # from the centre: the direction IS the normalised surface point
    d = osc.sample_light(np.zeros(m, dtype=np.int32), far, draws)
    # recover the face from the direction: scale so that the point lies on the box
    p = d / np.max(np.abs(d) / np.array([1, 2, 3]), axis=1, keepdims=True)
    face = np.argmax(np.abs(p) / np.array([1, 2, 3]), axis=1)
    frac = np.bincount(face, minlength=3) / m
    w = np.array([2 * 3, 1 * 3, 1 * 2], dtype=float); w /= w.sum()
    assert np.allclose(frac, w, atol=0.005)
    for a in range(3):
        sel = face == a
        assert abs((p[sel, a] > 0).mean() - 0.5) < 0.01                 # gen_bool(0.5)
        others = [b for b in range(3) if b != a]
        for b in others:
            u = p[sel, b] / [1, 2, 3][b]
            assert abs(u.mean()) < 0.01 and abs(u.std() - 1 / math.sqrt(3)) < 0.01   # uniform on [-1, 1)


def test_nearest_hit_uses_the_first_hit_of_a_box_and_skips_nothing(oracle):
    """bvh.rs:268-276: a primitive competes with points.0[0] -- the entry hit from outside, the exit hit from inside."""
    fl = _one_light_scene(oracle, oracle.SHAPE_BOX, [1, 1, 1], [0, 0, 0], [0, 0, 0, 1])
    osc = oracle.OracleScene(fl)
    rays = np.array([[0, 0, 5, 0, 0, -1.0], [0, 0, 0.5, 0, 0, -1.0], [0, 0, 5, 0, 0, 1.0]])
    h = osc.trace_hits(rays)
    assert h[0, 0] == pytest.approx(4.0, rel=1e-7) and h[0, 1:4].tolist() == [0, 0, 1] and h[0, 8] == 1
    assert h[1, 0] == pytest.approx(1.5, rel=1e-7) and h[1, 1:4].tolist() == [0, 0, 1] and h[1, 8] == 0      # inside: exit face z = -1, normal flipped inward
    assert np.isinf(h[2, 0]) and h[2, 7] == -1


def test_rotated_object_keeps_normal_shading_in_object_space(oracle):
    """geometry.rs:245-249 rotates normal_geometry back to world space and leaves normal_shading alone -- restated, not fixed."""
    c, s = math.cos(math.pi / 4), math.sin(math.pi / 4)
    fl = _one_light_scene(oracle, oracle.SHAPE_BOX, [1, 1, 1], [0, 0, 0], [0, s, 0, c])     # 90 degrees about y: object +x = world -z
    h = oracle.OracleScene(fl).trace_hits(np.array([[0, 0, -5, 0, 0, 1.0]]))
    assert np.allclose(h[0, 1:4], [0, 0, -1], atol=1e-12)               # world-space geometric normal faces the ray
    assert np.allclose(h[0, 4:7], [1, 0, 0], atol=1e-12)                # shading normal still in object space


# ------------------------------------------------------------------------------------------------ own spec: plane, ellipsoid, dielectric
def test_plane_and_ellipsoid_known_answers(oracle):
    (t, n, o), = oracle.intersect_shape(oracle.SHAPE_PLANE, [0, 1, 0], [0, 2, 0], [0, -1, 0])
    assert t == 2.0 and n.tolist() == [0, 1, 0] and o
    (t, n, o), = oracle.intersect_shape(oracle.SHAPE_PLANE, [0, 1, 0], [0, -2, 0], [0, 1, 0])
    assert t == 2.0 and n.tolist() == [0, -1, 0] and not o               # from below: the normal faces the ray
    assert oracle.intersect_shape(oracle.SHAPE_PLANE, [0, 1, 0], [0, 2, 0], [1, 0, 0]) == []      # parallel
    assert oracle.intersect_shape(oracle.SHAPE_PLANE, [0, 1, 0], [0, 2, 0], [0, 1, 0]) == []      # behind
    hits = oracle.intersect_shape(oracle.SHAPE_ELLIPSOID, [2, 1, 1], [5, 0, 0], [-1, 0, 0])
    assert [h[0] for h in hits] == [3.0, 7.0] and hits[0][1].tolist() == [1, 0, 0] and hits[0][2] and hits[1][1].tolist() == [1, 0, 0] and not hits[1][2]
    (t, n, o), = oracle.intersect_shape(oracle.SHAPE_ELLIPSOID, [2, 1, 1], [0, 0, 0], [0, 1, 0])
    assert t == 1.0 and n.tolist() == [0, -1, 0] and not o
    # normal of an ellipsoid = normalize(p / r^2)
    d = np.array([-1.0, -0.1, -0.02]); d /= np.linalg.norm(d)
    (t, n, o), _ = oracle.intersect_shape(oracle.SHAPE_ELLIPSOID, [2, 1, 0.5], [5, 0.5, 0.1], d)
    p = np.array([5, 0.5, 0.1]) + t * d
    assert abs(((p / [2, 1, 0.5]) ** 2).sum() - 1) < 1e-12
    g = p / np.array([2, 1, 0.5]) ** 2
    assert np.allclose(n, g / np.linalg.norm(g), atol=1e-12)


@pytest.mark.parametrize("rotated", [False, True])
def test_light_ellipsoid_pdf_normalisation(oracle, rotated):
    """Own spec, same methodology: ellipsoid light radii (1, 2, 3) at (0, 0, 5) seen from the origin."""
    rng = np.random.default_rng(12)
    q = rng.random(4) if rotated else np.array([0.0, 0, 0, 1])
    q = q / np.linalg.norm(q)
    osc = oracle.OracleScene(_one_light_scene(oracle, oracle.SHAPE_ELLIPSOID, [1, 2, 3], [0, 0, 5], q))
    n = 2_000_000
    l = _sphere(n, rng)
    pdf = osc.pdf_light(np.zeros((n, 3)), l)
    est, err = pdf.mean() * 4 * math.pi, pdf.std() / math.sqrt(n) * 4 * math.pi
    assert abs(est - 1.0) < max(5 * err, 0.01), (est, err)
    m = 200_000
    ls = osc.sample_light(np.zeros(m, dtype=np.int32), np.zeros((m, 3)), _sphere(m, rng))
    ps = osc.pdf_light(np.zeros((m, 3)), ls)
    assert (ps > 0).all()
    assert (1.0 / ps).mean() == pytest.approx((pdf > 0).mean() * 4 * math.pi, rel=0.02)


def test_dielectric_direction_choice(oracle):
    n = np.array([[0, 0, 1.0]])
    v = np.array([[math.sin(0.5), 0, math.cos(0.5)]])
    # entering glass (eta = 1/1.5): Snell, reflect iff u < Schlick
    r0 = ((1 / 1.5 - 1) / (1 / 1.5 + 1)) ** 2
    R = r0 + (1 - r0) * (1 - math.cos(0.5)) ** 5
    refl = oracle.dielectric(n, v, [[1.5, 1, R * 0.999]])[0]
    refr = oracle.dielectric(n, v, [[1.5, 1, R * 1.001]])[0]
    assert refl[3] == 0 and np.allclose(refl[:3], [-math.sin(0.5), 0, math.cos(0.5)], atol=1e-15)
    sin2 = math.sin(0.5) / 1.5
    assert refr[3] == 1 and np.allclose(refr[:3], [-sin2, 0, -math.sqrt(1 - sin2 ** 2)], atol=1e-15)
    # leaving glass beyond the critical angle: total internal reflection whatever u is
    v2 = np.array([[math.sin(0.9), 0, math.cos(0.9)]])
    assert math.sin(0.9) * 1.5 > 1
    assert oracle.dielectric(n, v2, [[1.5, 0, 0.999999]])[0][3] == 0
    # and below it: refraction bends away from the normal
    v3 = np.array([[math.sin(0.3), 0, math.cos(0.3)]])
    out = oracle.dielectric(n, v3, [[1.5, 0, 0.999999]])[0]
    assert out[3] == 1 and out[0] == pytest.approx(-1.5 * math.sin(0.3), rel=1e-12)


def test_planes_are_scanned_after_the_bvh_with_strict_less(oracle):
    """rendering.rs:215-224: infinite primitives compete with the BVH hit under the running bound; nearest wins."""
    fl = oracle.parse_text_scene(text_path("practice3_1"))
    osc = oracle.OracleScene(fl)
    rays = np.array([[0, 2, 0, 0, -1, 0.0],            # straight down: the plane y = 0 at t = 2
                     [3, 10, -6, 0, -1, 0.0],          # down onto the box top (y = 4.5) before the plane
                     [0, 2, 0, 0, 1, 0.0]])            # up: nothing
    r = osc.trace_primary(rays)
    assert r["tri_id"].tolist() == [0, 2, -1]
    assert r["t"][0] == 2.0 and r["t"][1] == pytest.approx(5.5, rel=1e-7)
    assert r["second_t"][1] == pytest.approx(10.0)      # the plane behind the box


def test_oracle_render_of_text_scenes_is_deterministic_and_finite(oracle):
    fl = oracle.parse_text_scene(text_path("practice3_5"), 24, 24, 16)
    a = oracle.OracleScene(fl).render(seed=0, n_threads=2)
    b = oracle.OracleScene(fl).render(seed=0, n_threads=3)
    assert np.array_equal(a["rgb"], b["rgb"]) and np.isfinite(a["mean"]).all() and a["stats"]["nan_pixels"] == 0
    assert a["stats"]["segments"] == 24 * 24 * 16 * 6                    # a closed room: every path runs its 6 segments
    assert a["stats"]["light_tri_tests"] > 0
