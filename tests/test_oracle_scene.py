"""Oracle scene-level checks: loader facts, the BVH invariants the reference asserts at run time
(bvh.rs:31-35, 299-322), sampler pdf normalisation with the methodology of the reference's own test
(tests.rs:22-49, made two-sided and seeded) and render determinism."""
import numpy as np
import pytest

from conftest import scene_path

FACTS = {  # scene: (tris, light ids, nodes, leaves, depth)
    "practice7_1": (36, [10, 11], 25, 13, 8),
    "practice7_4": (92, [10, 11], 59, 30, 8),
}


@pytest.mark.parametrize("name", sorted(FACTS))
def test_scene_facts_and_bvh_invariants(oracle, name):
    tris, lights, nodes, leaves, depth = FACTS[name]
    fl = oracle.convert_gltf_to_scene(scene_path(name), 32, 32, 1)
    assert fl.n_tris == tris and fl.light_ids.tolist() == lights and fl.ray_depth == 6 and not fl.bg_color.any()
    sc = oracle.OracleScene(fl)
    info = sc.info()
    assert (info["n_nodes"], info["n_leaves"], info["depth"]) == (nodes, leaves, depth)
    assert info["n_light_nodes"] == 1 and info["validate_failures"] == 0           # validate_bvh, rendering.rs:22
    nd = sc.bvh_nodes()
    assert nd[-1, 8] == 0 and nd[-1, 9] == tris                                    # root is the LAST node (bvh.rs:31-35)
    leaf = nd[:, 6] < 0
    assert (nd[leaf, 9] <= 4).all()                                                 # bvh.rs:89
    assert sorted(sc.bvh_order().tolist()) == list(range(tris))


def test_large_scene_facts(oracle):
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_3"), 32, 32, 1)
    assert fl.n_tris == 99950 and fl.light_ids.tolist() == [10, 11]
    # roughness clamp (gltf_to_scene.rs:221) and emissive strength (:223-231)
    assert fl.tri_material[:, 4].min() >= 0.03
    assert np.allclose(fl.tri_emission[10], [5, 5, 5])


def test_practice7_4_material_and_camera_rules(oracle):
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_4"), 64, 64, 1)
    assert np.allclose(fl.camera_position, [0, 0, 7]) and np.allclose(fl.camera_forward, [0, 0, -1])
    assert fl.camera_fov_x == fl.camera_fov_y == pytest.approx(0.6911112070083618)  # aspect 1 -> fov_x = fov_y
    sphere = fl.tri_material[12:]
    assert np.allclose(sphere[:, :3], 1) and np.allclose(sphere[:, 3], 1) and np.allclose(sphere[:, 4], 0.03)
    assert np.allclose(fl.tri_emission[10:12], 10)


def _sphere(n, rng):
    x = rng.normal(size=(n, 3))
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("rough", [0.2, 0.5, 1.0])
def test_vndf_pdf_normalisation_uniform(oracle, rough):
    """tests.rs:22-49 methodology (mean pdf over uniform sphere directions * 4 pi), two-sided, fixed seed.
    The reference formula (distributions.rs:276-297) is also positive for half vectors BELOW the horizon, which the
    sampler never produces (max(0, Nh.z), :232): over half vectors with h.n > 0 it integrates to 1, over the whole
    sphere to more than 1 for rough surfaces (10/9 at roughness 1 for this v).  Both are pinned here."""
    rng = np.random.default_rng(7)
    n = np.array([0.0, 0.0, 1.0]); v = np.array([0.6, 0.0, 0.8])
    cnt = 400_000
    l = _sphere(cnt, rng)
    pdf = oracle.pdf_vndf(np.tile(n, (cnt, 1)), l, np.tile(v, (cnt, 1)), np.full(cnt, rough))
    h = l + v
    sampler_support = h[:, 2] > 0
    assert pdf[sampler_support].sum() / cnt * 4 * np.pi == pytest.approx(1.0, abs=0.03)
    total = pdf.sum() / cnt * 4 * np.pi
    assert total >= 1.0 - 0.03
    if rough == 1.0:
        assert total == pytest.approx(10.0 / 9.0, abs=0.02)


@pytest.mark.parametrize("rough", [0.03, 0.2, 0.5, 1.0])
def test_vndf_sampler_matches_pdf(oracle, rough):
    """Sampler <-> pdf consistency with an importance-sampled integrator (thin lobes): E_l~sampler[g(l)/pdf(l)] = int g."""
    rng = np.random.default_rng(11)
    n = np.array([0.0, 0.0, 1.0]); v = np.array([0.6, 0.0, 0.8])
    cnt = 200_000
    u = rng.random((cnt, 2))
    l = oracle.sample_vndf(np.tile(n, (cnt, 1)), np.tile(v, (cnt, 1)), np.full(cnt, rough), u)
    assert np.allclose(np.linalg.norm(l, axis=1), 1.0, atol=1e-9)
    pdf = oracle.pdf_vndf(np.tile(n, (cnt, 1)), l, np.tile(v, (cnt, 1)), np.full(cnt, rough))
    assert (pdf > 0).all()
    # g = cosine pdf (integrates to 1 over the upper hemisphere, 0 below): estimate int_{lobe support} g
    g = oracle.pdf_cosine(np.tile(n, (cnt, 1)), l)
    est = np.mean(g / pdf)
    if rough >= 0.5:
        assert est == pytest.approx(1.0, abs=0.05)
    else:
        assert 0.0 < est < 1.2      # thin lobe: g/pdf has most mass outside the lobe, only sanity


def test_cosine_sampler_and_pdf(oracle):
    rng = np.random.default_rng(3)
    cnt = 200_000
    n = _sphere(1, rng)[0]
    l = oracle.sample_cosine(np.tile(n, (cnt, 1)), _sphere(cnt, rng))
    assert (l @ n > -1e-12).all()
    assert np.mean(l @ n) == pytest.approx(2 / 3, abs=5e-3)                        # E[cos] of a cosine lobe
    u = _sphere(cnt, rng)
    assert np.mean(oracle.pdf_cosine(np.tile(n, (cnt, 1)), u)) * 4 * np.pi == pytest.approx(1.0, abs=0.01)


def test_light_sampler_matches_pdf(oracle):
    """Uniform-by-count light choice + area sampling (distributions.rs:111-125,151-158) vs the all-hits pdf (:160-184):
    E_l~sampler[1/pdf] = solid angle subtended by the lights."""
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_4"), 16, 16, 1)
    sc = oracle.OracleScene(fl)
    rng = np.random.default_rng(5)
    cnt = 100_000
    p = np.tile(np.array([0.3, -1.5, 0.4]), (cnt, 1))
    l = sc.sample_light(rng.integers(0, 2, cnt), p, rng.random((cnt, 2)))
    pdf = sc.pdf_light(p, l)
    assert (pdf > 0).all()
    omega_est = np.mean(1.0 / pdf)
    # solid angle by brute force: fraction of uniform directions that hit a light
    u = _sphere(400_000, rng)
    rays = np.concatenate([np.tile(p[0], (u.shape[0], 1)), u], axis=1)
    frac = (sc.pdf_light(rays[:, :3], rays[:, 3:]) > 0).mean()
    assert omega_est == pytest.approx(frac * 4 * np.pi, rel=0.05)


def test_render_is_deterministic_and_seedable(oracle):
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_1"), 24, 24, 8)
    sc = oracle.OracleScene(fl)
    a = sc.render(seed=0, n_threads=4)
    b = sc.render(seed=0, n_threads=2)
    c = sc.render(seed=1, n_threads=4)
    assert np.array_equal(a["mean"], b["mean"]) and np.array_equal(a["rgb"], b["rgb"])   # one stream per ROW (rendering.rs:50-51)
    assert not np.array_equal(a["mean"], c["mean"])
    st = a["stats"]
    assert st["samples"] == 24 * 24 * 8 and st["nan_pixels"] == 0 and st["vndf_assert_fail"] == 0
    assert 1.0 <= st["segments"] / st["samples"] <= 6.0                                # ray_depth = 6
    # row subset renders only those rows
    d = sc.render(seed=0, n_threads=2, rows=(4, 12, 4))
    assert np.array_equal(d["mean"][4], a["mean"][4]) and np.array_equal(d["mean"][8], a["mean"][8]) and not d["mean"][5].any()


def test_oracle_renders_are_bit_stable(oracle):
    """The oracle is the checker: its arithmetic must not drift when it is refactored (round 2 generalised it from triangles to
    Object3D primitives).  48x32x32-spp renders (seed 0 = the reference's per-row xoshiro seeding) are pinned by a hash of the
    f64 image and by two work counters; the practice7_* entries are the values the ROUND-1 oracle produced (checked against the
    round-1 sources when the generalisation landed), the text-scene entries pin the own-spec paths from round 2 on."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN, SCENES
    want = json.load(open(os.path.join(GOLDEN, "oracle_checksums.json")))
    for name, w in want.items():
        g = os.path.join(SCENES, name + ".gltf")
        fl = oracle.convert_gltf_to_scene(g, 48, 32, 32) if os.path.exists(g) else oracle.parse_text_scene(os.path.join(SCENES, name + ".txt"), 48, 32, 32)
        r = oracle.OracleScene(fl).render(seed=0, n_threads=3)
        got = {"sha16": hashlib.sha256(r["mean"].tobytes()).hexdigest()[:16], "attempts": r["stats"]["attempts"], "node_tests": r["stats"]["node_tests"]}
        if got != {k: w[k] for k in got}:
            # a different libm (sin / cos / log dispatch by CPU model) may flip a last bit and with it one accept / reject decision: then
            # the bit pin cannot hold on this host, but the render must still be the same estimator -- and the mismatch is reported
            import warnings
            warnings.warn(f"oracle bit pin differs on this host for {name}: {got} vs {w}")
            assert abs(float(r["mean"].mean()) - w["mean_radiance"]) <= 0.03 * w["mean_radiance"], (name, got, w)
            assert abs(got["attempts"] - w["attempts"]) <= 0.02 * w["attempts"] and abs(got["node_tests"] - w["node_tests"]) <= 0.02 * w["node_tests"]


def test_reference_own_test_vndf_in_its_exact_configuration(oracle):
    """The ONLY live test the reference ships for this path, restated verbatim: `tests::test_vndf` (tests.rs:43-49) calls
    `test_distribution(VndfDistribution, n = z, v = normalize(z + x) with x = Vec3f::z() (sic: so v = z), roughness 0.04)`
    (tests.rs:22-41): 1 000 000 uniform sphere directions l, avg = mean pdf(point 0, n, l, v, material), and asserts
    `(avg * 4 pi - 1.0) < 0.05` -- one-sided, with `thread_rng` (here: a fixed seed).  alpha = 0.0016 makes the lobe so thin
    that uniform sampling almost never resolves it, so the reference's assertion mostly checks "no blow-up"; it holds for the oracle.
    The two-sided statement (integral = 1) is checked next to it with an integrator that does resolve the lobe."""
    rng = np.random.default_rng(4349)
    cnt = 1_000_000
    x = rng.normal(size=(cnt, 3))
    l = x / np.linalg.norm(x, axis=1, keepdims=True)                    # random_unit_vec, tests.rs:11-20
    z = np.array([0.0, 0.0, 1.0])
    v = (z + z) / np.linalg.norm(z + z)                                 # tests.rs:45-47
    pdf = oracle.pdf_vndf(np.tile(z, (cnt, 1)), l, np.tile(v, (cnt, 1)), np.full(cnt, 0.04))
    avg = np.nanmean(np.where(np.isfinite(pdf), pdf, 0.0))
    assert (avg * 4 * np.pi - 1.0) < 0.05                               # tests.rs:40, as written
    # the same integral, resolved: polar angle of l around the mirror direction (= z) log-uniform in [1e-7, pi]
    L = np.log(np.pi / 1e-7)
    th = 1e-7 * np.exp(rng.random(cnt) * L); ph = rng.random(cnt) * 2 * np.pi
    l2 = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    p2 = oracle.pdf_vndf(np.tile(z, (cnt, 1)), l2, np.tile(v, (cnt, 1)), np.full(cnt, 0.04))
    above = (l2 + v) @ z > 0
    w = np.where(above & np.isfinite(p2), p2, 0.0) * (th * L * 2 * np.pi * np.sin(th))
    assert abs(w.mean() - 1.0) < max(5 * w.std() / np.sqrt(cnt), 0.01), w.mean()
