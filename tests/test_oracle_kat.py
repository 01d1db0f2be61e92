"""Oracle pinned against the hand-derived known answers of SURVEY.md A.3 (f64 restatements of the cited reference
formulas) and against the committed kat.json (regression of the oracle itself).  Parity status: unpinned -- the
reference ships no golden vectors for this path."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, scene_path

nz = lambda v: np.array(v, float) / np.linalg.norm(v)  # noqa: E731
V, L, N, L2 = nz([-1, 0, 1]), nz([1, 0, 1]), np.array([0.0, 0.0, 1.0]), nz([0.3, 0.4, 0.8])


def test_specular_brdf_kat(oracle):
    # rendering.rs:157-184, the "good" case of the reference's own test_metalic_brdf (rendering.rs:186-199)
    assert oracle.specular_brdf([L], [N], [V], [N], [0.05])[0] == pytest.approx(25464.711317661277, rel=1e-12)


def test_brdf_kats(oracle):
    assert np.allclose(oracle.brdf([L], [N], [V], [[1, 0.25, 0.125, 0, 0.5]])[0], [0.40882489, 0.18013577, 0.14202092], rtol=1e-7)
    assert np.allclose(oracle.brdf([L], [N], [V], [[1, 1, 1, 1, 0.03]])[0], [196487.50450496] * 3, rtol=1e-10)
    assert np.allclose(oracle.brdf([L2], [N], [V], [[0.8, 0.2, 0.2, 0, 1]])[0], [0.24847644, 0.06515668, 0.06515668], rtol=1e-7)


def test_pdf_kats(oracle):
    assert oracle.pdf_vndf([N], [L], [V], [0.5])[0] == pytest.approx(1.7733440536689282, rel=1e-12)
    assert oracle.pdf_vndf([N], [L2], [V], [0.5])[0] == pytest.approx(0.22510628210609815, rel=1e-12)
    assert oracle.pdf_cosine([N], [L2])[0] == pytest.approx(0.26992624363190704, rel=1e-12)


def test_color_to_pixel_kat(oracle):
    # rendering.rs:250-262
    got = oracle.color_to_pixel([[0, 0, 0], [0.18, 0.18, 0.18], [0.5, 1, 2], [10, 10, 10], [0.01, 0.05, 0.25]])
    assert got.tolist() == [[0, 0, 0], [140, 140, 140], [205, 231, 245], [255, 255, 255], [20, 62, 163]]
    # NaN -> 0 (NaN.clamp stays NaN, `as u8` saturates NaN to 0); huge inputs saturate; the ACES fit maps x = -1 to
    # 2.48/1.98 > 1, so negative radiance (possible: SURVEY.md A.1 item 3) comes out white, like the reference
    assert oracle.color_to_pixel([[np.nan, -1.0, 1e30]]).tolist() == [[0, 255, 255]]
    assert oracle.color_to_pixel([[-0.001, -0.01, 0.0]]).tolist() == [[0, 0, 0]]


def test_scene_kats(oracle):
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_4"), 64, 64, 1)
    sc = oracle.OracleScene(fl)
    ray = sc.primary_rays([[0, 0]], [[0.5, 0.5]])[0]                     # rendering.rs:71-84
    assert np.allclose(ray[:3], [0, 0, 7])
    assert np.allclose(ray[3:], [-0.31681527, 0.31681527, -0.89401128], atol=1e-8)
    p = np.array([[0.0, -1.99999, 0.0]])
    cen = fl.tri_v[fl.light_ids[0]].reshape(3, 3).mean(0)
    d = (cen - p[0]) / np.linalg.norm(cen - p[0])
    assert sc.pdf_light(p, [d])[0] == pytest.approx(11.831245537422978, rel=1e-9)   # distributions.rs:160-184
    # the pdf ignores occlusion (the mirror sphere sits between that floor point and the light); an unoccluded
    # ray towards the same centroid must report the light triangle at its centroid
    p2 = np.array([[0.0, 1.5, 1.0]])
    d2 = (cen - p2[0]) / np.linalg.norm(cen - p2[0])
    hit = sc.trace_primary(np.concatenate([p2, [d2]], axis=1))
    assert hit["tri_id"][0] == fl.light_ids[0] and hit["t"][0] == pytest.approx(np.linalg.norm(cen - p2[0]), rel=1e-9)
    assert hit["u"][0] == pytest.approx(1 / 3, abs=1e-9) and hit["v"][0] == pytest.approx(1 / 3, abs=1e-9)


def test_committed_kat_json_matches_oracle(oracle):
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        k = json.load(f)
    assert oracle.specular_brdf([L], [N], [V], [N], [0.05])[0] == pytest.approx(k["specular_brdf_r005"], rel=1e-13)
    assert np.allclose(oracle.brdf([L], [N], [V], [[1, 0.25, 0.125, 0, 0.5]])[0], k["brdf_diel_r05"], rtol=1e-13)
    assert oracle.pdf_vndf([N], [L2], [V], [0.5])[0] == pytest.approx(k["vndf_pdf_l2"], rel=1e-13)
    assert oracle.color_to_pixel([[0.5, 1, 2]]).tolist()[0] == k["color_to_pixel"][2]


def test_xoshiro_and_splitmix(oracle):
    """xoshiro256** seeded through SplitMix64 (rand_xoshiro seed_from_u64), against an independent Python restatement
    and the published SplitMix64 outputs for seed 0."""
    M = (1 << 64) - 1

    def splitmix(state):
        state = (state + 0x9E3779B97F4A7C15) & M
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return state, z ^ (z >> 31)

    def rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & M

    for seed in (0, 1, 3840 * 2159, 2 ** 63 + 5):
        st, s = seed, []
        for _ in range(4):
            st, z = splitmix(st)
            s.append(z)
        if seed == 0:
            assert s == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F, 0xF88BB8A8724C81EC]
        exp = []
        for _ in range(8):
            exp.append((rotl((s[1] * 5) & M, 7) * 9) & M)
            t = (s[1] << 17) & M
            s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45)
        assert oracle.rng_u64(seed, 8).tolist() == exp


def test_triangle_and_aabb_primitives(oracle):
    # geometry.rs:93-138: inclusive edges, two-sided, t > 0
    tri = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    hit, tuv = oracle.intersect_triangle([0.25, 0.25, 1], [0, 0, -1], tri)
    assert hit and np.allclose(tuv, [1.0, 0.25, 0.25])
    assert oracle.intersect_triangle([0.25, 0.25, -1], [0, 0, 1], tri)[0]          # back face hits too
    assert oracle.intersect_triangle([0.5, 0.5, 1], [0, 0, -1], tri)[0]            # u+v == 1 edge is inclusive
    assert oracle.intersect_triangle([0.0, 0.3, 1], [0, 0, -1], tri)[0]            # u == 0 edge is inclusive
    assert not oracle.intersect_triangle([0.6, 0.6, 1], [0, 0, -1], tri)[0]
    assert not oracle.intersect_triangle([0.25, 0.25, -1], [0, 0, -1], tri)[0]     # behind the origin
    assert not oracle.intersect_triangle([0.25, 0.25, 1], [1, 0, 0], tri)[0]       # parallel: det == 0
    # bvh.rs:157-166 via geometry.rs:140-194: entry when outside, exit when inside, none when behind
    hit, t, outer = oracle.aabb_first_hit([0, 0, 5], [0, 0, -1], [-1, -1, -1], [1, 1, 1])
    assert hit and outer and t == pytest.approx(4.0, abs=1e-6)
    hit, t, outer = oracle.aabb_first_hit([0, 0, 0], [0, 0, -1], [-1, -1, -1], [1, 1, 1])
    assert hit and not outer and t == pytest.approx(1.0, abs=1e-6)
    assert not oracle.aabb_first_hit([0, 0, 5], [0, 0, 1], [-1, -1, -1], [1, 1, 1])[0]


@pytest.mark.parametrize("rough", [0.03, 0.2, 0.5, 1.0])
def test_device_vndf_construction_draws_the_reference_distribution(oracle, rough):
    """The device samples GGX visible normals with spherical caps in a Duff frame (oracle.sample_vndf_cap), the reference
    with Heitz 2018 in its helper-axis frame (distributions.rs:209-234, 264-274 = oracle.sample_vndf).  Same
    distribution: two-sample Kolmogorov-Smirnov on frame-independent projections of the reflected direction, and the
    importance-sampling identity E[g(l)/pdf(l)] = integral of g against the reference's own pdf (:276-297)."""
    from scipy import stats
    rng = np.random.default_rng(int(rough * 1000) + 7)
    cnt = 200_000
    alpha = rough * rough
    for n, v in [(N, nz([0.3, -0.2, 0.93])), (nz([0.5, -0.7, 0.2]), nz([0.9, -0.3, 0.6])), (nz([-0.2, 0.1, -0.97]), nz([0.1, 0.8, -0.4]))]:
        if np.dot(n, v) <= 0:
            v = nz(v - 2 * np.dot(n, v) * n)
        nn, vv, rr = np.tile(n, (cnt, 1)), np.tile(v, (cnt, 1)), np.full(cnt, rough)
        a = oracle.sample_vndf(nn, vv, rr, rng.random((cnt, 2)))           # reference construction
        b = oracle.sample_vndf_cap(nn, vv, rr, rng.random((cnt, 2)))       # device construction
        assert np.allclose(np.linalg.norm(b, axis=1), 1.0, atol=1e-12)
        # frame built from (n, v) only: e1 = v's tangential direction, e2 = n x e1
        e1 = nz(v - np.dot(n, v) * n); e2 = np.cross(n, e1)

        def proj(l):
            m = l + v; m /= np.linalg.norm(m, axis=1, keepdims=True)     # half vector = sampled visible normal
            mz = m @ n
            t2 = (1 - mz ** 2) / np.maximum(mz ** 2, 1e-300) / alpha ** 2  # tan^2(theta_m) / alpha^2: O(1) at every roughness
            return t2, np.arctan2(m @ e2, m @ e1), l @ n
        for pa, pb in zip(proj(a), proj(b)):
            assert stats.ks_2samp(pa, pb).pvalue > 1e-4
        # sampler <-> reference pdf: E_b[ 1{l.n > 0} cos / pdf ] must equal E_a[ the same ] (both estimate the same integral)
        def est(l):
            p = oracle.pdf_vndf(nn, l, vv, rr)
            w = np.where((l @ n > 0) & (p > 0), np.maximum(l @ n, 0) / np.where(p > 0, p, 1), 0.0)
            return w.mean(), w.std() / np.sqrt(cnt)
        (ma, sa), (mb, sb) = est(a), est(b)
        assert abs(ma - mb) < 5 * np.hypot(sa, sb) + 1e-12, (rough, ma, mb)
