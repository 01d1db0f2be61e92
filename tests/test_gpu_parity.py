"""GPU parity tests (-m gpu): every check calls the CUDA path THROUGH THE C ABI (librt_b200.so) and compares with
the oracle (oracle/, f64 CPU restatement of the reference) on the same inputs, or with the committed golden
fixtures the oracle produced (tests/golden/make_golden.py).

Tolerances (north_star): primary-ray hit ids bit-exact except at exact ties, t within 1e-5 relative; converged
images within a stated RMSE / mean-luminance tolerance (Monte-Carlo noise; the RNG streams differ: Philox on
the device, xoshiro256** per row in the reference)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, scene_path

pytestmark = pytest.mark.gpu

M32 = 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ Philox
def philox4x32_10(ctr, key):
    c = [int(x) for x in ctr]
    k = [int(x) for x in key]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k[1]) & M32, p0 & M32]
        k = [(k[0] + 0x9E3779B9) & M32, (k[1] + 0xBB67AE85) & M32]
    return c


def test_philox_known_answers_and_device(gpu_rt):
    # Random123 kat_vectors, philox4x32-10
    assert philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox4x32_10([M32] * 4, [M32] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    rng = np.random.default_rng(0)
    x = rng.integers(0, 2 ** 32, size=(257, 4), dtype=np.uint64).astype(np.uint32)
    x[0] = 0
    out = gpu_rt.eval_fn(gpu_rt.FN_PHILOX, x.view(np.float32)).view(np.uint32)
    TAG = 0x52544232
    for i in range(x.shape[0]):
        assert out[i].tolist() == philox4x32_10([x[i, 0], x[i, 1], x[i, 2], TAG], [x[i, 3], 0])


# ------------------------------------------------------------------------------------------------ unit functions
def _unit(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def _shading_inputs(rng, cnt):
    n = _unit(rng, cnt)
    v = _unit(rng, cnt)
    v = np.where(((v * n).sum(1) < 0)[:, None], -v, v)
    v = v + 0.05 * n
    v /= np.linalg.norm(v, axis=1, keepdims=True)          # n.v > ~0.05
    l = _unit(rng, cnt)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    n, v, l = f32(n), f32(v), f32(l)
    # renormalise in f64 AFTER rounding so both sides see the same (almost unit) vectors
    return n, v, l


def test_brdf_matches_oracle(gpu_rt, oracle):
    rng = np.random.default_rng(1)
    cnt = 20000
    n, v, l = _shading_inputs(rng, cnt)
    base = rng.random((cnt, 3)).astype(np.float32)
    metallic = rng.choice([0.0, 1.0, 0.5, 0.25], size=cnt).astype(np.float32)
    rough = rng.choice([0.03, 0.1, 0.2, 0.35, 0.5, 1.0], size=cnt).astype(np.float32)
    x = np.concatenate([l, n, v, base, metallic[:, None], rough[:, None]], axis=1)
    got = gpu_rt.eval_fn(gpu_rt.FN_BRDF, x).astype(np.float64)
    ref = oracle.brdf(l, n, v, np.concatenate([base, metallic[:, None], rough[:, None]], axis=1))
    ok = np.isfinite(ref).all(1)
    scale = np.abs(ref[ok]).max(1, keepdims=True) + 1e-6
    err = np.abs(got[ok] - ref[ok]) / scale
    tol = np.where(rough[ok] < 0.1, 5e-3, 5e-4)[:, None]           # thin lobes amplify FP32 rounding of n x h
    assert (err <= tol).mean() > 0.999, float(err.max())
    assert np.median(err) < 2e-5


def test_pdfs_match_oracle(gpu_rt, oracle):
    rng = np.random.default_rng(2)
    cnt = 20000
    n, v, l = _shading_inputs(rng, cnt)
    rough = rng.choice([0.03, 0.2, 0.5, 1.0], size=cnt).astype(np.float32)
    got = gpu_rt.eval_fn(gpu_rt.FN_PDF_COSINE, np.concatenate([n, l], axis=1))[:, 0]
    assert np.allclose(got, oracle.pdf_cosine(n, l), rtol=2e-6, atol=1e-7)
    got = gpu_rt.eval_fn(gpu_rt.FN_PDF_VNDF, np.concatenate([n, l, v, rough[:, None]], axis=1))[:, 0].astype(np.float64)
    ref = oracle.pdf_vndf(n, l, v, rough)
    err = np.abs(got - ref) / (np.abs(ref) + 1e-9)
    tol = np.where(rough < 0.1, 5e-3, 5e-4)
    assert (err <= tol).mean() > 0.999, float(err.max())


def test_samplers_match_oracle(gpu_rt, oracle):
    rng = np.random.default_rng(3)
    cnt = 20000
    n, v, _ = _shading_inputs(rng, cnt)
    u = rng.random((cnt, 2)).astype(np.float32)
    out = gpu_rt.eval_fn(gpu_rt.FN_SAMPLE_COSINE, np.concatenate([n, u], axis=1)).astype(np.float64)
    sphere = out[:, 3:6]
    assert np.allclose(np.linalg.norm(sphere, axis=1), 1.0, atol=1e-5)
    # uniform on the sphere: z = 1 - 2 u1 exactly, mean ~ 0
    assert np.allclose(sphere[:, 2], 1 - 2 * u[:, 0].astype(np.float64), atol=1e-6)
    ref = oracle.sample_cosine(n, sphere)                            # distributions.rs:62 on the device's sphere point
    good = np.linalg.norm(sphere + n, axis=1) > 1e-2
    assert np.allclose(out[good, :3], ref[good], atol=2e-5)
    for rough in (0.03, 0.2, 0.5, 1.0):
        r = np.full((cnt, 1), rough, dtype=np.float32)
        got = gpu_rt.eval_fn(gpu_rt.FN_SAMPLE_VNDF, np.concatenate([n, v, r, u], axis=1)).astype(np.float64)
        ref = oracle.sample_vndf_cap(n, v, r[:, 0], u)                  # f64 restatement of the device's (cap) construction
        d = np.linalg.norm(got - ref, axis=1)
        assert np.quantile(d, 0.999) < 5e-4 and np.median(d) < 5e-6, (rough, d.max())


def test_light_sampler_and_pdf_match_oracle(gpu_rt, oracle):
    name = "practice7_4"
    sc = gpu_rt.Scene.from_gltf(scene_path(name), 16, 16, 1)
    osc = oracle.OracleScene(oracle.convert_gltf_to_scene(scene_path(name), 16, 16, 1))
    rng = np.random.default_rng(4)
    cnt = 20000
    p = (rng.random((cnt, 3)) * 3.6 - 1.8).astype(np.float32)
    idx = rng.integers(0, 2, cnt)
    uv = rng.random((cnt, 2)).astype(np.float32)
    got = sc.eval(gpu_rt.FN_SAMPLE_LIGHT, np.concatenate([p, idx[:, None].astype(np.float32), uv], axis=1)).astype(np.float64)
    ref = osc.sample_light(idx, p, uv)
    assert np.allclose(got, ref, atol=3e-6)
    l = np.ascontiguousarray(got, dtype=np.float32)
    gp = sc.eval(gpu_rt.FN_PDF_LIGHT, np.concatenate([p, l], axis=1))[:, 0].astype(np.float64)
    rp = osc.pdf_light(p, l)
    both = (gp > 0) & (rp > 0)
    assert both.mean() > 0.995                                      # edge samples may fall just outside in one of the two
    assert np.allclose(gp[both], rp[both], rtol=2e-3)
    # directions that miss the lights: pdf 0 on both sides
    l2 = np.ascontiguousarray(_unit(rng, cnt), dtype=np.float32)
    g2 = sc.eval(gpu_rt.FN_PDF_LIGHT, np.concatenate([p, l2], axis=1))[:, 0]
    r2 = osc.pdf_light(p, l2)
    assert ((g2 > 0) == (r2 > 0)).mean() > 0.999
    # mixture pdf (distributions.rs:194-201)
    n, v, _ = _shading_inputs(rng, cnt)
    rough = np.full((cnt, 1), 0.5, dtype=np.float32)
    gm = sc.eval(gpu_rt.FN_PDF_MIX, np.concatenate([p, n, l2, v, rough], axis=1))[:, 0].astype(np.float64)
    mat = np.concatenate([np.ones((cnt, 3)), np.zeros((cnt, 1)), rough], axis=1)
    rm = osc.pdf_mix(p, n, l2, v, mat)
    agree = (g2 > 0) == (r2 > 0)
    assert np.allclose(gm[agree], rm[agree], rtol=2e-3, atol=1e-7)
    sc.close()


# ------------------------------------------------------------------------------------------------ primary hits
def _centre_rays(osc, W, H, step=1):
    xs, ys = np.meshgrid(np.arange(0, W, step), np.arange(0, H, step))
    xy = np.stack([xs.ravel(), ys.ravel()], axis=1).astype(np.int32)
    return xy, osc.primary_rays(xy, np.full((xy.shape[0], 2), 0.5))


@pytest.mark.parametrize("name,W,H,step", [("practice7_1", 512, 512, 1), ("practice7_4", 512, 512, 1), ("practice7_4", 3840, 2160, 4),
                                          ("practice7_2", 512, 512, 2), ("practice7_3", 512, 512, 2)])
def test_primary_hit_parity(gpu_rt, oracle, name, W, H, step):
    """SURVEY.md 8d parity rule 1: ids equal on >= 99.99 % of the rays and on 100 % of the rays whose oracle
    barycentric margin > 1e-5 and whose gap to the second-nearest hit > 1e-5 t; |t_gpu - t_ref| / t_ref <= 1e-5."""
    fl = oracle.convert_gltf_to_scene(scene_path(name), W, H, 1)
    osc = oracle.OracleScene(fl)
    sc = gpu_rt.Scene.from_gltf(scene_path(name), W, H, 1)
    xy, rays = _centre_rays(osc, W, H, step)
    ref = osc.trace_primary(rays, want_second=False)
    for precision in (32, 64):
        tid, t = sc.trace_primary(rays, precision=precision)
        same = tid == ref["tri_id"]
        hit = same & (ref["tri_id"] >= 0)
        rel = np.abs(t[hit] - ref["t"][hit]) / ref["t"][hit]
        tol = 1e-5 if precision == 32 else 1e-12
        assert rel.max() <= tol, (precision, float(rel.max()))
        assert np.isinf(t[same & (ref["tri_id"] < 0)]).all()
        bad = np.nonzero(~same)[0]
        n_tie = n_one_sided = 0
        if bad.size:
            # SURVEY.md 8d, strictly: EVERY id mismatch must be a tie -- the oracle's hit sits on an edge (barycentric margin <= 1e-5)
            # or another triangle lies within 1e-5 t of it (second_t: brute force over all triangles, evaluated for the mismatches
            # only).  Pixel-centre rays of these symmetric scenes DO land exactly on shared edges (image diagonals = wall corners,
            # quad diagonals), so ties are not rare: ~0.1 % of the rays.  FP32 and f64 are held to the SAME rule.
            sub = osc.trace_primary(rays[bad], want_second=True)
            margin = np.minimum(np.minimum(sub["u"], sub["v"]), 1 - sub["u"] - sub["v"])
            gap = np.abs(sub["second_t"] - sub["t"])
            both = (sub["tri_id"] >= 0) & (tid[bad] >= 0)
            is_tie = both & ((margin <= 1e-5) | (gap <= 1e-5 * np.abs(sub["t"])))
            n_tie = int(is_tie.sum())
            assert (is_tie | ~both).all(), (name, precision, bad[both & ~is_tie][:10], margin[both & ~is_tie][:10])
            # one side hits, the other misses (silhouettes; the +2e-6 barycentric overlap of the FP32 tri_test): NOT a tie -- counted
            # separately and bounded: at most 1e-4 of the rays (ids equal on >= 99.99 %), and the hit that exists must be within 1e-4 of an edge
            one_sided = ~both
            n_one_sided = int(one_sided.sum())
            assert n_one_sided <= max(1, int(1e-4 * rays.shape[0])), (name, precision, n_one_sided)
            gpu_only = one_sided & (tid[bad] >= 0)
            oracle_only = one_sided & (sub["tri_id"] >= 0)
            assert (margin[oracle_only] <= 1e-4).all()
            if gpu_only.any():                                         # the GPU's extra hit: within 1e-4 of an edge of that triangle (oracle barycentrics of the same ray)
                for k in np.nonzero(gpu_only)[0]:
                    tri = fl.tri_v[tid[bad][k]].reshape(3, 3)
                    o, d = rays[bad][k, :3], rays[bad][k, 3:]
                    e1, e2 = tri[1] - tri[0], tri[2] - tri[0]
                    M = np.stack([e1, e2, -d], axis=1)
                    uvt = np.linalg.solve(M, o - tri[0])
                    assert min(uvt[0], uvt[1], 1 - uvt[0] - uvt[1]) >= -1e-4, (name, k, uvt)
        print(f"{name} {W}x{H} fp{precision}: id match {same.mean():.6f}, mismatches {bad.size} (ties {n_tie}, one-sided {n_one_sided}), max rel t err {rel.max():.2e}")
    # the device's own FP32 camera rays agree with the oracle's f64 rays (rendering.rs:71-84)
    dev_rays = sc.primary_rays(xy[:4096], np.full((min(4096, xy.shape[0]), 2), 0.5))
    assert np.allclose(dev_rays, rays[:4096], atol=2e-6)
    sc.close()


# ------------------------------------------------------------------------------------------------ converged images
def _lum(img):
    return img @ np.array([0.2126, 0.7152, 0.0722])


def _golden(name, W, H, spp):
    g = np.load(os.path.join(GOLDEN, f"converged_{name}_{W}x{H}_{spp}.npz"))
    return g["mean"].astype(np.float64), g["var"].astype(np.float64), json.loads(str(g["stats"]))


@pytest.mark.parametrize("name,W,H,ref_spp,gpu_spp", [("practice7_4", 64, 64, 16384, 16384), ("practice7_1", 64, 64, 16384, 16384),
                                                     ("practice7_4", 64, 36, 8192, 8192), ("practice7_2", 32, 32, 4096, 4096),
                                                     ("practice7_3", 32, 32, 4096, 4096)])
def test_converged_image_parity(gpu_rt, name, W, H, ref_spp, gpu_spp):
    """SURVEY.md 8d parity rule 2 against the committed oracle renders (tests/golden/): whole-frame mean luminance
    within 1 %, each block of a 4x4 grid within 2 % (or 4 sigma of its own Monte-Carlo noise), per-pixel RMSE within
    1.5x the noise floor predicted from the oracle's per-pixel sample variance."""
    ref, var, ost = _golden(name, W, H, ref_spp)
    sc = gpu_rt.Scene.from_gltf(scene_path(name), W, H, gpu_spp)
    img, st = sc.render_linear(seed=12345, collect_stats=True)
    img = img.astype(np.float64)
    assert np.isfinite(img).all()
    assert st["samples"] == W * H * gpu_spp and st["attempt_cap_hits"] <= 1e-6 * st["samples"] and st["nonfinite_samples"] <= 1e-6 * st["samples"]
    # same estimator -> same path statistics as the oracle (segments per sample, attempts per vertex)
    seg_o = ost["segments"] / ost["samples"]; att_o = ost["attempts"] / ost["vertices"]
    assert st["segments"] / st["samples"] == pytest.approx(seg_o, rel=0.01)
    assert st["attempts"] / st["vertices"] == pytest.approx(att_o, rel=0.01)
    lg, lr = _lum(img), _lum(ref)
    assert abs(lg.mean() - lr.mean()) / lr.mean() <= 0.01, (lg.mean(), lr.mean())
    lvar = _lum(var) * (1.0 / gpu_spp + 1.0 / ref_spp)              # upper bound of the luminance variance of the difference
    by, bx = H // 4, W // 4
    for j in range(4):
        for i in range(4):
            sl = (slice(j * by, (j + 1) * by), slice(i * bx, (i + 1) * bx))
            diff = abs(lg[sl].mean() - lr[sl].mean())
            sigma = np.sqrt(lvar[sl].sum()) / lg[sl].size
            assert diff <= max(0.02 * lr[sl].mean(), 4 * sigma), (name, j, i, diff, lr[sl].mean(), sigma)
    rmse = np.sqrt(np.mean((img - ref) ** 2))
    floor = np.sqrt(np.mean(var * (1.0 / gpu_spp + 1.0 / ref_spp)))
    assert rmse <= 1.5 * floor, (rmse, floor)
    sc.close()


def test_image_statistics_per_pixel_zscores(gpu_rt):
    """Per-pixel z-scores of (GPU - oracle) against the oracle's variance estimate: a biased estimator (a 'fixed'
    rejection loop, a missing emission term, wrong normals) shows up as a shifted or widened z distribution."""
    name, W, H, spp = "practice7_4", 64, 64, 16384
    ref, var, _ = _golden(name, W, H, spp)
    sc = gpu_rt.Scene.from_gltf(scene_path(name), W, H, spp)
    img = sc.render_linear(seed=777)[0].astype(np.float64)
    z = (_lum(img) - _lum(ref)) / np.sqrt(_lum(var) * 2.0 / spp + 1e-12)
    z = z[np.isfinite(z)]
    assert abs(np.median(z)) < 0.15 and abs(z.mean()) < 0.25, (np.median(z), z.mean())
    assert np.quantile(np.abs(z), 0.9) < 3.0
    sc.close()


def test_converged_frame_256x144_and_rgb8_criterion(gpu_rt, oracle):
    """VERDICT r1 item 2 / SURVEY.md 8d: practice7_4 at 256x144 (the 16:9 stretch of the north-star frame), 4 096 spp on both sides:
    (a) the linear-radiance gates of test_converged_image_parity on 147 456 pixels (16x9 grid of blocks), (b) on the RGB8 OUTPUT
    (rt_render bytes vs the oracle's color_to_pixel bytes) RMSE <= 2/255, mean signed error ~ 0."""
    name, W, H, spp = "practice7_4", 256, 144, 4096
    g = np.load(os.path.join(GOLDEN, f"converged_{name}_{W}x{H}_{spp}.npz"))
    ref, var, rgb_ref = g["mean"].astype(np.float64), g["var"].astype(np.float64), g["rgb"]
    sc = gpu_rt.Scene.from_gltf(scene_path(name), W, H, spp)
    img, st = sc.render_linear(seed=2024, collect_stats=True)
    img = img.astype(np.float64)
    assert st["samples"] == W * H * spp and st["attempt_cap_hits"] <= 1e-6 * st["samples"] and st["nonfinite_samples"] <= 1e-6 * st["samples"]
    lg, lr = _lum(img), _lum(ref)
    assert abs(lg.mean() - lr.mean()) / lr.mean() <= 0.005, (lg.mean(), lr.mean())
    lvar = _lum(var) * (2.0 / spp)
    for j in range(9):
        for i in range(16):
            sl = (slice(j * 16, (j + 1) * 16), slice(i * 16, (i + 1) * 16))
            diff = abs(lg[sl].mean() - lr[sl].mean())
            sigma = np.sqrt(lvar[sl].sum()) / lg[sl].size
            assert diff <= max(0.02 * lr[sl].mean(), 4.5 * sigma), (j, i, diff, lr[sl].mean(), sigma)
    rmse, floor = np.sqrt(np.mean((img - ref) ** 2)), np.sqrt(np.mean(var * (2.0 / spp)))
    assert rmse <= 1.5 * floor, (rmse, floor)
    # per-pixel differences are centred.  Two independent renders of the same estimator at the same spp have a SYMMETRIC difference
    # whatever the per-pixel distribution: the median is 0 and GPU > oracle on half of the pixels (sign test, 4 sigma).  (The mean of
    # z-scores built from the ORACLE's variance estimate is not symmetric: a pixel whose rare bright sample the oracle has not seen
    # gets a small sigma, so it is not used as a gate.)
    z = (lg - lr) / np.sqrt(lvar + 1e-12)
    assert abs(np.median(z)) < 0.05, float(np.median(z))
    nz = lg != lr
    assert abs((lg > lr)[nz].mean() - 0.5) < 4 * 0.5 / np.sqrt(nz.sum()), float((lg > lr)[nz].mean())
    # RGB8 output at this sample count: the difference to the oracle's bytes must be pure Monte-Carlo noise, i.e. no larger than the
    # difference between two independent ORACLE renders (seed 0 vs seed 1, committed as rgb / rgb2).  Measured: ~9/255 -- SURVEY.md
    # 8d's "RMSE <= 2/255 at >= 4 096 spp" is not reachable at 4 096 spp by ANY renderer of this estimator, the reference included;
    # test_rgb8_criterion below meets the 2/255 figure at the sample count it really needs.
    u8, _ = sc.render(seed=2024)
    d8 = u8.astype(np.float64) - rgb_ref.astype(np.float64)
    floor8 = np.sqrt(np.mean((g["rgb2"].astype(np.float64) - rgb_ref.astype(np.float64)) ** 2))
    rmse8 = np.sqrt(np.mean(d8 ** 2))
    assert rmse8 <= 1.15 * floor8, (rmse8, floor8)
    assert abs(d8.mean()) < 0.1, d8.mean()
    print(f"256x144x4096: linear rmse {rmse:.5f} (floor {floor:.5f}), RGB8 rmse {rmse8:.3f}/255 (oracle vs oracle {floor8:.3f}/255), mean signed {d8.mean():+.4f}/255")
    sc.close()


def test_rgb8_criterion(gpu_rt):
    """SURVEY.md 8d parity rule 2, last clause: "on the RGB8 output, RMSE <= 2/255 at >= 4 096 spp".  Two independent 4 096-spp renders
    of practice7_4 differ by ~9/255 (previous test), so the figure needs ~20x more samples per side: the committed oracle render of the
    16:9 frame at 64x36 with 131 072 spp against rt_render at 524 288 spp (1.2e9 camera paths, well under a second of B200 time)."""
    name, W, H, ospp, gspp = "practice7_4", 64, 36, 131072, 524288
    g = np.load(os.path.join(GOLDEN, f"converged_{name}_{W}x{H}_{ospp}.npz"))
    sc = gpu_rt.Scene.from_gltf(scene_path(name), W, H, gspp)
    u8, st = sc.render(seed=77, collect_stats=True)
    assert st["samples"] == W * H * gspp and st["attempt_cap_hits"] <= 1e-6 * st["samples"] and st["nonfinite_samples"] <= 1e-6 * st["samples"]
    d8 = u8.astype(np.float64) - g["rgb"].astype(np.float64)
    rmse8 = np.sqrt(np.mean(d8 ** 2))
    assert rmse8 <= 2.0, rmse8
    assert abs(d8.mean()) < 0.15, d8.mean()
    lin, _ = sc.render_linear(seed=77)
    ref = g["mean"].astype(np.float64)
    lg, lr = _lum(lin.astype(np.float64)), _lum(ref)
    assert abs(lg.mean() - lr.mean()) / lr.mean() < 0.002, (lg.mean(), lr.mean())
    print(f"64x36: GPU {gspp} spp vs oracle {ospp} spp: RGB8 rmse {rmse8:.3f}/255, mean signed {d8.mean():+.4f}/255, max |d| {np.abs(d8).max():.0f}")
    sc.close()


def _log_polar_directions(axis, n, rng, theta_min=1e-6):
    """Directions around `axis` with the polar angle log-uniform in [theta_min, pi]: an importance-sampled integrator that
    resolves a lobe of any width.  Returns (l, 1/q) with q the density per solid angle."""
    a = axis / np.linalg.norm(axis)
    t1 = np.cross(a, [0.0, 1.0, 0.0] if abs(a[1]) < 0.9 else [1.0, 0.0, 0.0]); t1 /= np.linalg.norm(t1)
    t2 = np.cross(a, t1)
    L = np.log(np.pi / theta_min)
    th = theta_min * np.exp(rng.random(n) * L)
    ph = rng.random(n) * 2 * np.pi
    l = np.cos(th)[:, None] * a + np.sin(th)[:, None] * (np.cos(ph)[:, None] * t1 + np.sin(ph)[:, None] * t2)
    inv_q = th * L * 2 * np.pi * np.sin(th)
    return l, inv_q


@pytest.mark.parametrize("rough", [0.03, 0.2, 0.5, 1.0])
def test_device_pdf_normalisation(gpu_rt, rough):
    """Methodology of the reference's own test (tests.rs:22-49: integral of the pdf over the sphere = 1), two-sided, fixed seed, with
    an importance-sampled integrator so that the alpha = 9e-4 lobe of roughness 0.03 is resolved, for each DEVICE pdf:
    cosine (distributions.rs:65-67), VNDF (:276-297; over half vectors above the horizon -- the reference's formula also gives mass to
    half vectors below it, which its sampler never draws, :232), lights (:160-184) and the mixture (:194-201)."""
    rng = np.random.default_rng(31)
    n = np.array([0.0, 0.0, 1.0]); v = np.array([0.6, 0.0, 0.8])
    cnt = 1_000_000
    mirror = 2 * n.dot(v) * n - v
    l, inv_q = _log_polar_directions(mirror, cnt, rng)
    l32 = np.ascontiguousarray(l, dtype=np.float32)
    l64 = l32.astype(np.float64); l64 /= np.linalg.norm(l64, axis=1, keepdims=True)
    N = np.tile(n.astype(np.float32), (cnt, 1)); V = np.tile(v.astype(np.float32), (cnt, 1)); R = np.full((cnt, 1), rough, dtype=np.float32)
    pv = gpu_rt.eval_fn(gpu_rt.FN_PDF_VNDF, np.concatenate([N, l32, V, R], axis=1))[:, 0].astype(np.float64)
    above = ((l64 + v) @ n) > 0
    w = np.where(above & np.isfinite(pv), pv, 0.0) * inv_q
    est, err = w.mean(), w.std() / np.sqrt(cnt)
    assert abs(est - 1.0) < max(5 * err, 0.01), ("vndf", rough, est, err)
    pc = gpu_rt.eval_fn(gpu_rt.FN_PDF_COSINE, np.concatenate([N, l32], axis=1))[:, 0].astype(np.float64)
    wc = pc * inv_q
    assert abs(wc.mean() - 1.0) < max(5 * wc.std() / np.sqrt(cnt), 0.01), ("cosine", wc.mean())
    # lights + mixture on practice7_4 from a point on the floor (two triangle lights in the ceiling corners)
    sc = gpu_rt.Scene.from_gltf(scene_path("practice7_4"), 16, 16, 1)
    P = np.tile(np.array([[0.3, -1.99, 0.2]], dtype=np.float32), (cnt, 1))
    up = np.array([0.0, 1.0, 0.0])
    vv = np.array([0.0, 0.8, 0.6])
    lu = rng.normal(size=(cnt, 3)); lu /= np.linalg.norm(lu, axis=1, keepdims=True)
    lu32 = np.ascontiguousarray(lu, dtype=np.float32)
    pl = sc.eval(gpu_rt.FN_PDF_LIGHT, np.concatenate([P, lu32], axis=1))[:, 0].astype(np.float64) * 4 * np.pi
    assert abs(pl.mean() - 1.0) < max(5 * pl.std() / np.sqrt(cnt), 0.01), ("light", pl.mean())
    # mixture = mean of the three (one-sample MIS, uniform component choice): integrate with the log-polar proposal around the mirror
    # direction for the VNDF part and uniformly for the rest -- two estimates of two different integrals, both must be right:
    #   uniform directions resolve cosine + lights (the VNDF lobe adds its share only for rough surfaces)
    if rough >= 0.2:
        Nn = np.tile(up.astype(np.float32), (cnt, 1)); Vv = np.tile(vv.astype(np.float32), (cnt, 1))
        pm = sc.eval(gpu_rt.FN_PDF_MIX, np.concatenate([P, Nn, lu32, Vv, R], axis=1))[:, 0].astype(np.float64) * 4 * np.pi
        pvu = gpu_rt.eval_fn(gpu_rt.FN_PDF_VNDF, np.concatenate([Nn, lu32, Vv, R], axis=1))[:, 0].astype(np.float64) * 4 * np.pi
        pcu = gpu_rt.eval_fn(gpu_rt.FN_PDF_COSINE, np.concatenate([Nn, lu32], axis=1))[:, 0].astype(np.float64) * 4 * np.pi
        assert np.allclose(pm, (pvu + pcu + pl) / 3.0, rtol=2e-3, atol=1e-6)
        above_u = ((lu + vv) @ up) > 0
        tot = np.where(above_u, pvu, 0.0) / 3 + pcu / 3 + pl / 3
        assert abs(tot.mean() - 1.0) < max(5 * tot.std() / np.sqrt(cnt), 0.02), ("mix", rough, tot.mean())
    sc.close()


# ------------------------------------------------------------------------------------------------ API behaviour
def test_render_u8_matches_oracle_tonemap_and_is_deterministic(gpu_rt, oracle):
    sc = gpu_rt.Scene.from_gltf(scene_path("practice7_4"), 96, 64, 64)
    a, st = sc.render(seed=5)
    b, _ = sc.render(seed=5)
    c, _ = sc.render(seed=6)
    assert a.shape == (64, 96, 3) and a.dtype == np.uint8
    assert np.array_equal(a, b)                                        # same seed -> identical bytes
    assert not np.array_equal(a, c)
    assert st["kernel_launches"] == 3 and st["kernel_ms"] > 0
    lin, _ = sc.render_linear(seed=5)
    exp = oracle.color_to_pixel(lin.reshape(-1, 3).astype(np.float64)).reshape(64, 96, 3)   # rendering.rs:250-262 in f64
    assert np.array_equal(a, exp)
    # global-memory scene variant == shared-memory scene variant
    g, _ = sc.render(seed=5, kernel_variant=1)
    s, _ = sc.render(seed=5, kernel_variant=2)
    assert np.abs(g.astype(int) - s.astype(int)).max() <= 1
    sc.close()


def test_kernel_variants_render_the_same_image(gpu_rt):
    """The per-lane megakernel (v1) and the warp-local wavefront (v2, 64 or 96 slots per warp) evaluate the same
    estimator with the same Philox counters (pixel, sample, call#): their images must agree to FP32 rounding."""
    W, H, spp = 96, 64, 128
    for name in ("practice7_4", "practice7_2"):
        sc = gpu_rt.Scene.from_gltf(scene_path(name), W, H, spp)
        ref, st1 = sc.render_linear(seed=21, kernel_variant=10, collect_stats=True)
        for kv in (20, 30, 21) + ((22,) if name == "practice7_4" else ()):   # placement 2 = shared memory: small scenes only
            img, st2 = sc.render_linear(seed=21, kernel_variant=kv, collect_stats=True)
            for k in ("samples", "segments", "vertices", "attempts"):
                assert abs(st2[k] - st1[k]) <= 1e-4 * st1[k], (name, kv, k, st1[k], st2[k])
            assert st2["attempt_cap_hits"] == 0 and st2["nonfinite_samples"] == 0
            # identical paths except where FP32 contraction differs by an ulp and flips a hit: compare robustly
            diff = np.abs(img.astype(np.float64) - ref.astype(np.float64))
            scale = np.abs(ref).mean()
            assert np.median(diff) <= 1e-5 * scale, (name, kv, float(np.median(diff)), float(scale))
            assert np.mean(diff) <= 2e-2 * scale, (name, kv, float(np.mean(diff)), float(scale))
        sc.close()


def test_sample_sharding_is_consistent(gpu_rt):
    """8e: rendering samples [0,s) in one call or as disjoint shards accumulated on the device gives the same image up
    to FP32 summation order -- the property the multi-GPU path relies on (counter-based RNG keyed by sample index)."""
    import torch
    W, H, spp = 80, 48, 96
    sc = gpu_rt.Scene.from_gltf(scene_path("practice7_1"), W, H, spp)
    full, _ = sc.render_linear(seed=9)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    acc = torch.zeros(H * W * 4, dtype=torch.float32, device="cuda")
    stream = ts.cuda_stream
    assert stream != 0
    for lo, hi in ((0, 32), (32, 33), (33, 96)):
        sc.render_accumulate_device(acc.data_ptr(), stream, seed=9, sample_begin=lo, sample_end=hi)
    torch.cuda.synchronize()
    a = acc.view(H, W, 4).cpu().numpy()
    assert np.all(a[..., 3] == spp)
    assert np.allclose(a[..., :3] / spp, full, rtol=2e-5, atol=1e-6)
    rgb = torch.zeros(H * W * 3, dtype=torch.uint8, device="cuda")
    gpu_rt.resolve_device(acc.data_ptr(), W, H, rgb.data_ptr(), stream)
    torch.cuda.synchronize()
    u8, _ = sc.render(seed=9)
    assert np.abs(rgb.view(H, W, 3).cpu().numpy().astype(int) - u8.astype(int)).max() <= 1
    sc.close()


def test_error_codes_on_gpu(gpu_rt):
    sc = gpu_rt.Scene.from_gltf(scene_path("practice7_1"), 8, 8, 0)
    with pytest.raises(gpu_rt.RtError) as e:                          # reference: panics on the empty reduce (rendering.rs:60)
        sc.render()
    assert e.value.code == gpu_rt.RT_ERR_INVALID
    sc.set_frame(8, 8, 4)
    with pytest.raises(gpu_rt.RtError):
        sc.render(sample_begin=3, sample_end=9)
    img, _ = sc.render()
    assert img.shape == (8, 8, 3)
    sc.close()
    # empty scene: every ray misses -> background (black) everywhere
    e0 = gpu_rt.Scene.from_arrays(width=8, height=8, samples=2, ray_depth=6, bg_color=[0.25, 0.5, 1.0], camera_position=[0, 0, 0], camera_forward=[0, 0, -1],
                                 camera_right=[1, 0, 0], camera_up=[0, 1, 0], camera_fov_x=1.0, camera_fov_y=1.0, tri_v=np.zeros((0, 9)),
                                 tri_n=np.zeros((0, 9)), tri_material=np.zeros((0, 5)), tri_emission=np.zeros((0, 3)))
    lin, _ = e0.render_linear()
    assert np.allclose(lin, [0.25, 0.5, 1.0])
    e0.close()


def test_ray_depth_zero_is_black_and_bad_depth_is_rejected(gpu_rt, oracle):
    """rendering.rs:93-95: recursion_depth <= 0 returns black before the scene is looked at -- also for the camera ray; a negative
    depth is an error here, not a wrapped bit field."""
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_1"), 24, 16, 4)
    fl.bg_color = np.array([0.3, 0.3, 0.3])
    for kv in (0, 10):
        sc = gpu_rt.Scene.from_flat(fl, ray_depth=0)
        lin, st = sc.render_linear(kernel_variant=kv)
        assert not lin.any() and st["samples"] == 24 * 16 * 4
        u8, _ = sc.render(kernel_variant=kv)
        assert not u8.any()
        sc.close()
    fl.ray_depth = 0
    assert not oracle.OracleScene(fl).render()["mean"].any()              # the oracle agrees
    sc = gpu_rt.Scene.from_flat(fl, ray_depth=-1)
    with pytest.raises(gpu_rt.RtError) as e:
        sc.render()
    assert e.value.code == gpu_rt.RT_ERR_INVALID
    sc.close()
    # max_attempts above the 7-bit attempt counter is clamped to 127 for every kernel: both kernels still render the same image
    sc = gpu_rt.Scene.from_gltf(scene_path("practice7_1"), 32, 32, 64)
    a, _ = sc.render_linear(seed=3, kernel_variant=10, max_attempts=1000)
    b, _ = sc.render_linear(seed=3, kernel_variant=30, max_attempts=1000)
    d = np.abs(a.astype(np.float64) - b)
    assert np.median(d) <= 1e-5 * np.abs(a).mean()
    sc.close()


def test_many_lights_use_the_light_bvh(gpu_rt, oracle):
    """More than 8 emissive triangles switch the pdf to the all-hits walk of the light BVH (bvh.rs:174-229)."""
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_4"), 32, 32, 512)
    emi = fl.tri_emission.copy()
    emi[12:40] = [0.5, 0.4, 0.3]                                     # part of the icosphere glows
    fl.tri_emission = emi
    osc = oracle.OracleScene(fl)
    assert osc.info()["n_lights"] == 30
    sc = gpu_rt.Scene.from_arrays(width=32, height=32, samples=2048, ray_depth=6, bg_color=fl.bg_color, camera_position=fl.camera_position,
                                  camera_forward=fl.camera_forward, camera_right=fl.camera_right, camera_up=fl.camera_up, camera_fov_x=fl.camera_fov_x,
                                  camera_fov_y=fl.camera_fov_y, tri_v=fl.tri_v, tri_n=fl.tri_n, tri_material=fl.tri_material, tri_emission=emi)
    rng = np.random.default_rng(8)
    cnt = 5000
    p = (rng.random((cnt, 3)) * 3.0 - 1.5).astype(np.float32); p[:, 1] = 1.5
    l = np.ascontiguousarray(_unit(rng, cnt), dtype=np.float32)
    gp = sc.eval(gpu_rt.FN_PDF_LIGHT, np.concatenate([p, l], axis=1))[:, 0].astype(np.float64)
    rp = osc.pdf_light(p, l)
    agree = (gp > 0) == (rp > 0)
    assert agree.mean() > 0.995 and np.allclose(gp[agree], rp[agree], rtol=3e-3, atol=1e-7)
    ref = osc.render(seed=0, n_threads=0, want_var=True)
    img = sc.render_linear(seed=3)[0].astype(np.float64)
    lg, lr = _lum(img).mean(), _lum(ref["mean"]).mean()
    assert abs(lg - lr) / lr < 0.03, (lg, lr)
    sc.close()


def test_cli_matches_library(gpu_rt, tmp_path):
    """raytracing-engine <scene> <w> <h> <samples> <out.ppm> (main.rs:37-43) == rt_render bytes behind the P6 header."""
    out = tmp_path / "o.ppm"
    r = subprocess.run([gpu_rt.CLI_PATH, scene_path("practice7_1"), "40", "24", "16", str(out), str(tmp_path / "pic")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Scene finite primitives: 36, light sources: 2" in r.stdout and "Rendering took" in r.stdout
    data = out.read_bytes()
    hdr = b"P6\n40 24\n255\n"
    assert data.startswith(hdr) and len(data) == len(hdr) + 40 * 24 * 3
    sc = gpu_rt.Scene.from_gltf(scene_path("practice7_1"), 40, 24, 16)
    img, _ = sc.render(seed=0)
    assert data[len(hdr):] == img.tobytes()
    from PIL import Image                                               # optional 6th argument: {arg}.png with the same pixels (main.rs:68-72)
    assert np.array_equal(np.asarray(Image.open(tmp_path / "pic.png").convert("RGB")), img)
    sc.close()


def test_non_black_background_matches_oracle_and_both_kernels(gpu_rt, oracle):
    """`None => scene.bg_color` (rendering.rs:125) with a non-black background (glTF scenes always load black,
    gltf_to_scene.rs:65; the flat-scene entry point takes any colour): escaping paths carry throughput x background."""
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_1"), 32, 32, 1024)
    fl.bg_color = np.array([0.4, 0.6, 0.9])
    osc = oracle.OracleScene(fl)
    ref = osc.render(seed=0, n_threads=0)
    sc = gpu_rt.Scene.from_arrays(width=32, height=32, samples=1024, ray_depth=6, bg_color=fl.bg_color, camera_position=fl.camera_position,
                                  camera_forward=fl.camera_forward, camera_right=fl.camera_right, camera_up=fl.camera_up, camera_fov_x=fl.camera_fov_x,
                                  camera_fov_y=fl.camera_fov_y, tri_v=fl.tri_v, tri_n=fl.tri_n, tri_material=fl.tri_material, tri_emission=fl.tri_emission)
    a, _ = sc.render_linear(seed=9, kernel_variant=10)
    b, _ = sc.render_linear(seed=9, kernel_variant=30)
    assert np.allclose(a, b, rtol=2e-3, atol=2e-4), float(np.abs(a - b).max())
    lg, lr = _lum(b.astype(np.float64)).mean(), _lum(ref["mean"]).mean()
    assert abs(lg - lr) / lr < 0.02, (lg, lr)
    black = gpu_rt.Scene.from_gltf(scene_path("practice7_1"), 32, 32, 1024)
    c, _ = black.render_linear(seed=9)
    assert _lum(b.astype(np.float64)).mean() > 1.05 * _lum(c.astype(np.float64)).mean()      # the background really contributes
    sc.close(); black.close()


@pytest.mark.parametrize("name", ["practice7_2", "practice7_4"])
def test_gpu_bvh_builder_gives_the_same_hits(gpu_rt, monkeypatch, name):
    """the GPU LBVH builder (RT_BVH_BUILDER=gpu; auto picks it for meshes of >= 1 000 000 triangles) replaces create_bvh_tree (bvh.rs:26-144).  Tree shape is
    not observable: the flattened tree must pass validate_bvh (bvh.rs:299-322 restated) and every primary ray must
    report the same triangle and distance as with the host SAH tree (f64 triangle tests: exact ties aside)."""
    W = H = 256
    monkeypatch.setenv("RT_BVH_BUILDER", "host")
    host = gpu_rt.Scene.from_gltf(scene_path(name), W, H, 16)
    monkeypatch.setenv("RT_BVH_BUILDER", "gpu")
    dev = gpu_rt.Scene.from_gltf(scene_path(name), W, H, 16)
    monkeypatch.delenv("RT_BVH_BUILDER")
    monkeypatch.setenv("RT_BVH_GPU_MIN_TRIS", "32768")                   # auto: the GPU builder from this many triangles on (default 1 000 000)
    auto = gpu_rt.Scene.from_gltf(scene_path(name), W, H, 16)
    assert auto.info()["bvh_builder"] == (1 if name == "practice7_2" else 0)
    auto.close()
    monkeypatch.delenv("RT_BVH_GPU_MIN_TRIS")
    ih, idv = host.info(), dev.info()
    assert ih["bvh_builder"] == 0 and idv["bvh_builder"] == 1
    assert idv["bvh_validate_failures"] == 0 and idv["max_leaf_size"] <= 4 and idv["n_tris"] == ih["n_tris"]
    if name == "practice7_2":                                            # the rebuild of the top (regraft_top_sah, twice) is on: far fewer box tests than the bare Morton tree
        monkeypatch.setenv("RT_BVH_BUILDER", "gpu"); monkeypatch.setenv("RT_BVH_TOP_SAH", "0"); monkeypatch.setenv("RT_BVH_TOP_AGGLO", "0")
        morton = gpu_rt.Scene.from_gltf(scene_path(name), 96, 96, 8)
        monkeypatch.delenv("RT_BVH_BUILDER"); monkeypatch.delenv("RT_BVH_TOP_SAH"); monkeypatch.delenv("RT_BVH_TOP_AGGLO")
        dev.set_frame(96, 96, 8)
        _, s_m = morton.render_linear(seed=4, collect_stats=True)
        _, s_r = dev.render_linear(seed=4, collect_stats=True)
        dev.set_frame(W, H, 16)
        assert s_r["node_tests"] < 0.80 * s_m["node_tests"], (s_r["node_tests"], s_m["node_tests"])
        assert morton.info()["bvh_validate_failures"] == 0
        morton.close()
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    xy = np.stack([xs.ravel(), ys.ravel()], axis=1).astype(np.int32)
    rays = host.primary_rays(xy, np.full((xy.shape[0], 2), 0.5))
    ia, ta = host.trace_primary(rays, precision=64)
    ib, tb = dev.trace_primary(rays, precision=64)
    same = ia == ib
    assert same.mean() > 0.995, same.mean()
    assert np.array_equal(ta, tb)                                       # same distance everywhere (same f64 triangle arithmetic) ...
    assert ((ia >= 0) == (ib >= 0)).all()                               # ... so an id mismatch is an exact tie: the ray runs along an edge
    #                                                                     shared by two triangles (the image diagonal = the back wall's diagonal)
    # the render goes through the same tree: same path statistics and image up to FP32 tie-breaks
    a, sa = host.render_linear(seed=4, collect_stats=True)
    b, sb = dev.render_linear(seed=4, collect_stats=True)
    for k in ("samples", "segments", "vertices"):
        assert abs(sa[k] - sb[k]) <= 2e-4 * sa[k], (k, sa[k], sb[k])
    assert abs(_lum(a.astype(np.float64)).mean() - _lum(b.astype(np.float64)).mean()) < 0.02 * _lum(a.astype(np.float64)).mean()
    print(f"{name}: host SAH {ih['bvh_build_ms']:.1f} ms / {ih['n_nodes']} nodes / depth {ih['bvh_depth']};  GPU LBVH {idv['bvh_build_ms']:.2f} ms / "
          f"{idv['n_nodes']} nodes / depth {idv['bvh_depth']};  box tests per segment {sa['node_tests'] / sa['segments']:.1f} vs {sb['node_tests'] / sb['segments']:.1f}")
    host.close(); dev.close()


@pytest.mark.parametrize("name", ["practice7_2", "practice7_3"])
def test_quantised_nodes_walk_like_the_octant_nodes(gpu_rt, monkeypatch, name):
    """Global-memory triangle scenes walk 32-byte nodes whose planes are 16-bit grid coordinates (rt_device.cuh pair_step_quant,
    rt_api.cu quant_nodes; RT_NODE_FORMAT=0 keeps the 112-byte octant nodes).  The quantised boxes are conservative, so the
    FP32 nearest hit must be the same triangle at the same distance -- for camera rays AND for rays that start up to 50 scene
    diameters outside the mesh (the slack of the grid planes covers the FP32 rounding of far origins) -- and the rendered frame
    must have the same path statistics."""
    W = H = 192
    monkeypatch.setenv("RT_NODE_FORMAT", "0")
    octant = gpu_rt.Scene.from_gltf(scene_path(name), W, H, 8)
    monkeypatch.setenv("RT_NODE_FORMAT", "2")
    quant = gpu_rt.Scene.from_gltf(scene_path(name), W, H, 8)
    monkeypatch.delenv("RT_NODE_FORMAT")
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    xy = np.stack([xs.ravel(), ys.ravel()], axis=1).astype(np.int32)
    cam = octant.primary_rays(xy, np.full((xy.shape[0], 2), 0.5))
    d = octant.desc()
    v = np.asarray(d["tri_v"], np.float64).reshape(-1, 3)
    lo, hi = v.min(axis=0), v.max(axis=0)
    rng = np.random.default_rng(11)
    n_far = 20000
    target = lo + rng.random((n_far, 3)) * (hi - lo)
    dirs = rng.normal(size=(n_far, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dist = np.linalg.norm(hi - lo) * rng.uniform(0.0, 50.0, size=(n_far, 1))
    far = np.concatenate([target - dirs * dist, dirs], axis=1)
    for rays, label in ((cam, "camera"), (far, "far origins")):
        ia, ta = octant.trace_primary(rays, precision=32)
        ib, tb = quant.trace_primary(rays, precision=32)
        assert ((ia >= 0) == (ib >= 0)).all(), label
        hit = ia >= 0
        assert hit.mean() > 0.3, (label, hit.mean())
        assert np.allclose(ta[hit], tb[hit], rtol=1e-6, atol=0.0), label            # same triangle arithmetic: a different id is a tie in t
        assert (ia == ib).mean() > 0.995, (label, (ia == ib).mean())                   # (visiting order differs where entry distances tie: shared edges)
    a, sa = octant.render_linear(seed=3, collect_stats=True)
    b, sb = quant.render_linear(seed=3, collect_stats=True)
    for k in ("samples", "segments", "vertices", "attempts"):
        assert abs(sa[k] - sb[k]) <= 2e-4 * sa[k], (k, sa[k], sb[k])
    assert sb["node_tests"] <= 1.02 * sa["node_tests"]                                # the grid costs < 2 % more box tests
    assert abs(_lum(a.astype(np.float64)).mean() - _lum(b.astype(np.float64)).mean()) < 0.01 * _lum(a.astype(np.float64)).mean()
    octant.close(); quant.close()


def test_full_size_frame_properties(gpu_rt):
    """BASELINE.json's headline frame size (3840x2160, practice7_4) at a low sample count, through size-independent
    properties: every pixel receives exactly `spp` samples; the same seed gives identical bytes; two sample shards
    accumulated on the device equal the single call up to FP32 summation order; no attempt cap / non-finite sample; the
    frame's mean luminance and its 16x9 block means agree with the committed converged 64x36 oracle render (same camera,
    `fov_x = aspect * fov_y` stretch included) within Monte-Carlo tolerance."""
    import torch
    W, H, spp = 3840, 2160, 8
    sc = gpu_rt.Scene.from_gltf(scene_path("practice7_4"), W, H, spp)
    _, st = sc.render_linear(seed=77, collect_stats=True)                # the instrumented build: counters only
    assert st["samples"] == W * H * spp and st["attempt_cap_hits"] == 0 and st["nonfinite_samples"] == 0
    lin, _ = sc.render_linear(seed=77)                                   # the product build (the one the shards below run)
    assert np.isfinite(lin).all()
    a8, _ = sc.render(seed=77)
    b8, _ = sc.render(seed=77)
    assert np.array_equal(a8, b8)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    acc = torch.zeros(H * W * 4, dtype=torch.float32, device="cuda")
    for lo, hi in ((0, 3), (3, 8)):
        sc.render_accumulate_device(acc.data_ptr(), ts.cuda_stream, seed=77, sample_begin=lo, sample_end=hi)
    torch.cuda.synchronize()
    a = acc.view(H, W, 4)
    assert bool((a[..., 3] == spp).all())
    shard = (a[..., :3] / spp).cpu().numpy()
    err = np.abs(shard.astype(np.float64) - lin)
    assert (err <= 2e-5 * np.abs(lin) + 1e-6).mean() > 0.999             # FP32 summation order only ...
    assert err.max() <= 1e-5 * max(1.0, float(np.abs(lin).max()))         # ... also where large terms of both signs cancel
    # (the instrumented and the product build are different compilations: an ulp may flip a hit, so they are not compared pixel by pixel)
    ref, var, _ = _golden("practice7_4", 64, 36, 8192)
    blocks = _lum(lin.astype(np.float64)).reshape(36, H // 36, 64, W // 64).mean(axis=(1, 3))     # 60x60-pixel blocks = the oracle's pixels
    lr = _lum(ref)
    assert abs(blocks.mean() - lr.mean()) / lr.mean() < 0.01, (blocks.mean(), lr.mean())
    coarse_g = blocks.reshape(9, 4, 16, 4).mean(axis=(1, 3)); coarse_r = lr.reshape(9, 4, 16, 4).mean(axis=(1, 3))
    # a 60x60 block of jittered samples covers the oracle pixel's footprint: equal in expectation.  4x4 groups of them within
    # 2 % or 5 sigma of the Monte-Carlo noise predicted from the oracle's per-pixel variance (highlights are heavy-tailed)
    lvar = _lum(var) * (1.0 / 8192 + 1.0 / (spp * (H // 36) * (W // 64)))
    sigma = np.sqrt(lvar.reshape(9, 4, 16, 4).sum(axis=(1, 3))) / 16.0
    assert np.all(np.abs(coarse_g - coarse_r) <= np.maximum(0.02 * coarse_r, 5 * sigma)), float(np.abs(coarse_g - coarse_r).max())
    sc.close()


def test_single_process_multi_gpu_render(gpu_rt, monkeypatch, tmp_path):
    """rt_render_multi (8e; north_star 'main.rs (device and multi-GPU orchestration)'): samples sharded over the GPUs of the
    box from one process, ONE reduce of the framebuffer (NCCL through dlopen, or peer copies with RT_NO_NCCL=1), resolve
    on device 0.  Same image as the single-GPU call up to FP32 summation order; needs >= 2 GPUs."""
    if gpu_rt.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(gpu_rt.device_count(), 4)
    W, H, spp = 256, 144, 66
    scenes = [gpu_rt.Scene.from_gltf(scene_path("practice7_4"), W, H, spp, device=g) for g in range(n)]
    one, st1 = scenes[0].render(seed=5, collect_stats=True)
    for no_nccl in ("0", "1"):
        monkeypatch.setenv("RT_NO_NCCL", no_nccl)
        multi, stn = gpu_rt.render_multi(scenes, seed=5, collect_stats=True)
        assert stn["samples"] == W * H * spp == st1["samples"]
        assert abs(stn["segments"] - st1["segments"]) <= 1e-4 * st1["segments"]
        d = np.abs(multi.astype(int) - one.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 0.01, (no_nccl, int(d.max()), float((d > 0).mean()))
    monkeypatch.delenv("RT_NO_NCCL")
    out = tmp_path / "m.ppm"
    env = dict(os.environ, RT_GPUS=str(n), RT_SEED="5")
    r = subprocess.run([gpu_rt.CLI_PATH, scene_path("practice7_4"), str(W), str(H), str(spp), str(out)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    hdr = f"P6\n{W} {H}\n255\n".encode()
    cli = np.frombuffer(out.read_bytes()[len(hdr):], dtype=np.uint8).reshape(H, W, 3)
    assert np.abs(cli.astype(int) - one.astype(int)).max() <= 1
    for s in scenes:
        s.close()


@pytest.mark.parametrize("kv", [10, 30])
def test_tile_sharding_is_consistent(gpu_rt, kv):
    """8e fallback (fewer samples than GPUs): interleaved 8x4-pixel tile shards accumulated on the device give the single-call
    image exactly (every pixel is rendered by exactly one shard with the same Philox counters), for both kernels."""
    import torch
    W, H, spp, n = 100, 61, 3, 4                                         # frame not a multiple of the tile size
    sc = gpu_rt.Scene.from_gltf(scene_path("practice7_4"), W, H, spp)
    full, _ = sc.render_linear(seed=3, kernel_variant=kv)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    acc = torch.zeros(H * W * 4, dtype=torch.float32, device="cuda")
    tot = 0
    for g in range(n):
        st = sc.render_accumulate_device(acc.data_ptr(), ts.cuda_stream, want_stats=True, seed=3, kernel_variant=kv, tile_shard=(g, n))
        tot += st["samples"]
    torch.cuda.synchronize()
    a = acc.view(H, W, 4).cpu().numpy()
    assert tot == W * H * spp and np.all(a[..., 3] == spp)
    assert np.array_equal(a[..., :3] / spp, full)
    part, _ = sc.render_linear(seed=3, kernel_variant=kv, tile_shard=(1, n))   # one shard alone: its tiles only, zeros elsewhere
    own = np.zeros((H, W), dtype=bool)
    for y in range(H):
        for x in range(W):
            own[y, x] = ((y // 4) * ((W + 7) // 8) + x // 8) % n == 1
    assert np.array_equal(part[own], full[own]) and not part[~own].any()
    sc.close()


def test_automatic_scene_placement_keeps_two_blocks_per_sm(gpu_rt, oracle):
    """A mid-size scene (184 triangles, ~40 KB blob) would fit the shared-memory staging limit but push the render block past
    half of the SM's shared memory; the automatic placement must then read the scene through L1/L2 instead (two resident
    blocks per SM), and both placements must render the same image."""
    fl = oracle.convert_gltf_to_scene(scene_path("practice7_4"), 96, 54, 64)
    far = fl.tri_v.copy(); far[:, 0::3] += 50.0                           # a second copy of the room, out of sight
    kw = dict(width=96, height=54, samples=64, ray_depth=6, bg_color=fl.bg_color, camera_position=fl.camera_position, camera_forward=fl.camera_forward,
              camera_right=fl.camera_right, camera_up=fl.camera_up, camera_fov_x=fl.camera_fov_x, camera_fov_y=fl.camera_fov_y,
              tri_v=np.concatenate([fl.tri_v, far]), tri_n=np.concatenate([fl.tri_n, fl.tri_n]),
              tri_material=np.concatenate([fl.tri_material, fl.tri_material]), tri_emission=np.concatenate([fl.tri_emission, np.zeros_like(fl.tri_emission)]))
    sc = gpu_rt.Scene.from_arrays(**kw)
    assert sc.info()["n_tris"] == 184 and sc.info()["scene_in_shared_memory"] == 1           # eligible ...
    auto, st = sc.render_linear(seed=2)
    assert st["scene_in_shared_memory"] == 0 and st["blocks_per_sm"] == 2, st                 # ... but not at the price of occupancy
    forced, st2 = sc.render_linear(seed=2, kernel_variant=2)
    assert st2["scene_in_shared_memory"] == 1 and st2["blocks_per_sm"] == 1, st2
    d = np.abs(auto.astype(np.float64) - forced)
    assert np.median(d) <= 1e-5 * np.abs(auto).mean() and d.mean() <= 2e-2 * np.abs(auto).mean()
    sc.close()


def test_frame_buffer_cache_is_transparent_and_bounded(gpu_rt):
    """rt_scene_destroy parks the big frame buffers for the next scene (include/rt_api.h, rt_release_device_cache): renders must not
    depend on what a recycled buffer held (other frame size, other scene, tile shards, ray_depth 0), and repeated create / render /
    destroy cycles must not grow the device memory in use."""
    import torch
    ref = {}
    for name, W, H, spp in (("practice7_4", 640, 360, 8), ("practice7_1", 512, 512, 4)):
        sc = gpu_rt.Scene.from_gltf(scene_path(name), W, H, spp)
        ref[name] = sc.render(seed=3)[0].copy()
        sc.close()
    gpu_rt.release_device_cache()
    free0 = torch.cuda.mem_get_info()[0]
    for k in range(12):
        name, W, H, spp = (("practice7_4", 640, 360, 8), ("practice7_1", 512, 512, 4))[k % 2]
        sc = gpu_rt.Scene.from_gltf(scene_path(name), W, H, spp)
        if k % 3 == 0:
            sc.render_linear(seed=9, tile_shard=(0, 2))                  # leaves zeros / partial data in the recycled layers
        img, _ = sc.render(seed=3)
        assert np.array_equal(img, ref[name]), (k, name)
        sc.close()
    used_cached = free0 - torch.cuda.mem_get_info()[0]
    assert used_cached < 400e6, used_cached                                # at most the parked blocks of two small frames
    gpu_rt.release_device_cache()
    assert free0 - torch.cuda.mem_get_info()[0] < 64e6


def test_reference_own_test_vndf_on_the_device(gpu_rt):
    """`tests::test_vndf` (tests.rs:43-49, the only live test of the reference for this path) evaluated with the DEVICE pdf: n = z,
    v = normalize(z + z) = z, roughness 0.04, 1e6 uniform directions, `(avg * 4 pi - 1) < 0.05` as written (one-sided) -- and the
    two-sided version of the same integral with an integrator that resolves the alpha = 0.0016 lobe."""
    rng = np.random.default_rng(4349)
    cnt = 1_000_000
    z = np.array([0.0, 0.0, 1.0], dtype=np.float32)
    N = np.tile(z, (cnt, 1)); V = np.tile(z, (cnt, 1)); R = np.full((cnt, 1), 0.04, dtype=np.float32)
    l = rng.normal(size=(cnt, 3)); l /= np.linalg.norm(l, axis=1, keepdims=True)
    pdf = gpu_rt.eval_fn(gpu_rt.FN_PDF_VNDF, np.concatenate([N, l.astype(np.float32), V, R], axis=1))[:, 0].astype(np.float64)
    avg = np.where(np.isfinite(pdf), pdf, 0.0).mean()
    assert (avg * 4 * np.pi - 1.0) < 0.05
    l2, inv_q = _log_polar_directions(np.array([0.0, 0.0, 1.0]), cnt, rng, theta_min=1e-7)
    l32 = np.ascontiguousarray(l2, dtype=np.float32)
    p2 = gpu_rt.eval_fn(gpu_rt.FN_PDF_VNDF, np.concatenate([N, l32, V, R], axis=1))[:, 0].astype(np.float64)
    above = (l32.astype(np.float64) + z) @ z > 0
    w = np.where(above & np.isfinite(p2), p2, 0.0) * inv_q
    assert abs(w.mean() - 1.0) < max(5 * w.std() / np.sqrt(cnt), 0.01), w.mean()
