#!/usr/bin/env python
"""Fills the rows of BASELINE.md section 3 on a GPU box: for each (scene, W, H, spp) the CPU oracle rate on all host cores
(bounded row sample), the 1-GPU rate (library CUDA events around the render kernel, best of 3), the FP32 algorithmic-flop
roofline fraction (same constants as bench.py), primary-hit parity against the oracle and, where a committed converged
oracle render exists (tests/golden), the mean-luminance difference of a GPU render at that size.  One JSON line per row."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O          # noqa: E402  (test infrastructure: CPU column and parity checks only)
import rtb200 as rt         # noqa: E402
C_NODE, C_TRI, C_ATTEMPT, C_SHADE = 20.0, 50.0, 260.0, 120.0
ROWS = [("practice3_1", 640, 480, 64), ("practice3_5", 512, 512, 64), ("practice7_1", 512, 512, 256), ("practice7_4", 512, 512, 1024), ("practice7_4", 3840, 2160, 1024),
        ("practice7_2", 512, 512, 64), ("practice7_2", 1920, 1080, 64), ("practice7_3", 512, 512, 64), ("practice7_3", 1920, 1080, 64), ("practice3_3", 512, 512, 64),
        ("practice3_4", 512, 512, 64), ("working", 400, 400, 256)]
GOLD = {"practice7_1": (64, 64, 16384), "practice7_4": (64, 64, 16384), "practice7_2": (32, 32, 4096), "practice7_3": (32, 32, 4096), "practice3_1": (80, 60, 4096),
        "practice3_5": (64, 64, 4096), "practice3_3": (64, 64, 4096), "practice3_4": (64, 64, 4096), "working": (50, 50, 16384)}


def scene_file(name):
    g = os.path.join(ROOT, "scenes", name + ".gltf")
    return g if os.path.exists(g) else os.path.join(ROOT, "scenes", name + ".txt")


def flat(path, W, H, s):
    return O.convert_gltf_to_scene(path, W, H, s) if path.endswith(".gltf") else O.parse_text_scene(path, W, H, s)


peak, _ = rt.measure_fp32_peak(0)
lum = lambda a: a @ np.array([0.2126, 0.7152, 0.0722])   # noqa: E731
for name, W, H, spp in ROWS:
    path = scene_file(name)
    cap = 64 if name == "working" else 0
    sc = rt.Scene.from_file(path, W, H, spp)
    best = None
    for _ in range(3):
        _, st = sc.render(seed=1)
        best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
    gpu = W * H * spp / best / 1e3
    sc.set_frame(max(64, W // 8), max(36, H // 8), 32)
    _, cs = sc.render_linear(seed=1, collect_stats=True)
    n = cs["samples"]
    flop = (cs["node_tests"] * C_NODE + cs["tri_tests"] * C_TRI + cs["vertices"] * C_SHADE + cs["attempts"] * C_ATTEMPT) / n
    # CPU oracle: every k-th row of the same frame at a reduced spp, ~8 s
    rows = 16; step = max(1, H // rows)
    def run(s):
        osc = O.OracleScene(flat(path, W, H, s), max_attempts=cap)
        r = osc.render(seed=0, n_threads=os.cpu_count(), rows=(step // 2, H, step), want_rgb=False, want_mean=False)
        osc.close(); return r["stats"]
    p = run(1); rate = p["samples"] / max(p["seconds"], 1e-6)
    s2 = int(max(1, min(spp, rate * 8.0 / (len(range(step // 2, H, step)) * W))))
    cst = run(s2)
    cpu = cst["samples"] / cst["seconds"] / 1e6
    # primary-hit parity at 256x256 pixel centres
    osc = O.OracleScene(flat(path, 256, 256, 1))
    xs, ys = np.meshgrid(np.arange(256), np.arange(256)); xy = np.stack([xs.ravel(), ys.ravel()], axis=1).astype(np.int32)
    rays = osc.primary_rays(xy, np.full((len(xy), 2), 0.5)); ref = osc.trace_primary(rays, want_second=False)
    sc.set_frame(256, 256, 1)
    tid, t = sc.trace_primary(rays, precision=32)
    same = float((tid == ref["tri_id"]).mean()); hit = (tid == ref["tri_id"]) & (tid >= 0)
    trel = float(np.max(np.abs(t[hit] - ref["t"][hit]) / ref["t"][hit]))
    bad = np.nonzero(tid != ref["tri_id"])[0]
    n_tie = 0
    if bad.size:
        sub = osc.trace_primary(rays[bad], want_second=True)
        margin = np.minimum(np.minimum(sub["u"], sub["v"]), 1 - sub["u"] - sub["v"])
        with np.errstate(invalid="ignore"):
            n_tie = int((((margin <= 1e-5) | (np.abs(sub["second_t"] - sub["t"]) <= 1e-5 * np.abs(sub["t"]))) & (sub["tri_id"] >= 0) & (tid[bad] >= 0)).sum())
    osc.close()
    # image parity against the committed converged oracle render
    gw, gh, gs = GOLD[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", f"converged_{name}_{gw}x{gh}_{gs}.npz"))
    sc.set_frame(gw, gh, gs)
    img, _ = sc.render_linear(seed=5)
    dl = float(abs(lum(img.astype(np.float64)).mean() - lum(g["mean"]).mean()) / lum(g["mean"]).mean())
    rmse = float(np.sqrt(np.mean((img - g["mean"]) ** 2)))
    print(json.dumps({"scene": name, "W": W, "H": H, "spp": spp, "cpu_msamples_s": round(cpu, 2), "cpu_cores": os.cpu_count(), "gpu_msamples_s": round(gpu, 1),
                      "kernel_ms": round(best, 2), "flop_per_sample": round(flop, 1), "fp32_roofline_frac": round(flop * gpu * 1e6 / 1e12 / peak, 4),
                      "primary_hit_same": round(same, 6), "primary_mismatches": int(bad.size), "primary_mismatches_that_are_ties": n_tie, "primary_t_max_rel": trel,
                      "blocks_per_sm": st["blocks_per_sm"], "block_threads": st["block_threads"], "scene_in_shared_memory": st["scene_in_shared_memory"], "image_mean_lum_rel_diff": round(dl, 5), "image_rmse": round(rmse, 5),
                      "image_at": f"{gw}x{gh}x{gs}"}), flush=True)
    sc.close()
