#!/usr/bin/env python
"""Performance sweep on one GPU: renders a fixed frame under different BVH build knobs / kernel variants and prints
one JSON line per configuration (kernel time from the library's CUDA events, work counters from the stats build).

    python tools/sweep.py [--scene practice7_4] [--width 1920 --height 1080 --spp 256]"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200 as rt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="practice7_4")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spp", type=int, default=256)
ap.add_argument("--leaf", default="1,2,4")
ap.add_argument("--cost", default="0.5,1.0,2.0")
ap.add_argument("--variants", default="10,20,30")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
path = os.path.join(ROOT, "scenes", a.scene + ".gltf")
for leaf, cost in itertools.product(a.leaf.split(","), a.cost.split(",")):
    os.environ["RT_BVH_MAX_LEAF"], os.environ["RT_BVH_TRAV_COST"] = leaf, cost
    sc = rt.Scene.from_gltf(path, a.width, a.height, a.spp)
    info = sc.info()
    sc.set_frame(max(64, a.width // 8), max(36, a.height // 8), 32)
    _, st = sc.render_linear(seed=1, collect_stats=True)
    sc.set_frame(a.width, a.height, a.spp)
    n = st["samples"]
    for variant in [int(v) for v in a.variants.split(",")]:
        best = None
        for _ in range(a.reps):
            _, s = sc.render(seed=1, kernel_variant=variant)
            best = s["kernel_ms"] if best is None else min(best, s["kernel_ms"])
        print(json.dumps({"scene": a.scene, "max_leaf": leaf, "trav_cost": cost, "variant": variant, "nodes": info["n_nodes"], "depth": info["bvh_depth"],
                          "kernel_ms": round(best, 3), "msamples_s": round(a.width * a.height * a.spp / best / 1e3, 1),
                          "box_tests_per_seg": round(st["node_tests"] / st["segments"], 2), "tri_tests_per_seg": round(st["tri_tests"] / st["segments"], 2),
                          "seg_per_sample": round(st["segments"] / n, 3), "att_per_vertex": round(st["attempts"] / st["vertices"], 3)}), flush=True)
    sc.close()
