"""Throughput of the general-primitive path on the text scenes at their native DIMENSIONS (BASELINE.json configs 1-2) and of the
oracle on the same frames (bounded rows) -- numbers for BASELINE.md.  python tools/text_bench.py [spp]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import numpy as np
import oracle as O
import rtb200 as rt

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for name in ["practice3_1", "practice3_2", "practice3_3", "practice3_4", "practice3_5", "working"]:
    path = os.path.join(ROOT, "scenes", name + ".txt")
    sc = rt.Scene.from_text(path, 0, 0, spp)
    d = sc.desc(); W, H, S = d["width"], d["height"], d["samples"]
    out = np.zeros((H, W, 3), np.uint8)
    sc.render_into(out, seed=0)
    best = None
    for _ in range(3):
        st = sc.render_into(out, seed=0)
        best = st if best is None or st["total_ms"] < best["total_ms"] else best
    _, stc = sc.render_linear(seed=0, collect_stats=True)
    fl = O.parse_text_scene(path, W, H, max(1, min(S, 8)))
    osc = O.OracleScene(fl, max_attempts=64)
    step = max(1, H // 32)
    r = osc.render(seed=0, n_threads=0, rows=(step // 2, H, step), want_rgb=False, want_mean=False)
    print(json.dumps({"scene": name, "W": W, "H": H, "spp": S, "gpu_msamples_s": W * H * S / best["total_ms"] / 1e3, "gpu_kernel_ms": best["kernel_ms"], "gpu_total_ms": best["total_ms"],
                      "segments_per_sample": stc["segments"] / stc["samples"], "attempts_per_vertex": stc["attempts"] / max(1, stc["vertices"]),
                      "node_tests_per_segment": stc["node_tests"] / stc["segments"], "prim_tests_per_segment": stc["tri_tests"] / stc["segments"],
                      "regs": best["regs_per_thread"], "block": best["block_threads"], "blocks_per_sm": best["blocks_per_sm"], "smem_scene": best["scene_in_shared_memory"],
                      "oracle_msamples_s": r["stats"]["samples"] / r["stats"]["seconds"] / 1e6, "oracle_threads": os.cpu_count()}), flush=True)
    sc.close()
