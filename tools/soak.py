#!/usr/bin/env python
"""Soak run: every shipped scene (glTF and text) at 1920x1080x256 spp (working.txt: 64) with the instrumented kernel; prints the robustness counters
(attempt-cap hits, dropped non-finite samples) and the per-sample work statistics next to the oracle-independent invariants."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rtb200 as rt
for name in ("practice7_1", "practice7_2", "practice7_3", "practice7_4", "practice3_1", "practice3_2", "practice3_3", "practice3_4", "practice3_5", "working"):
    ext = ".gltf" if name.startswith("practice7") else ".txt"
    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", name + ext), 1920, 1080, 64 if name == "working" else 256)
    img, st = sc.render_linear(seed=11, collect_stats=True)
    n = st["samples"]
    print(json.dumps({"scene": name, "samples": n, "finite_image": bool(np.isfinite(img).all()), "attempt_cap_hits": st["attempt_cap_hits"],
                      "nonfinite_samples": st["nonfinite_samples"], "segments_per_sample": round(st["segments"] / n, 4),
                      "attempts_per_vertex": round(st["attempts"] / st["vertices"], 4), "box_tests_per_segment": round(st["node_tests"] / st["segments"], 2),
                      "tri_tests_per_segment": round(st["tri_tests"] / st["segments"], 2), "kernel_ms": round(st["kernel_ms"], 1), "min": float(img.min()), "max": float(img.max())}), flush=True)
    sc.close()
