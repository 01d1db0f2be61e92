"""GPU LBVH + SAH rebuild of the top (RT_BVH_TOP_SAH clusters): build time, box tests per segment, Msamples/s.  python tools/lbvh_top_sweep.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import rtb200 as rt
for name in ("practice7_2", "practice7_3"):
    for builder, tops in (("host", (0,)), ("gpu", tuple(int(v) for v in os.environ.get("TOPS", "0,64,512,4096,8192,16384").split(",")))):
        for top in tops:
            os.environ.pop("RT_BVH_BUILDER", None)
            if builder == "gpu": os.environ["RT_BVH_BUILDER"] = "gpu"
            os.environ["RT_BVH_TOP_SAH"] = str(top)
            sc = rt.Scene.from_gltf(os.path.join(ROOT, "scenes", name + ".gltf"), 1920, 1080, 64)
            info = sc.info()
            sc.set_frame(480, 270, 16)
            _, st = sc.render_linear(seed=1, collect_stats=True)
            sc.set_frame(1920, 1080, 64)
            best = min(sc.render(seed=1)[1]["kernel_ms"] for _ in range(3))
            print(json.dumps({"scene": name, "builder": builder, "top_clusters": top, "build_ms": round(info["bvh_build_ms"], 2), "nodes": info["n_nodes"], "depth": info["bvh_depth"],
                              "validate_failures": info["bvh_validate_failures"], "box_tests_per_seg": round(st["node_tests"] / st["segments"], 2),
                              "tri_tests_per_seg": round(st["tri_tests"] / st["segments"], 2), "msamples_s": round(1920 * 1080 * 64 / best / 1e3, 1)}), flush=True)
            sc.close()
