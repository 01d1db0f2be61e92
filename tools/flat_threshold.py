"""Where does the flat scan stop paying?  practice3_5.txt plus k small boxes / ellipsoids scattered in the room, rendered with the
tree walk (RT_FLAT_SCAN_MAX=0) and with the flat scan (RT_FLAT_SCAN_MAX=64): python tools/flat_threshold.py [W H spp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import rtb200 as rt
W, H, spp = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1024, 1024, 64)
base = open(os.path.join(ROOT, "scenes", "practice3_5.txt")).read()
rng = np.random.default_rng(3)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for extra in (0, 2, 4, 8, 12, 16, 24):
    text = base
    for k in range(extra):
        p = rng.uniform(-4, 4, 3)
        if k % 2 == 0:
            text += f"\nNEW_PRIMITIVE\nBOX 0.3 0.3 0.3\nPOSITION {p[0]:.3f} {p[1]:.3f} {p[2]:.3f}\nCOLOR 0.8 0.8 0.8\n"
        else:
            text += f"\nNEW_PRIMITIVE\nELLIPSOID 0.3 0.4 0.3\nPOSITION {p[0]:.3f} {p[1]:.3f} {p[2]:.3f}\nCOLOR 0.8 0.8 0.8\n"
    path = os.path.join(ROOT, "gpurun_out", f"flat_threshold_{extra}.txt")
    open(path, "w").write(text)
    row = {"primitives": 8 + extra}
    for label, cap in (("tree", "0"), ("flat", "64")):
        os.environ["RT_FLAT_SCAN_MAX"] = cap
        sc = rt.Scene.from_file(path, W, H, spp)
        out = np.zeros((H, W, 3), np.uint8)
        best = min(sc.render_into(out, seed=1)["kernel_ms"] for _ in range(3))
        row[label + "_msamples_s"] = round(W * H * spp / best / 1e3, 1)
        sc.close()
    print(row, flush=True)
