"""Renders one frame of any scene file through the C ABI (profiling aid): python tools/render_once.py <scene> <W> <H> <spp> [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import rtb200 as rt
path, W, H, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
sc = rt.Scene.from_file(path, W, H, spp)
d = sc.desc()
out = np.zeros((d["height"], d["width"], 3), np.uint8)
for _ in range(reps):
    st = sc.render_into(out, seed=1)
print({k: st[k] for k in ("kernel_ms", "block_threads", "blocks_per_sm", "regs_per_thread", "scene_in_shared_memory")}, d["width"] * d["height"] * d["samples"] / st["kernel_ms"] / 1e3, "Msamples/s")
