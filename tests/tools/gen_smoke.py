"""Quick GPU check of the general-primitive path against the oracle (development aid; the real gates are tests/test_gpu_general.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import oracle as O
import rtb200 as rt

lum = lambda a: a @ np.array([0.2126, 0.7152, 0.0722])
for name in sys.argv[1:] or ["practice3_1", "practice3_5", "practice3_2", "practice3_3", "practice3_4"]:
    path = os.path.join(ROOT, "scenes", name + ".txt")
    fl = O.parse_text_scene(path)
    W, H = fl.width // 4, fl.height // 4
    fl.width, fl.height, fl.samples = W, H, 256
    osc = O.OracleScene(fl, max_attempts=64)
    sc = rt.Scene.from_text(path, W, H, 256)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    rays = osc.primary_rays(xy, np.full((len(xy), 2), 0.5))
    ref = osc.trace_primary(rays, want_second=False)
    tid, t = sc.trace_primary(rays, precision=32)
    same = tid == ref["tri_id"]
    hit = same & (tid >= 0)
    rel = np.abs(t[hit] - ref["t"][hit]) / ref["t"][hit]
    print(name, "ids equal", same.mean(), "mismatch", int((~same).sum()), "max rel t", rel.max() if rel.size else None)
    hg, ho = sc.trace_hits(rays), osc.trace_hits(rays)
    print("   ng err", np.abs(hg[hit, 1:4] - ho[hit, 1:4]).max(), "ns err", np.abs(hg[hit, 4:7] - ho[hit, 4:7]).max(), "outer eq", (hg[hit, 8] == ho[hit, 8]).mean())
    t0 = time.time(); o = osc.render(seed=0, n_threads=0, want_var=True); t1 = time.time()
    img, st = sc.render_linear(seed=3, collect_stats=True)
    img = img.astype(np.float64)
    lg, lo = lum(img).mean(), lum(o["mean"]).mean()
    rmse = np.sqrt(np.mean((img - o["mean"]) ** 2)); floor = np.sqrt(np.mean(o["var"] * 2.0 / 256))
    print("   lum gpu %.5f oracle %.5f rel %.4f  rmse %.5f floor %.5f  seg/sample gpu %.3f oracle %.3f  att/vert gpu %.3f  cap %d nonfinite %d  kernel %.1f ms oracle %.1f s"
          % (lg, lo, (lg - lo) / lo, rmse, floor, st["segments"] / st["samples"], o["stats"]["segments"] / o["stats"]["samples"],
             st["attempts"] / max(1, st["vertices"]), st["attempt_cap_hits"], st["nonfinite_samples"], st["kernel_ms"], t1 - t0))
    from PIL import Image
    u8, _ = sc.render(seed=3)
    Image.fromarray(u8).save(os.path.join(ROOT, "gpurun_out", f"gen_{name}.png"))
    Image.fromarray(o["rgb"]).save(os.path.join(ROOT, "gpurun_out", f"gen_{name}_oracle.png"))
