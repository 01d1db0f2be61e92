"""Do the GPU path and the oracle agree on the NEGATIVE tail of the estimator (the reference weighs with the geometric normal but
accepts on the shading normal, which stays in object space for rotated boxes -- negative samples are its behaviour)?  Per-pixel
means of practice3_5.txt at 192x192x64: fraction of negative pixels and low quantiles, GPU vs oracle (different RNG streams)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import numpy as np
import oracle
import rtb200 as rt
for name, W, H, spp in (("practice3_5", 192, 192, 64), ("practice3_2", 192, 144, 64)):
    path = os.path.join(ROOT, "scenes", name + ".txt")
    fl = oracle.parse_text_scene(path, W, H, spp)
    om = oracle.OracleScene(fl).render(seed=3, want_rgb=False)["mean"]
    sc = rt.Scene.from_text(path, W, H, spp)
    gm, _ = sc.render_linear(seed=3)
    gm = gm.astype(np.float64)
    for label, m in (("oracle", om), ("gpu", gm)):
        lum = m @ np.array([0.2126, 0.7152, 0.0722])
        print(name, label, "mean %.5f" % lum.mean(), "negative pixels %.5f" % (lum < 0).mean(), "min %.3f" % lum.min(),
              "q0.1%% %.4f q1%% %.4f q50%% %.4f q99.9%% %.3f" % tuple(np.quantile(lum, [0.001, 0.01, 0.5, 0.999])))
    sc.close()
