"""Kernel time of the headline frame as a function of the sample count: t = a + b * spp.  `a` is the per-launch constant that limits
multi-GPU scaling (every GPU pays it once per frame while its share of the samples shrinks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import rtb200 as rt
W, H = 3840, 2160
sc = rt.Scene.from_gltf(os.path.join(ROOT, "scenes", "practice7_4.gltf"), W, H, 1024)
acc = torch.zeros(H * W * 4, dtype=torch.float32, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
rows = []
for spp in [1, 2, 4, 8, 16, 32, 64, 128, 256]:
    best = 1e30
    for _ in range(3):
        s = sc.render_accumulate_device(acc.data_ptr(), st.cuda_stream, want_stats=True, seed=0, sample_begin=0, sample_end=spp)
        best = min(best, s["kernel_ms"])
    rows.append((spp, best, s["n_chunks"]))
    print(spp, "spp:", round(best, 3), "ms", s["n_chunks"], "chunks", round(W * H * spp / best / 1e3, 1), "Msamples/s", flush=True)
x = np.array([r[0] for r in rows[4:]], float); y = np.array([r[1] for r in rows[4:]])
b, a = np.polyfit(x, y, 1)
print("fit over spp >= 16: a = %.3f ms per launch, b = %.4f ms per spp (%.1f Msamples/s asymptotic)" % (a, b, W * H / b / 1e3))
