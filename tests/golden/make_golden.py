"""Generates the committed golden fixtures of tests/golden/ by running the ORACLE (oracle/, the f64 CPU restatement
of the reference) in this container.  The reference's own Rust binary cannot be built here (no cargo/rustc), and
the reference ships no fixtures of its own, so these are oracle outputs: "parity unpinned".

    python tests/golden/make_golden.py            # regenerates every fixture (several minutes on 8 cores)

Fixtures (float32 to keep them small):
  converged_<scene>_<W>x<H>_<spp>.npz : mean (H,W,3) linear radiance, var (H,W,3) per-sample variance, spp, seed,
                                        stats (segments/sample etc.) -- the "converged reference" of SURVEY.md 8d.
  kat.json                            : known-answer values of SURVEY.md A.3 re-evaluated by the oracle.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CONVERGED = [  # scene, W, H, spp
    ("practice7_4", 64, 64, 16384),
    ("practice7_1", 64, 64, 16384),
    ("practice7_2", 32, 32, 4096),
    ("practice7_3", 32, 32, 4096),
    ("practice7_4", 64, 36, 8192),   # 16:9 like the 4K north-star frame (fov_x = fov_y stretch)
]


def converged():
    for scene, W, H, spp in CONVERGED:
        out = os.path.join(HERE, f"converged_{scene}_{W}x{H}_{spp}.npz")
        if os.path.exists(out) and "--force" not in sys.argv:
            print("keep", out)
            continue
        fl = O.convert_gltf_to_scene(os.path.join(ROOT, "scenes", scene + ".gltf"), W, H, spp)
        sc = O.OracleScene(fl)
        t0 = time.time()
        r = sc.render(seed=0, n_threads=0, want_rgb=True, want_var=True)
        st = r["stats"]
        print(scene, W, H, spp, "%.1fs" % (time.time() - t0), st)
        np.savez_compressed(out, mean=r["mean"].astype(np.float32), var=r["var"].astype(np.float32), rgb=r["rgb"], spp=spp, seed=0,
                            stats=json.dumps(st))


def kat():
    nz = lambda v: np.array(v, float) / np.linalg.norm(v)  # noqa: E731
    v, l, n, l2 = nz([-1, 0, 1]), nz([1, 0, 1]), np.array([0.0, 0.0, 1.0]), nz([0.3, 0.4, 0.8])
    fl = O.convert_gltf_to_scene(os.path.join(ROOT, "scenes", "practice7_4.gltf"), 64, 64, 1)
    sc = O.OracleScene(fl)
    p = np.array([[0.0, -1.99999, 0.0]])
    cen = fl.tri_v[fl.light_ids[0]].reshape(3, 3).mean(0)
    d = (cen - p[0]) / np.linalg.norm(cen - p[0])
    k = {
        "specular_brdf_r005": float(O.specular_brdf([l], [n], [v], [n], [0.05])[0]),
        "brdf_diel_r05": O.brdf([l], [n], [v], [[1, 0.25, 0.125, 0, 0.5]])[0].tolist(),
        "brdf_metal_r003": O.brdf([l], [n], [v], [[1, 1, 1, 1, 0.03]])[0].tolist(),
        "brdf_diel_r1": O.brdf([l2], [n], [v], [[0.8, 0.2, 0.2, 0, 1]])[0].tolist(),
        "vndf_pdf_mirror_dir": float(O.pdf_vndf([n], [l], [v], [0.5])[0]),
        "vndf_pdf_l2": float(O.pdf_vndf([n], [l2], [v], [0.5])[0]),
        "cosine_pdf_l2": float(O.pdf_cosine([n], [l2])[0]),
        "light_pdf_7_4": float(sc.pdf_light(p, [d])[0]),
        "color_to_pixel": O.color_to_pixel([[0, 0, 0], [0.18, 0.18, 0.18], [0.5, 1, 2], [10, 10, 10], [0.01, 0.05, 0.25]]).tolist(),
        "primary_ray_7_4_64_00": sc.primary_rays([[0, 0]], [[0.5, 0.5]])[0].tolist(),
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(k, f, indent=1)
    print("kat.json written")


if __name__ == "__main__":
    kat()
    converged()
