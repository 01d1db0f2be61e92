"""compute-sanitizer is closed on the B200 pool (it left GPUs needing a reset), so memory safety is checked with the library's own
instrument: `make -C csrc debug` builds the SAME sources with -DRT_DEBUG_BOUNDS (_build_dbg/librt_b200.so), in which every
shared-memory pool / queue / traversal-stack access, every scene-blob load, every primitive index and every framebuffer store of
the kernels is range-checked and counted (rt_debug_bounds_violations).  The parity workloads must run with all counters zero; a
deliberately lowered blob limit must make the instrument fire (so a zero means something)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

DBG_LIB = os.path.join(ROOT, "raytracing-course-2024_b200", "_build_dbg", "librt_b200.so")

WORKLOAD = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import rtb200 as rt
S = os.path.join(%(root)r, "scenes")
rt.debug_bounds_violations(reset=True)
rng = np.random.default_rng(0)
def rays(sc, n):
    d = sc.desc()
    xy = np.stack([rng.integers(0, d["width"], n), rng.integers(0, d["height"], n)], 1).astype(np.int32)
    return sc.primary_rays(xy, rng.random((n, 2)))
# triangle scenes: shared-memory and global-memory placement, both kernels, stats build, tile shards, odd frame sizes
for name, W, H, spp in (("practice7_4", 97, 61, 24), ("practice7_1", 64, 64, 16), ("practice7_2", 96, 96, 8), ("practice7_3", 64, 48, 8)):
    sc = rt.Scene.from_gltf(os.path.join(S, name + ".gltf"), W, H, spp)
    for kv in (0, 1, 10, 20, 30) + ((2,) if name in ("practice7_4", "practice7_1") else ()):
        sc.render(seed=1, kernel_variant=kv)
    sc.render_linear(seed=2, collect_stats=True)
    sc.render_linear(seed=2, tile_shard=(1, 3))
    sc.render_linear(seed=2, sample_begin=3, sample_end=7)
    r = rays(sc, 4096)
    sc.trace_primary(r, precision=32); sc.trace_primary(r, precision=64); sc.trace_hits(r)
    sc.close()
# general-primitive scenes: boxes, ellipsoids, planes, dielectrics, box / ellipsoid lights, the light BVH (working.txt: 14 lights)
for name, W, H, spp in (("practice3_1", 80, 60, 16), ("practice3_4", 64, 64, 16), ("practice3_5", 64, 64, 16), ("working", 50, 50, 8)):
    sc = rt.Scene.from_text(os.path.join(S, name + ".txt"), W, H, spp)
    for kv in (0, 1) + ((2,) if name != "working" else ()):
        sc.render(seed=1, kernel_variant=kv)
    sc.render_linear(seed=2, collect_stats=True)
    r = rays(sc, 4096)
    sc.trace_primary(r, precision=32); sc.trace_hits(r)
    if name in ("practice3_5", "working"):
        p = (rng.random((2048, 3)) * 6 - 3).astype(np.float32)
        l = rng.normal(size=(2048, 3)); l = (l / np.linalg.norm(l, axis=1, keepdims=True)).astype(np.float32)
        sc.eval(rt.FN_PDF_LIGHT, np.concatenate([p, l], 1))
    sc.close()
# empty scene, ray_depth 0, deep path
e = rt.Scene.from_arrays(width=16, height=16, samples=2, ray_depth=6, bg_color=[0.2, 0.2, 0.2], camera_position=[0, 0, 0], camera_forward=[0, 0, -1], camera_right=[1, 0, 0],
                         camera_up=[0, 1, 0], camera_fov_x=1.0, camera_fov_y=1.0, tri_v=np.zeros((0, 9)), tri_n=np.zeros((0, 9)), tri_material=np.zeros((0, 5)), tri_emission=np.zeros((0, 3)))
e.render(); e.close()
print("COUNTS " + json.dumps(rt.debug_bounds_violations()))
'''


def _run(extra_env):
    env = dict(os.environ, RT_B200_LIB=DBG_LIB, **extra_env)
    r = subprocess.run([sys.executable, "-c", WORKLOAD % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("COUNTS ")][-1]
    return json.loads(line[len("COUNTS "):])


def test_parity_workloads_run_clean_under_the_bounds_checked_build(gpu_rt):
    assert os.path.exists(DBG_LIB), "run __graft_entry__.build() (make -C csrc debug)"
    counts = _run({})
    assert counts == {"pool": 0, "queue": 0, "stack": 0, "blob": 0, "prim": 0, "layer": 0, "chunk": 0}, counts


def test_the_instrument_fires_when_a_limit_is_lowered(gpu_rt):
    counts = _run({"RT_DEBUG_BLOB_LIMIT": "1024"})
    assert counts["blob"] > 0 and counts["pool"] == 0 and counts["stack"] == 0, counts


def test_regular_build_refuses_the_debug_query(gpu_rt):
    with pytest.raises(gpu_rt.RtError) as e:
        gpu_rt.debug_bounds_violations()
    assert e.value.code == gpu_rt.RT_ERR_INVALID
