import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import rtb200 as rt
W, H, S = 3840, 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 128
scene = rt.Scene.from_gltf(os.path.join(ROOT, "scenes", "practice7_4.gltf"), W, H, S)
desc = scene.desc()
rgb_host = torch.zeros(H * W * 3, dtype=torch.uint8).pin_memory()
pageable = np.zeros(H * W * 3, dtype=np.uint8)
for rep in range(4):
    t0 = time.perf_counter()
    sc2 = rt.Scene.from_arrays(width=W, height=H, samples=S, ray_depth=desc["ray_depth"], bg_color=desc["bg_color"], camera_position=desc["camera_position"],
                               camera_forward=desc["camera_forward"], camera_right=desc["camera_right"], camera_up=desc["camera_up"], camera_fov_x=desc["camera_fov_x"],
                               camera_fov_y=desc["camera_fov_y"], tri_v=desc["tri_v"], tri_n=desc["tri_n"], tri_material=desc["tri_material"], tri_emission=desc["tri_emission"])
    t1 = time.perf_counter()
    st = sc2.render_into(rgb_host.numpy(), seed=0)
    t2 = time.perf_counter()
    st2 = sc2.render_into(rgb_host.numpy(), seed=0)
    t3 = time.perf_counter()
    st3 = sc2.render_into(pageable, seed=0)
    t4 = time.perf_counter()
    sc2.close()
    t5 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms, first render wall {1e3*(t2-t1):.1f} (device total {st['total_ms']:.1f}, kernel {st['kernel_ms']:.1f}, render {st['render_ms']:.1f}, resolve {st['resolve_ms']:.2f}), "
          f"second render wall {1e3*(t3-t2):.1f} (device {st2['total_ms']:.1f}), pageable wall {1e3*(t4-t3):.1f} (device {st3['total_ms']:.1f}), close {1e3*(t5-t4):.1f} ms", flush=True)
