#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path tracer: Msamples/s (camera paths) on practice7_4.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full frame of the north-star workload (BASELINE.json configs[4]):
scenes/practice7_4.gltf at 3840x2160, 1024 spp, ray_depth 6 = 8.49e9 camera paths.  With N GPUs the frame is
sample-sharded (rank r renders samples [r*S/N, (r+1)*S/N) of every pixel), reduced once with NCCL and resolved on
rank 0: total work is fixed -> "scaling": "strong".

Keys of the JSON line (rank 0):
  value        whole-job Msamples/s, scene resident on the device, device-timed (CUDA events, max over ranks)
  e2e          same metric through the C ABI with HOST buffers: rt_scene_create from host arrays (BVH build +
               H2D scene upload) + render + (reduce) + resolve + D2H of the W*H*3 result, every step
  roofline     FP32-issue roofline of the render kernel (the scene is ~21 KB and shared-memory
               resident, so HBM is not the bound): achieved = algorithmic flop/sample (fixed constants x counts
               measured by the kernel's own stats mode on the TIMED frame shape, 16 spp) x samples / kernel time;
               peak = FFMA micro-benchmark run in this process ("measured here"); hbm_* = the secondary HBM figures
               from the launch's actual chunk count; traffic = ncu DRAM bytes of this launch when profiles/ holds a
               capture of THIS build (hash of the kernel sources), else null
  phases       per-frame device time of render (path tracing + layer sum) / reduce / resolve, max and min over ranks
  cpu_baseline the oracle (f64 CPU restatement of the reference; the Rust binary cannot be built here) on a
               bounded sample of the same workload on this box's host cores
`--impl reference` times that CPU restatement as its own arm (rank 0 only)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE = "practice7_4"
TRAFFIC_FILE = "r2_bench_traffic.json"
KERNEL_SOURCES = ("rt_render_wave.cuh", "rt_device.cuh", "rt_kernels.cu", "rt_kernels.h")


def kernel_source_sha16():
    """Identity of the render kernel's build: sha256 over the kernel sources.  An ncu capture under profiles/ is quoted as
    `roofline.traffic` only when it was taken from exactly these sources (tools/ncu_traffic.py stores the same hash)."""
    import hashlib
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, "raytracing-course-2024_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


C_NODE, C_TRI, C_ATTEMPT, C_SHADE = 20.0, 50.0, 260.0, 120.0        # flop constants of SURVEY.md 8d


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--scene", default=SCENE)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of one oracle sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=0, help="kernel_variant passed to the library (0 auto)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if not (t0 <= ts <= t1 + 0.3):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except Exception:
                continue
            for k, nm in enumerate(names):
                if len(f) > 3 + k and f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "n_samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_sample(args, seconds, n_threads=0):
    """Times the oracle on a bounded sample of the workload: every `step`-th row of the full-size frame at a reduced
    spp, sized from a 1 s probe to take ~`seconds`.  Returns (Msamples/s, cores, description)."""
    import oracle as O
    cores = n_threads or os.cpu_count() or 1
    path = os.path.join(ROOT, "scenes", args.scene + ".gltf")
    W, H = args.width, args.height
    rows = 16
    step = max(1, H // rows)

    def run(spp):
        sc = O.OracleScene(O.convert_gltf_to_scene(path, W, H, spp))
        r = sc.render(seed=0, n_threads=cores, rows=(step // 2, H, step), want_rgb=False, want_mean=False)
        sc.close()
        return r["stats"]
    probe = run(1)
    rate = probe["samples"] / max(probe["seconds"], 1e-6)
    n_rows = len(range(step // 2, H, step))
    spp = int(max(1, min(args.spp, rate * seconds / (n_rows * W))))
    st = run(spp)
    ms = st["samples"] / st["seconds"] / 1e6
    desc = (f"{args.scene} {W}x{H}: rows {step // 2}::{step} ({n_rows} rows x {W} px) at {spp} spp = {st['samples']} paths in {st['seconds']:.1f} s; "
            f"oracle = f64 C++ restatement of the reference (not the Rust binary), std::thread over rows")
    return ms, cores, desc, st


def run_reference(args, rank):
    if rank != 0:
        return
    per_step = max(2.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        oracle_sample(args, per_step)
    t0 = time.time()
    vals, desc, cores = [], "", 0
    tot_samples, tot_sec = 0, 0.0
    for _ in range(args.steps):
        ms, cores, desc, st = oracle_sample(args, per_step)
        vals.append(ms); tot_samples += st["samples"]; tot_sec += st["seconds"]
    value = tot_samples / tot_sec / 1e6
    line = {
        "impl": "reference", "metric": "Msamples/s (camera paths) on practice7_4", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_sec / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "scene file scenes/practice7_4.gltf (fixed input)",
        "config": {"workload": f"{args.scene} {args.width}x{args.height} {args.spp} spp ray_depth 6 (bounded row/spp sample per step)"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


def smem_roofline(counts, samples, kern_ms, clk):
    per_sample = counts["node_tests"] / 2 * 48.0 + counts["tri_tests"] * 48.0 + counts["attempts"] * 144.0 + counts["segments"] * 216.0
    mhz = (clk or {}).get("sm_mhz") or 1965.0
    peak = 148 * 128.0 * mhz * 1e6 / 1e12
    ach = per_sample * samples / (kern_ms * 1e-3) / 1e12
    return {"algorithmic_bytes_per_sample": per_sample, "achieved_tbs": ach, "peak_tbs": peak, "frac": ach / peak,
            "note": "ncu: the pipe is ~84 % busy at this rate (bank conflicts and partially filled wavefronts cost 2.6x the algorithmic bytes)"}


def kernel_name(st):
    space = "SmemSpace" if st["scene_in_shared_memory"] else "GmemSpace"
    if st["kernel"] == 1:
        return f"render_kernel<{space},false> (v1 per-lane megakernel)"
    mode = {2: "0 (while-while bursts)", 3: "1 (phased bursts)"}.get(st["kernel"], "?")
    return f"render_wave_kernel<{space},false,MODE={mode},BLOCK={st['block_threads']},MINB={st['blocks_per_sm']}> (v3 warp-local wavefront)"


# ------------------------------------------------------------------------------------------------ B200 arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import rtb200 as rt
    from raytracing_course_2024_b200 import multigpu

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    W, H, S = args.width, args.height, args.spp
    path = os.path.join(ROOT, "scenes", args.scene + ".gltf")
    scene = rt.Scene.from_gltf(path, W, H, S, device=local_rank)
    desc = scene.desc()                                    # host arrays for the e2e leg
    info = scene.info()
    stream = torch.cuda.Stream(device=dev)                 # a real (non-default) stream: the library launches on it
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0
    accum = torch.zeros(H * W * 4, dtype=torch.float32, device=dev)
    rgb_dev = torch.zeros(H * W * 3, dtype=torch.uint8, device=dev)
    rgb_host = torch.zeros(H * W * 3, dtype=torch.uint8).pin_memory()
    lo, hi = multigpu.shard_range(S, rank, world)

    def resolve(acc_ptr, rgb_ptr, s):
        rt.resolve_device(acc_ptr, W, H, rgb_ptr, s)

    ev_pairs = []
    n_chunks_seen = []

    def step(record=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record else None
        multigpu.render_frame_sharded(scene, accum, rgb_dev, seed=args.seed, samples=S, rank=rank, world=world, stream_ptr=sp, resolve=resolve,
                                      mark=(lambda k: ev[k].record(stream)) if record else None, kernel_variant=args.variant)
        if record:
            ev_pairs.append(ev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- roofline inputs measured before the timed region: FFMA peak + work counters (stats kernel, reduced frame)
    peak_tflops, sm_max_mhz = rt.measure_fp32_peak(local_rank)
    flop_per_sample = None
    counts = None
    if rank == 0:
        # counters of the TIMED frame shape (same pixels, same camera), 16 spp through the instrumented build: ~50 ms of kernel time
        scene.set_frame(W, H, min(16, S))
        stats_acc = torch.zeros(H * W * 4, dtype=torch.float32, device=dev)
        st = scene.render_accumulate_device(stats_acc.data_ptr(), sp, want_stats=True, seed=args.seed, collect_stats=True, kernel_variant=args.variant)
        del stats_acc
        scene.set_frame(W, H, S)
        kcfg = scene.render_accumulate_device(accum.data_ptr(), sp, want_stats=True, seed=args.seed, sample_begin=lo, sample_end=min(hi, lo + 8),
                                              kernel_variant=args.variant)                  # the timed build's launch configuration (8 samples)

        n = st["samples"]
        counts = {k: st[k] / n for k in ("segments", "vertices", "attempts", "node_tests", "tri_tests", "light_tri_tests")}
        counts["attempt_cap_hits"] = st["attempt_cap_hits"]; counts["nonfinite_samples"] = st["nonfinite_samples"]
        flop_per_sample = (counts["node_tests"] * C_NODE + counts["tri_tests"] * C_TRI + counts["vertices"] * C_SHADE + counts["attempts"] * C_ATTEMPT)

    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.time()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(stream)
    for _ in range(args.steps):
        step(record=True)
    e_end.record(stream)
    barrier()
    t_wall1 = time.time()
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    ms_total = max_over_ranks(e_start.elapsed_time(e_end))
    nst = max(1, len(ev_pairs))
    my_phase = [sum(e[k].elapsed_time(e[k + 1]) for e in ev_pairs) / nst for k in range(4)]      # zero, render, reduce, resolve
    kern_ms_local = my_phase[1]
    kern_ms = max_over_ranks(kern_ms_local)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, my_phase)
    else:
        gathered = [my_phase]
    phases = {name: {"max": max(g[k] for g in gathered), "min": min(g[k] for g in gathered)} for k, name in enumerate(("zero_ms", "render_ms", "reduce_ms", "resolve_ms"))}
    phases["resolve_ms"] = {"rank0": gathered[0][3]}
    # a rank that finishes its render early waits inside the reduce: the collective's own cost is its MIN over ranks, and
    # (max render - min render) is the rank skew of the persistent kernel's tail
    phases["note"] = "reduce min over ranks = the collective itself; render max - min = rank skew (absorbed by the other ranks' reduce wait)"
    total_samples = float(W) * H * S
    value = total_samples * args.steps / (ms_total * 1e-3) / 1e6

    # ---- e2e: host buffers in, host bytes out, through the C ABI, every step
    def e2e_step():
        sc2 = rt.Scene.from_arrays(width=W, height=H, samples=S, ray_depth=desc["ray_depth"], bg_color=desc["bg_color"],
                                   camera_position=desc["camera_position"], camera_forward=desc["camera_forward"], camera_right=desc["camera_right"],
                                   camera_up=desc["camera_up"], camera_fov_x=desc["camera_fov_x"], camera_fov_y=desc["camera_fov_y"], tri_v=desc["tri_v"],
                                   tri_n=desc["tri_n"], tri_material=desc["tri_material"], tri_emission=desc["tri_emission"], device=local_rank)
        if world == 1:
            sc2.render_into(rgb_host.numpy(), seed=args.seed, kernel_variant=args.variant)       # rt_render: D2H inside the call
        else:
            multigpu.render_frame_sharded(sc2, accum, rgb_dev, seed=args.seed, samples=S, rank=rank, world=world, stream_ptr=sp, resolve=resolve,
                                          kernel_variant=args.variant)
            if rank == 0:
                rgb_host.copy_(rgb_dev, non_blocking=True)
            torch.cuda.synchronize()
        sc2.close()
    e2e_step()
    barrier()
    e2e_n = max(1, min(args.steps, 2))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_n):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = total_samples * e2e_n / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        assert int(rgb_host.max()) > 0, "rendered frame is black"
        frac_samples = (hi - lo) / S
        achieved = flop_per_sample * total_samples * frac_samples / (kern_ms * 1e-3) / 1e12      # this rank's kernel
        full = scene.render_accumulate_device(accum.data_ptr(), sp, want_stats=True, seed=args.seed, sample_begin=lo, sample_end=hi, kernel_variant=args.variant)
        n_chunks = int(full["n_chunks"])                                                    # of the timed launch (same arguments)
        # the render kernel writes n_chunks x W*H float4 layers (one store per (pixel, chunk)) -- its algorithmic HBM bytes; the whole
        # step also reads them back in sum_layers and reads + writes the accumulator once
        hbm_bytes = float(W) * H * 16 * n_chunks
        hbm_step_bytes = float(W) * H * 16 * (2 * n_chunks + 2)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic, traffic_src = None, "unmeasured for this build (no ncu capture of these kernel sources under profiles/)"
        try:                                                  # measured DRAM bytes of this exact launch (ncu, committed under profiles/)
            tr = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
            if (W, H, S, world, args.scene) == (3840, 2160, 1024, 1, SCENE) and tr.get("kernel_source_sha16") == kernel_source_sha16():
                traffic = tr["dram__bytes_read.sum"] + tr["dram__bytes_write.sum"]
                traffic_src = (f"profiles/{TRAFFIC_FILE}: ncu dram__bytes_read.sum + dram__bytes_write.sum of the render kernel of `{tr.get('command')}`, "
                               f"kernel sources sha16 {tr.get('kernel_source_sha16')} = this build")
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ms, cores, cdesc, _ = oracle_sample(args, args.cpu_seconds)
            cpu = {"value": ms, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": cdesc}
        line = {
            "metric": "Msamples/s (camera paths) on practice7_4", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "scene file scenes/practice7_4.gltf (fixed input, 92 triangles); no synthetic tensors",
            "config": {"workload": f"{args.scene} {W}x{H} {S} spp ray_depth 6, sample-sharded over {world} GPU(s), Philox seed {args.seed}",
                       "l2_note": "inputs are a 21 KB scene staged in shared memory; each step rewrites %.0f MB of radiance layers and the 133 MB accumulator (> L2 126 MB; nothing is re-read across steps)" % (hbm_bytes / 1e6),
                       "scene_in_shared_memory": bool(info["scene_in_shared_memory"]), "bvh_nodes": info["n_nodes"], "kernel_variant": args.variant},
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(info["device_bytes"]) * world, "d2h_bytes_per_step": W * H * 3,
                    "steps": e2e_n, "ms_per_step": e2e_ms / e2e_n},
            "gpu_launches": args.steps * (2 * world + 1),
            "clocks": clk,
            "phases": phases,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops if peak_tflops else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": "FFMA micro-benchmark measured in this process (MEASURED_PEAKS.json has no FP32 entry)",
                         "kernel": kernel_name(kcfg), "launch": {k: kcfg[k] for k in ("block_threads", "blocks_per_sm", "grid_blocks", "regs_per_thread", "smem_bytes_per_block")},
                         "kernel_ms": kern_ms, "flop_per_sample": flop_per_sample, "per_sample": counts,
                         "per_sample_source": f"stats build on the timed frame shape {W}x{H} at {min(16, S)} spp", "n_chunks": n_chunks,
                         "kernel_source_sha16": kernel_source_sha16(),
                         "flop_constants": {"node": C_NODE, "tri": C_TRI, "attempt": C_ATTEMPT, "shade": C_SHADE},
                         "hbm_algorithmic_bytes": hbm_bytes, "hbm_step_bytes": hbm_step_bytes, "hbm_achieved_gbs": hbm_bytes / (kern_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs"),
                         "hbm_frac": (hbm_bytes / (kern_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                         "mrays_per_s": counts["segments"] * total_samples * frac_samples / (kern_ms * 1e-3) / 1e6,
                         # the resource ncu shows closest to saturation (profiles/): the shared-memory data pipe.  Algorithmic bytes per
                         # sample = box-pair steps x 48 B + triangle tests x 48 B + 144 B of per-triangle shading data per attempt + 216 B of
                         # slot / stack state per segment; peak = SMs x 128 B/clk x the SM clock seen during the run
                         "smem": smem_roofline(counts, total_samples * frac_samples, kern_ms, clk)},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    scene.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
