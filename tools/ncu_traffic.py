"""Turns `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` output of a bench.py run into
profiles/<out>.json: per-launch DRAM traffic of the render kernel at the headline workload, tagged with the hash of the kernel
sources it was captured from (bench.py quotes it as roofline.traffic only while that hash matches the build being timed).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:render_wave \
        --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline
    python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/r2_bench_traffic.json "<the command above>"
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha16  # noqa: E402


def main():
    src, out, command = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    launches = {}
    for r in rows:
        if "render_wave_kernel" not in r.get("Kernel Name", ""):
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "second": 1e9,
                 "nsecond": 1.0}.get(unit, 1.0)
        launches.setdefault(r["ID"], {"kernel": r["Kernel Name"]})[r["Metric Name"]] = val * scale
    full = [v for v in launches.values() if v.get("gpu__time_duration.sum", 0) > 1e9]      # the full-frame launches (> 1 s), not the stats / probe ones
    if not full:
        raise SystemExit("no full-frame render launch in " + src)
    pick = full[-1]
    res = {"command": command, "kernel": pick["kernel"], "kernel_source_sha16": kernel_source_sha16(),
           "dram__bytes_read.sum": pick["dram__bytes_read.sum"], "dram__bytes_write.sum": pick["dram__bytes_write.sum"],
           "gpu__time_duration.sum_ns": pick["gpu__time_duration.sum"], "n_full_frame_launches": len(full),
           "all_launches": [{k: v for k, v in x.items()} for x in launches.values()]}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "all_launches"}))


if __name__ == "__main__":
    main()
