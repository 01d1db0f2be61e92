#!/bin/bash
# Renders the headline frame with the CLI on N and N/2 GPUs of one box (single process, RT_GPUS) and compares the bytes.
# usage: tools/cli_multi_gpu_check.sh <n_gpus> [spp]
set -u
N=${1:-2}; SPP=${2:-1024}
E=raytracing-course-2024_b200/_build/raytracing-engine
mkdir -p gpurun_out
for g in $N $((N / 2)); do
  RT_GPUS=$g timeout 150 $E scenes/practice7_4.gltf 3840 2160 $SPP gpurun_out/cli_m$g.ppm | grep -E "msamples|error|took"
done
python - <<PY
import numpy as np
n = 3840 * 2160 * 3
a = np.fromfile("gpurun_out/cli_m$N.ppm", dtype=np.uint8)[-n:].astype(int)
b = np.fromfile("gpurun_out/cli_m$((N / 2)).ppm", dtype=np.uint8)[-n:].astype(int)
print("max byte diff", int(np.abs(a - b).max()), "fraction differing", float((a != b).mean()))
PY
rm -f gpurun_out/cli_m*.ppm
