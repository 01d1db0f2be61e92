#!/bin/bash
# A/B of the flat scan (RT_FLAT_SCAN_MAX=12, default) against the BVH walk (RT_FLAT_SCAN_MAX=0) on the text scenes
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_general.py tests/test_text_scene.py -m gpu -x -q > gpurun_out/flat_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/flat_tests.log
for sc in practice3_1 practice3_5 practice3_3 practice3_4; do
  for env in "RT_FLAT_SCAN_MAX=0" "RT_FLAT_SCAN_MAX=12" "RT_FLAT_SCAN_MAX=12 RT_WAVE_CFG=2" "RT_FLAT_SCAN_MAX=12 RT_WAVE_CFG=1" "RT_FLAT_SCAN_MAX=12 RT_WAVE_CFG=0"; do
    echo "== $sc $env" | tee -a gpurun_out/flat_ab.txt
    env $env timeout 300 python tools/render_once.py scenes/$sc.txt 2048 2048 64 3 2>&1 | tail -1 | tee -a gpurun_out/flat_ab.txt
  done
done
