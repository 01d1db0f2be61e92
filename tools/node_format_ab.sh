#!/bin/bash
# A/B of the node formats for global-memory scenes: RT_NODE_FORMAT 0 = octant pair nodes (LDG.128), 1 = 4-wide, 2 = compact + LDG.256
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_general.py -m gpu -x -q -k "hit_parity or converged_image or bvh_builder or kernel_variants or placement or native_size" > gpurun_out/nodes_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/nodes_tests.log
tail -3 gpurun_out/nodes_tests.log
run() { echo "== $SCENE $*" | tee -a gpurun_out/nodes_ab.txt; env "$@" timeout 600 python tools/sweep.py --scene $SCENE --width 1920 --height 1080 --spp 64 --leaf 2 --cost 2.0 --variants 30 --reps 2 2>&1 | tee -a gpurun_out/nodes_ab.txt; }
for SCENE in practice7_2 practice7_3; do
  export SCENE
  run RT_NODE_FORMAT=0
  run RT_NODE_FORMAT=2
  run RT_NODE_FORMAT=2 RT_WAVE_CFG=2


done
for f in 0 2; do echo "== working.txt format $f" | tee -a gpurun_out/nodes_ab.txt; RT_NODE_FORMAT=$f timeout 300 python tools/render_once.py scenes/working.txt 0 0 0 3 2>&1 | tail -1 | tee -a gpurun_out/nodes_ab.txt; done
