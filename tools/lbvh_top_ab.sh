#!/bin/bash
# GPU LBVH with and without the SAH rebuild of the top of the tree (RT_BVH_TOP_SAH = number of subtrees), against the host SAH builder
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bvh_builder" 2>&1 | tail -2
for SCENE in practice7_2 practice7_3; do
  echo "== $SCENE host SAH (leaf 2)" | tee -a gpurun_out/lbvh_top.txt
  timeout 300 python tools/sweep.py --scene $SCENE --width 1920 --height 1080 --spp 64 --leaf 2 --cost 2.0 --variants 30 --reps 2 2>&1 | tee -a gpurun_out/lbvh_top.txt
  for top in 0 512 4096 16384; do
    echo "== $SCENE GPU LBVH (leaf 4) top SAH $top" | tee -a gpurun_out/lbvh_top.txt
    RT_BVH_BUILDER=gpu RT_BVH_TOP_SAH=$top timeout 300 python tools/sweep.py --scene $SCENE --width 1920 --height 1080 --spp 64 --leaf 4 --cost 2.0 --variants 30 --reps 2 2>&1 | tee -a gpurun_out/lbvh_top.txt
  done
done
