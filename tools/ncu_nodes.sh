#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
for f in ${FORMATS:-2}; do
  RT_NODE_FORMAT=$f RT_WAVE_CFG=${CFG:-2} timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 1 -c 1 -o gpurun_out/nodes_fmt$f -f \
    python tools/sweep.py --scene practice7_2 --leaf 2 --cost 2.0 --variants 30 --width 1280 --height 720 --spp 32 --reps 1 > gpurun_out/ncu_nodes_fmt$f.log 2>&1
  echo "fmt $f rc=$?"
done
