#!/bin/bash
# Round-end evidence on one B200: tests, bench (both arms), ncu launch list + DRAM traffic of the timed build, spp fit, results tables.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
python -m pytest tests -m gpu -q -s > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_gpu_tests.log
# DRAM traffic of the render kernel of THIS build first, so that the bench line below can quote it (bench.py checks the kernel-source hash)
CMD='ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:render_wave --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline'
$CMD > gpurun_out/traffic_run.log 2>&1; echo "ncu traffic rc=$?"
python tools/ncu_traffic.py gpurun_out/traffic.csv gpurun_out/r2_bench_traffic.json "$CMD" | cut -c1-400
cp gpurun_out/r2_bench_traffic.json profiles/r2_bench_traffic.json
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err; echo "ref arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launches_run.log 2>&1; echo "ncu launches rc=$?"
python tools/spp_fit.py > gpurun_out/r2_spp_fit.txt 2>&1; tail -1 gpurun_out/r2_spp_fit.txt
python tests/tools/text_bench.py > gpurun_out/r2_text_scenes.jsonl 2>&1; echo "text bench rc=$?"
python tests/tools/results_table.py > gpurun_out/r2_results_table.jsonl 2>&1; echo "results table rc=$?"; cut -c1-260 gpurun_out/r2_results_table.jsonl
