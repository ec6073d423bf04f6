#!/bin/bash
# Run on the GPU box:  gpurun --timeout 1500 -- 'bash profiles/collect.sh'
# Collects the evidence bench.py's numbers rest on into gpurun_out/ (summarised into profiles/ afterwards
# with profiles/summarize_ncu.py on the CPU box).  Nothing measured under ncu is reported as a bench value.
set -x
mkdir -p gpurun_out
# 1. launch list of the bench command (config 2 at 1/16 of its layers: same launch geometry, 8192 groups per launch)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --scale 0.0625 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
# 2. full captures of the two dominant kernels at the bench's launch size
ncu --set full --clock-control none --import-source on -k regex:^compress_fast -s 6 -c 1 -f -o gpurun_out/prof_compress \
    python bench.py --scale 0.0625 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^decompress_fast -s 6 -c 1 -f -o gpurun_out/prof_decompress \
    python bench.py --scale 0.0625 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_d.log 2>&1
# 3. bench lines (not under a profiler)
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
kill $SMI
python bench.py --workload cfg3 --steps 10 --no-cpu > gpurun_out/bench_cfg3.json 2>&1
python bench.py --workload cfg4p --steps 5 --no-cpu > gpurun_out/bench_cfg4p.json 2>&1
python bench.py --workload cfg4 > gpurun_out/bench_cfg4.json 2>&1
python bench.py --workload cfg5 > gpurun_out/bench_cfg5.json 2>&1
python bench.py --workload ratios > gpurun_out/bench_ratios.json 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1
