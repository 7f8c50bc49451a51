#!/bin/bash
# The measurement recipe behind profiles/r02_conv1_*: headline bench, then the ncu launch list and one full capture of the top
# kernel on a short run of the SAME command (numbers printed under ncu are never bench values).  Run under gpurun.
set -x
python bench.py > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err; tail -c 1500 gpurun_out/bench_latest.err
SHORT="python bench.py --steps 3 --warmup 3 --images 256 --e2e-images 16 --no-cpu-baseline --no-configs --no-parity-check"
$SHORT > gpurun_out/plain_latest.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_latest.csv $SHORT > gpurun_out/ncu_latest.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma2 -s 3 -c 1 -f -o gpurun_out/prof_conv1_latest $SHORT > gpurun_out/ncu_full_latest.log 2>&1
tail -3 gpurun_out/ncu_full_latest.log
