set -x
python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; tail -c 3000 gpurun_out/bench_r01c.json
python bench.py --steps 3 --warmup 3 --images 256 --e2e-images 16 --no-cpu-baseline > gpurun_out/plain_r01c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 3 --warmup 3 --images 256 --e2e-images 16 --no-cpu-baseline > gpurun_out/ncu_r01c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma2 -s 3 -c 1 -f -o gpurun_out/prof_conv1_r01c python bench.py --steps 3 --warmup 3 --images 256 --e2e-images 16 --no-cpu-baseline > gpurun_out/ncu_full_r01c.log 2>&1
tail -3 gpurun_out/ncu_full_r01c.log
