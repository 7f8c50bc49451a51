python bench.py --steps 3 --warmup 3 --images 256 --e2e-images 16 --no-cpu-baseline > gpurun_out/plain_r01d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:umma2 -s 3 -c 1 -f -o gpurun_out/prof_conv1_r01d python bench.py --steps 3 --warmup 3 --images 256 --e2e-images 16 --no-cpu-baseline > gpurun_out/ncu_full_r01d.log 2>&1
tail -1 gpurun_out/ncu_full_r01d.log
