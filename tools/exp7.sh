python tools/bench_layers.py --images 64 --only L0 > gpurun_out/l0_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:umma2 -s 3 -c 1 -o gpurun_out/prof_l0 -f python tools/bench_layers.py --images 64 --only L0 > gpurun_out/l0_ncu.log 2>&1
tail -2 gpurun_out/l0_ncu.log
