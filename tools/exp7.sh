python tools/bench_layers.py --images 256 --only cfg4 > gpurun_out/cfg4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:umma2 -s 3 -c 1 -o gpurun_out/prof_cfg4b -f python tools/bench_layers.py --images 256 --only cfg4 > gpurun_out/cfg4_ncu.log 2>&1
tail -2 gpurun_out/cfg4_ncu.log; cat gpurun_out/cfg4_plain.log | cut -c1-200
