# launch list of the thin layers (kernel shares): L0 (im2col + 1x1), L7 (deconv OFM=3)
python tools/bench_layers.py --images 64 --only L0,L7 > gpurun_out/thin_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_thin.csv python tools/bench_layers.py --images 64 --only L0,L7 > gpurun_out/thin_ncu.log 2>&1
cat gpurun_out/thin_plain.log
