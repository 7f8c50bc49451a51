python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_v21.log; cat gpurun_out/pytest_v21.log
python tools/bench_layers.py --images 64 --only L0 2>&1 | tee gpurun_out/layers14.log | cut -c1-400
