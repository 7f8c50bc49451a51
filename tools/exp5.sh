python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_v20.log; cat gpurun_out/pytest_v20.log
python tools/bench_layers.py --images 64 --only L0,L1,L2,L5,L6 2>&1 | tee gpurun_out/layers13.log | cut -c1-330
FCB_U2_NO_STAGE=1 python tools/bench_layers.py --images 64 --only L0,L5,L6 2>&1 | cut -c1-330
