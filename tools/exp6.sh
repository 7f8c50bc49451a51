python -m pytest tests -m gpu -x -q -k "thin_output or dc_ or random or net" 2>&1 | tail -12 | cut -c1-400
python tools/bench_layers.py --images 64 --only L7 --check 2>&1 | cut -c1-400
FCB_U2_NO_DCOL=1 python tools/bench_layers.py --images 64 --only L7 2>&1 | cut -c1-100,250-400
