python -m pytest tests -m gpu -x -q -k "xn or xnor or random" 2>&1 | tail -3
python tools/bench_layers.py --images 1024 --only cfg3 --check 2>&1 | cut -c1-260
