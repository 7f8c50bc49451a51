python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for E in X=1 FCB_U2_NO_CHB=1; do for L in cfg4 cfg4n; do echo -n "$E $L: "; env $E python tools/bench_layers.py --images 512 --only $L 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms'],'ms',d['img_s'],'img/s',d['plan'][:110])"; done; done
