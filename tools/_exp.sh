python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_layers.py --images 64 --only L0 --check 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['plan']); print(d['ms'],'ms',d['img_s'],'img/s', d['GBs'],'GB/s', d['checked'])"
FCB_U2_NO_WL=1 python tools/bench_layers.py --images 64 --only L0 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NO_WL', d['plan'][60:140]); print(d['ms'],'ms',d['img_s'],'img/s')"
