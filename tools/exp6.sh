python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_layers.py --images 64 --only L0,L1,L2,L5,L6,L7,cfg4 2>&1 | python -c "
import sys,json
for l in sys.stdin.read().strip().splitlines():
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['layer'], d.get('ms'), 'ms', d.get('img_s'), 'img/s', d.get('TOPs_nonzero', d.get('TOPs')), 'TOPs')"
