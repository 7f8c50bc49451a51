python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_layers.py --images 64 --check 2>&1 | python -c "
import sys,json
for l in sys.stdin.read().strip().splitlines():
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    if d['layer'].startswith('L') : continue
    print(d['layer'], d.get('ms'), 'ms', d.get('img_s'), 'img/s', d.get('TOPs_nonzero', d.get('TOPs')), 'TOPs', d.get('checked'), (d.get('plan') or '')[:100])"
