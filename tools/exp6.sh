python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for E in FCB_U2_EPI4=0 FCB_U2_EPI4=1 X=1; do echo "== $E"; env $E python tools/bench_layers.py --images 256 --only L1,L2,L3,L4 2>&1 | python -c "
import sys,json
for l in sys.stdin.read().strip().splitlines():
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['layer'], d.get('ms'), 'ms', d.get('img_s'), 'img/s', d.get('TOPs_nonzero', d.get('TOPs')), 'TOPs', d['plan'][40:110])"; done
