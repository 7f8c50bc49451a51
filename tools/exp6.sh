python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_layers.py --images 64 --only L7,L6 --check 2>&1 | python -c "
import sys,json
for l in sys.stdin.read().strip().splitlines():
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['layer'], d['plan'][:150]); print('   ',d['ms'],'ms',d['img_s'],'img/s', d['GBs'],'GB/s', d['TOPs_nonzero'],'TOPs nz', d['checked'])"
FCB_U2_PROF=1 python tools/bench_layers.py --images 64 --only L7 2>&1 | grep "u2 prof" | tail -3
