python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "L0 prof"; FCB_U2_PROF=1 python tools/bench_layers.py --images 64 --only L0 2>&1 | grep "u2 prof" | tail -3
python tools/bench_layers.py --images 64 --only L0,L5,L6 2>&1 | python -c "
import sys,json
for l in sys.stdin.read().strip().splitlines():
    d=json.loads(l); print(d['layer'], d['plan'][:150]); print('   ',d['ms'],'ms',d['img_s'],'img/s', d['GBs'],'GB/s', d['TOPs_nonzero'],'TOPs nz')"
FCB_U2_NO_ALT=1 python tools/bench_layers.py --images 64 --only L0,L6 2>&1 | python -c "
import sys,json
for l in sys.stdin.read().strip().splitlines():
    d=json.loads(l); print('NO_ALT',d['layer'],d['ms'],'ms',d['img_s'],'img/s')"
