python -m pytest tests -m gpu -x -q 2>&1 | tail -3
L=simple_image_compression_network_b200/libfinnconv_b200.so
cp $L /tmp/new.so
run() { python tools/bench_layers.py --images 256 --only L1,L2,L6,L7 2>&1 | python -c "
import sys,json
for l in sys.stdin.read().strip().splitlines():
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('  ',d['layer'], d.get('ms'), 'ms', d.get('img_s'), 'img/s')"; }
for i in 1 2; do
echo "new:"; cp /tmp/new.so $L; run
echo "old (session start, d3d5ab0):"; cp tools/libfinnconv_old.so $L; run
done
cp /tmp/new.so $L
