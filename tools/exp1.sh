# perf decomposition of the umma_i8 resident-planes kernel (results wrong by design for DEBUG != 0)
run() { echo -n "$1 => "; env $1 BENCH_NOCHECK=1 python bench.py --images 1024 --steps 5 --warmup 3 --e2e-images 16 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3),'ms', round(d['value']),'img/s', d['roofline']['kernel'])"; }
run "X=0"
run "FCB_U2_DEBUG=1"
run "FCB_U2_DEBUG=2"
run "FCB_U2_DEBUG=3"
run "FCB_U2_DEBUG=4"
run "FCB_U2_DEBUG=8"
run "FCB_U2_DEBUG=11"
run "FCB_U2_FORCE=48,2,128,6"
run "FCB_U2_FORCE=30,8,256,4"
run "FCB_UMMA_V1=1"
