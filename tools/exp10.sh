L=simple_image_compression_network_b200/libfinnconv_b200.so
cp $L /tmp/new.so; cp tools/libfinnconv_prof.so $L
cat > /tmp/st1.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import bench_layers as B
from simple_image_compression_network_b200.desc import ACT_THRESHOLDS, KIND_CONV, LayerDesc
def stage(c, ofm, x, y, simd, pe, pool=2):
    return LayerDesc(kind=KIND_CONV, kernel_x=3, kernel_y=3, ifm_ch=c, ofm_ch=ofm, ifm_x=x, ifm_y=y, stride_x=1, stride_y=1, pad=1,
                     simd=simd, pe=pe, in_bits=8, w_bits=4, acc_bits=24, acc_signed=1, act_kind=ACT_THRESHOLDS, out_bits=8, num_th=255, pool=pool)
which = sys.argv[1]
if which == "s1": B.bench_layer("stage1", stage(3, 128, 768, 512, 3, 16), 32, 0xFF)
if which == "s2": B.bench_layer("stage2", stage(128, 128, 384, 256, 32, 16), 64, 0xFF)
P
for w in s1 s2; do for D in 0 16; do echo "== $w DEBUG=$D"; FCB_U2_DEBUG=$D FCB_U2_PROF=1 python /tmp/st1.py $w 2>&1 | grep -E "u2 prof|img_s" | tail -4 | cut -c1-200; done; done
cp /tmp/new.so $L
