run() { echo "== $*"; env "$@" python tools/kbench.py --snps 20920 2>&1 | tail -2 | head -1 | cut -c1-100; }
run VAFGPU_SPAN_CAP=96
run VAFGPU_SPAN_CAP=128
run VAFGPU_SPAN_CAP=192
run VAFGPU_SPAN_CAP=256
run VAFGPU_SPAN_CAP=1024
echo == 1k cap128; VAFGPU_SPAN_CAP=128 python tools/kbench.py --snps 1000 2>&1 | tail -2 | head -1 | cut -c1-100
echo == 1k cap256; VAFGPU_SPAN_CAP=256 python tools/kbench.py --snps 1000 2>&1 | tail -2 | head -1 | cut -c1-100
echo == k31 cap128; VAFGPU_SPAN_CAP=128 python tools/kbench.py --snps 20920 --k 31 2>&1 | tail -2 | head -1 | cut -c1-100
echo == k31 cap256; VAFGPU_SPAN_CAP=256 python tools/kbench.py --snps 20920 --k 31 2>&1 | tail -2 | head -1 | cut -c1-100
echo == small 64MB; python tools/kbench.py --snps 20920 --mb 64 2>&1 | tail -2 | head -1 | cut -c1-100
