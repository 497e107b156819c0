cd /root/repo
echo "== default (memset)"; timeout 120 python tools/quick_bench.py 1000 2>&1 | grep -E "^K |^G |all\(|jac\("
echo "== ikz build, runtime off"; A2DS_LIB=a2d-shells_b200/lib/variants/liba2ds_ikz.so A2DS_INKERNEL_ZERO=0 timeout 120 python tools/quick_bench.py 1000 2>&1 | grep -E "^K |^G |all\(|jac\("
for la in 2 4; do echo "== ikz on, ahead $la"; A2DS_LIB=a2d-shells_b200/lib/variants/liba2ds_ikz.so A2DS_ZERO_AHEAD=$la timeout 120 python tools/quick_bench.py 1000 2>&1 | grep -E "^K |^G |all\(|jac\("; done
echo "== ikz parity"; A2DS_LIB=a2d-shells_b200/lib/variants/liba2ds_ikz.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
