cd "$(dirname "$0")/.."
echo "== memset (A2DS_INKERNEL_ZERO=0)"; A2DS_INKERNEL_ZERO=0 timeout 120 python tools/quick_bench.py 1000 2>&1 | grep -E "^res|^K |^G |all\(|jac\("
for la in 2 3 5 8; do
echo "== in-kernel zeroing, ahead $la"; A2DS_ZERO_AHEAD=$la timeout 120 python tools/quick_bench.py 1000 2>&1 | grep -E "^K |^G |all\(|jac\("
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
