# A/B: steps without host I/O split into element ranges so that the matrices are zeroed next to
# the kernels (A2DS_STREAM_RESIDENT=1) against the memsets in front of one launch (development aid)
cd "$(dirname "$0")/.."
echo "== default (memsets in front)"; A2DS_STREAM_RESIDENT=0 timeout 120 python tools/quick_bench.py 1000 2>&1 | grep -E "^res|^K |^G |all\(|jac\("
for ch in 4 8 16; do
echo "== resident ranges, $ch chunks"; A2DS_STREAM_RESIDENT=1 A2DS_STREAM_CHUNKS=$ch timeout 120 python tools/quick_bench.py 1000 2>&1 | grep -E "^res|^K |^G |all\(|jac\("
done
