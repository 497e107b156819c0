# A/B of the default library against every variant in lib/variants (fused + single kernels), and
# the GPU parity file with each variant (development aid; run through gpurun)
cd "$(dirname "$0")/.."
echo "== default"; timeout 120 python tools/quick_bench.py 1000 2>&1 | grep -E "^res|^K |^G |all\(|jac\("
for lib in a2d-shells_b200/lib/variants/liba2ds_*.so; do
  [ -e "$lib" ] || continue
  echo "== $(basename $lib)"; A2DS_LIB=$lib timeout 120 python tools/quick_bench.py 1000 2>&1 | grep -E "^res|^K |^G |all\(|jac\("
  A2DS_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
done
