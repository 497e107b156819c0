#!/bin/bash
# compute-sanitizer passes over tools/sanitize.py + the final full GPU suite (through gpurun)
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/final_pytest.txt 2>&1
for tool in memcheck racecheck synccheck; do
  ( time timeout 170 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize.py 2>&1 | grep -v "^$" | tail -25 ) > gpurun_out/sanitizer_$tool.txt 2>&1
done
tail -n 3 gpurun_out/final_pytest.txt; for f in gpurun_out/sanitizer_*.txt; do tail -n 4 $f; done
