#!/bin/bash
# One-call GPU validation of the tree (run through gpurun): parity suite, smoke, bench lines,
# host-side mesh input timing on the GPU box's cores.  Most important first; each step has its
# own time limit so that a short budget still yields the earlier results.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
( time timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.txt 2>&1
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> /dev/null
timeout 150 python tools/bdf_bench.py 500 > gpurun_out/bdf_bench.txt 2>&1
tail -5 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -2; cat gpurun_out/bench_n1.json | cut -c1-400
