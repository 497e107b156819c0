#!/bin/bash
# end-to-end rate of bench.py against the number of element ranges of the streamed assembly
# (A2DS_STREAM_CHUNKS) — development aid, run through gpurun
mkdir -p gpurun_out; out=gpurun_out/stream_chunks_ab.txt; : > $out
for c in ${@:-4 6 8 10 12 16}; do
  A2DS_STREAM_CHUNKS=$c python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-parity 2>/dev/null |
    python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks', $c, 'value', round(d['value'] / 1e6, 2), 'e2e', round(d['e2e']['value'] / 1e6, 2))" >> $out
done
cat $out
