#!/bin/bash
# r2j: config 5 after the columnar merge, ncu window of the decoder's conv launches (54..65: 3x3 1024->256 tensor-bound, dense-block 1x1 / 3x3 g4)
mkdir -p gpurun_out
( timeout 400 python bench.py --workload wsi --patches 20000 --steps 1 --no-cpu-baseline > gpurun_out/r2j_bench_wsi20k.json 2> gpurun_out/r2j_bench_wsi20k.err ); echo "wsi rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2j_bench_wsi20k.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('detail'))"
bash profiles/capture_conv.sh r2j 54 12 | tail -4
