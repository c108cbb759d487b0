#!/bin/bash
# r2i: the reference arm on the GPU box's host cores, config 5 with its wall-clock split, one ncu window of the conv engine
mkdir -p gpurun_out
last() { python -c "
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d.get('value'), d.get('detail'), str(d.get('cpu_baseline', {}).get('sample'))[:300])" "$1"; }
( timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2i_ref_arm.json 2> gpurun_out/r2i_ref_arm.err ); echo "ref rc=$?"; last gpurun_out/r2i_ref_arm.json
( timeout 400 python bench.py --workload wsi --patches 20000 --steps 1 --no-cpu-baseline > gpurun_out/r2i_bench_wsi20k.json 2> gpurun_out/r2i_bench_wsi20k.err ); echo "wsi rc=$?"; last gpurun_out/r2i_bench_wsi20k.json
bash profiles/capture_conv.sh r2i 0 12 | tail -4
