#!/bin/bash
# r2k: WSI mode after the device-side instance-map paste; headline bench at 64 tiles per step; smoke(); ncu window with the 3x3 1024->256 decoder conv
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_wsi_gpu.py tests/test_hovernetplus_gpu.py tests/test_engines_gpu.py tests/test_stain_gpu.py -q -m gpu -x > gpurun_out/r2k_tests.log 2>&1 ); echo "tests rc=$? $(tail -1 gpurun_out/r2k_tests.log)"
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1 ); echo "smoke rc=$? $(tail -1 gpurun_out/r2k_smoke.log)"
( timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err ); echo "bench rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2k_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown_ms'], d['cpu_baseline']['value'], d['roofline']['frac'])"
( timeout 400 python bench.py --workload wsi --patches 20000 --steps 1 --no-cpu-baseline > gpurun_out/r2k_bench_wsi20k.json 2> gpurun_out/r2k_bench_wsi20k.err ); echo "wsi rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2k_bench_wsi20k.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('detail'))"
bash profiles/capture_conv.sh r2k 46 8 | tail -3
