#!/bin/bash
# r2n: side-stream D2H of the per-batch records, WSI paste without the slot-0 write contention
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_engines_gpu.py tests/test_contours_gpu.py tests/test_wsi_gpu.py -q -m gpu -x > gpurun_out/r2n_tests.log 2>&1 ); echo "tests rc=$? $(tail -1 gpurun_out/r2n_tests.log)"
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err ); echo "bench rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2n_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
( timeout 400 python bench.py --workload wsi --patches 20000 --steps 1 --no-cpu-baseline > gpurun_out/r2n_bench_wsi20k.json 2> gpurun_out/r2n_bench_wsi20k.err ); echo "wsi rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2n_bench_wsi20k.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('detail'))"
