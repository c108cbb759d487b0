#!/bin/bash
# final validation of a build: full GPU suite, smoke(), default bench (all legs), reference arm
tag=${1:-final}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu -x -s --durations=5 > gpurun_out/${tag}_tests.log 2>&1 ); echo "tests rc=$? $(tail -1 gpurun_out/${tag}_tests.log)"
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1 ); echo "smoke rc=$? $(tail -1 gpurun_out/${tag}_smoke.log)"
( timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err ); echo "bench rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/${tag}_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown_ms'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline_stain']['frac'], d['library_baseline'].get('forward_ms_tf32'))"
( timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref_arm.json 2> gpurun_out/${tag}_ref_arm.err ); echo "ref rc=$? $(cut -c1-160 gpurun_out/${tag}_ref_arm.json)"
