#!/bin/bash
# r2u (4 GPUs): result arrays on recycled host blocks — engine GPU tests, N=1 and N=4 headline bench
( timeout 300 python -m pytest tests/test_engines_gpu.py tests/test_hovernet_gpu.py tests/test_wsi_gpu.py -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1 ); echo "tests rc=$? $(tail -1 gpurun_out/r2u_tests.log)"
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2u_bench_n4.json 2> gpurun_out/r2u_bench_n4.err ); echo "bench4 rc=$?"
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err ); echo "bench1 rc=$?"
python - <<'P'
import json
for f in ("r2u_bench_n1", "r2u_bench_n4"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["e2e"])
    except Exception as ex:
        print(f, "unreadable", ex)
P
