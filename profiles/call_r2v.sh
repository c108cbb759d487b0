#!/bin/bash
# r2v (8 GPUs): headline bench with the recycled result blocks
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2v_bench_n8.json 2> gpurun_out/r2v_bench_n8.err ); echo "bench8 rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2v_bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"
