#!/bin/bash
# r2l (8 GPUs of one box): weak scaling of the headline bench, strong scaling of config 5
mkdir -p gpurun_out
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err ); echo "bench8 rc=$?"
python -c "
import json
d = [json.loads(l) for l in open('gpurun_out/r2l_bench_n8.json') if l.startswith('{')][-1]; print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload wsi --patches 20000 --steps 1 > gpurun_out/r2l_bench_wsi20k_n8.json 2> gpurun_out/r2l_bench_wsi20k_n8.err ); echo "wsi8 rc=$?"
python -c "
import json
d = [json.loads(l) for l in open('gpurun_out/r2l_bench_wsi20k_n8.json') if l.startswith('{')][-1]; print(d['value'], d['detail'])"
grep -v Generating gpurun_out/r2l_bench_wsi20k_n8.err | grep -i "error\|Traceback" | head -5
