#!/bin/bash
# r2t: final validation of the round-2 build + config 5 on one GPU with every fix
bash profiles/call_final.sh r2t
( timeout 400 python bench.py --workload wsi --patches 20000 --steps 1 --no-cpu-baseline > gpurun_out/r2t_bench_wsi20k.json 2> gpurun_out/r2t_bench_wsi20k.err ); echo "wsi rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2t_bench_wsi20k.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('detail'))"
( timeout 120 python -m tiatoolbox_b200.models.export resnet18-kather100k gpurun_out/r2t_resnet18.plan --batch 8 > gpurun_out/r2t_export.log 2>&1 ); echo "export rc=$? $(tail -1 gpurun_out/r2t_export.log)"; rm -f gpurun_out/r2t_resnet18.plan
grep -E "original vs reference|rel_l2" gpurun_out/r2t_tests.log | head
