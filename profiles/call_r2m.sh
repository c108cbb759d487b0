#!/bin/bash
# r2m: config 5 after the scatter fix, headline bench at the default batch, ncu --set full of the non-conv kernels (kept under gpurun's 64 MiB return limit)
mkdir -p gpurun_out
( timeout 400 python bench.py --workload wsi --patches 20000 --steps 1 --no-cpu-baseline > gpurun_out/r2m_bench_wsi20k.json 2> gpurun_out/r2m_bench_wsi20k.err ); echo "wsi rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2m_bench_wsi20k.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('detail'))"
( timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err ); echo "bench rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2m_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'])"
step() { name=$1; lim=$2; shift 2; ( timeout $lim "$@" ) > gpurun_out/r2m_$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/r2m_$name.log | cut -c1-120)"; }
step ncu_stain 300 ncu --set full --clock-control none --import-source on -k regex:stain_fused_kernel -s 2 -c 1 -o gpurun_out/r2m_stain_fused \
     python bench.py --workload macenko --steps 1 --warmup 1 --no-cpu-baseline
for pat in pp_watershed_v4 pp_contours pp_ccl_merge reinhard_apply_kernel reinhard_hist_kernel; do
  step ncu_$pat 300 ncu --set full --clock-control none -k regex:$pat -s 2 -c 1 -o gpurun_out/r2m_$pat python profiles/measure_new_kernels.py
done
du -sh gpurun_out; ls -la gpurun_out | tail -20
