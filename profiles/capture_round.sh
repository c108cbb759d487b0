#!/bin/bash
# One gpurun call that brings back everything a round needs from ONE B200 (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/capture_round.sh r2a'
# Outputs land in gpurun_out/<tag>_*; copy what is to be judged into profiles/.  Every step has its own timeout; ncu steps come last.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
step() { name=$1; lim=$2; shift 2; echo "== $name"; ( timeout $lim "$@" ) > $out/${tag}_$name.log 2>&1; echo "   rc=$? $(tail -1 $out/${tag}_$name.log | cut -c1-200)"; }

# 1. parity: the whole GPU suite (established modules first, newest last: tests/conftest.py)
step tests 600 python -m pytest tests -q -m gpu -x --durations=8
# 2. flood implementations 1-4 incl. the adversarial tie floods on the device + timings (printed with -s)
step floods 240 python -m pytest tests/test_watershed_v2_gpu.py -q -m gpu -s
# 3. benches (clock sampling is inside bench.py)
( timeout 400 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err ); tail -c 400 $out/${tag}_bench_n1.json
( timeout 200 python bench.py --workload macenko --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_macenko.json 2>> $out/${tag}_bench_n1.err )
( timeout 120 python profiles/measure_new_kernels.py > $out/${tag}_new_kernels.jsonl 2> $out/${tag}_new_kernels.err ); cat $out/${tag}_new_kernels.jsonl | cut -c1-160
( timeout 300 python bench.py --workload wsi --patches 2000 --no-cpu-baseline > $out/${tag}_bench_wsi_2000.json 2> $out/${tag}_bench_wsi.err ); tail -c 300 $out/${tag}_bench_wsi_2000.json
# 4. launch list of one short bench run (shares, not absolutes)
step ncu_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
     python bench.py --steps 2 --warmup 1 --no-cpu-baseline
# 5. full captures of the kernels that matter (3 launches each)
for pat in conv_gemm_kernel stain_fused_kernel; do      # launched by the headline bench
  step ncu_$pat 400 ncu --set full --clock-control none --import-source on -k regex:$pat -s 4 -c 3 -o $out/${tag}_$pat \
       python bench.py --steps 1 --warmup 1 --no-cpu-baseline
done
for pat in reinhard_apply_kernel reinhard_hist_kernel pp_watershed pp_contours; do      # launched by the new-kernel timing script
  step ncu_$pat 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 4 -c 3 -o $out/${tag}_$pat \
       python profiles/measure_new_kernels.py
done
ls -la $out | tail -30
