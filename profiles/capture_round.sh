#!/bin/bash
# One gpurun call that brings back everything a round needs from ONE B200 (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash profiles/capture_round.sh r2h'
# Outputs land in gpurun_out/<tag>_*; copy what is to be judged into profiles/.  Every step has its own timeout; ncu steps come last.
# Pass "quick" as 2nd argument to skip the full test suite and the ncu --set full captures.
tag=${1:-rX}
mode=${2:-full}
out=gpurun_out
mkdir -p $out
step() { name=$1; lim=$2; shift 2; echo "== $name"; ( timeout $lim "$@" ) > $out/${tag}_$name.log 2>&1; echo "   rc=$? $(tail -1 $out/${tag}_$name.log | cut -c1-200)"; }
bench() { name=$1; lim=$2; shift 2; echo "== bench $name"; ( timeout $lim python bench.py "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err ); echo "   rc=$? $(tail -c 300 $out/${tag}_bench_$name.json)"; }

# 1. parity: the whole GPU suite (-s: the tests print their measured errors and timings)
if [ "$mode" = full ]; then step tests 600 python -m pytest tests -q -m gpu -x -s --durations=8; fi
# 2. one bench line per BASELINE config (clock sampling is inside bench.py)
bench n1 500 --steps 10 --warmup 3                                    # config 4 headline: Macenko + HoVer-Net (+ rooflines, library / cpu baselines)
bench macenko 300 --workload macenko --steps 10 --warmup 3            # config 1
bench resnet18 300 --workload resnet18 --steps 10 --warmup 3          # config 2
bench fcn 400 --workload fcn_tissue_mask --steps 5 --warmup 3         # config 3
bench wsi20k 500 --workload wsi --patches 20000 --steps 1             # config 5 on one GPU
( timeout 120 python profiles/time_postproc.py 1 4 > $out/${tag}_time_postproc.jsonl 2> $out/${tag}_time_postproc.err ); cat $out/${tag}_time_postproc.jsonl
( timeout 120 python profiles/measure_new_kernels.py > $out/${tag}_new_kernels.jsonl 2> $out/${tag}_new_kernels.err ); cut -c1-160 $out/${tag}_new_kernels.jsonl
# 3. launch list of one short bench run (shares, not absolutes)
step ncu_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/${tag}_launches.csv \
     python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline --stain-roofline-tiles 1024
[ "$mode" = full ] || { ls -la $out | tail -30; exit 0; }
# 4. full captures of the kernels that matter.  gpurun copies back at most 64 MiB: an .ncu-rep costs ~2 MB per profiled launch, so the
#    conv engine's 79 launches per forward are captured in windows by `profiles/capture_conv.sh` (separate calls), not here.
step ncu_stain 300 ncu --set full --clock-control none --import-source on -k regex:stain_fused_kernel -s 2 -c 2 -o $out/${tag}_stain_fused \
     python bench.py --workload macenko --steps 1 --warmup 1 --no-cpu-baseline
for pat in pp_watershed_v4 pp_contours pp_ccl_merge reinhard_apply_kernel reinhard_hist_kernel; do
  step ncu_$pat 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 2 -o $out/${tag}_$pat \
       python profiles/measure_new_kernels.py
done
du -sh $out; ls -la $out | tail -40
