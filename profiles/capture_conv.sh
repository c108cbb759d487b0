#!/bin/bash
# ncu --set full of a window of the HoVer-Net forward's conv_gemm_kernel launches (launch indices as in the per-launch table of
# profiles/r1e_conv_gemm.md: 0-10 d0 stage, 11-23 d1, 24-42 d2, 43-53 d3 + conv_bot, 54-72 decoder level 3, 73-85 level 2, 86-89 level 1):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/capture_conv.sh r2i 0 12'
# ~2 MB per launch in the report; gpurun copies back at most 64 MiB per call, so keep the window <= 20 launches.
tag=${1:-rX}; skip=${2:-0}; count=${3:-12}
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s $skip -c $count -o gpurun_out/${tag}_conv_gemm_s${skip}_c${count} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline --stain-roofline-tiles 1024 > gpurun_out/${tag}_conv_ncu_s${skip}.log 2>&1
echo rc=$?; du -sh gpurun_out; ls -la gpurun_out
