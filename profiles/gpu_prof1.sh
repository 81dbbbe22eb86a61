#!/bin/bash
# one full ncu capture of a named kernel in a named workload: bash profiles/gpu_prof1.sh <tag> <kernel regex> <workload> [skip] [count]
tag=$1; k=$2; w=${3:-cfg5}; skip=${4:-4}; cnt=${5:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c $cnt -o gpurun_out/prof_$tag \
  python bench.py --workload $w --steps 4 --warmup 3 --no-configs --no-cpu-baseline --min-seconds 0 --noise-inputs > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full_$tag.log; ls -la gpurun_out/prof_$tag.ncu-rep
