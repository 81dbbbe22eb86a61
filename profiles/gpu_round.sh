#!/bin/bash
# One gpurun call of round 2: GPU parity tests, the bench line, the ncu launch list and full captures of the top kernels.
# Usage (from the repo root, under gpurun): bash profiles/gpu_round.sh <tag> [tests|notests]
tag=${1:-r02}
mkdir -p gpurun_out
if [ "${2:-tests}" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_$tag.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
  tail -25 gpurun_out/pytest_$tag.log
fi
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_$tag.json; tail -30 gpurun_out/bench_$tag.err
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 4 --warmup 3 --no-configs --no-cpu-baseline --min-seconds 0 --noise-inputs > gpurun_out/ncu_launch_$tag.log 2>&1
echo "ncu launches rc=$?"
# full captures of the top kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dc_lanes|k_dc_window|k_dc_compose|k_dc_resolve|k_chan_front|k_chan_tail|k_spec_fft' -s 60 -c 8 \
  -o gpurun_out/prof_$tag python bench.py --steps 4 --warmup 3 --no-configs --no-cpu-baseline --min-seconds 0 --noise-inputs > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
