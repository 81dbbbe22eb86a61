#!/bin/bash
# experiment call: quick parity subset, bench lines of several workloads with device timelines, ncu launch lists of the small configs
tag=${1:-exp}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "${2:-spectrum or cfg2 or cfg5_full or zoomed or dc_ or five_modems or several}" > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$tag.log
for w in cfg5 cfg3 cfg2 cfg1; do
  timeout 300 python bench.py --workload $w --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4 > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err
  python - <<PY
import json
try:
    l=json.load(open("gpurun_out/bench_${tag}_$w.json"))
    print("$w value %.1f ms/step %.4f e2e %.1f launches/step %.1f stages %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l["gpu_launches"]/l["steps"], {k: round(v,4) for k,v in l["roofline"]["stage_ms_per_block"].items()}))
except Exception as ex:
    print("$w no line", ex)
PY
  grep -A6 "device timeline" gpurun_out/bench_${tag}_$w.err | head -8
done
SDR_B200_DC_WINDOW=1 timeout 300 python bench.py --workload cfg5 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4 > gpurun_out/bench_${tag}_cfg5_win.json 2> gpurun_out/bench_${tag}_cfg5_win.err
python - <<PY
import json
try:
    l=json.load(open("gpurun_out/bench_${tag}_cfg5_win.json"))
    print("cfg5 (windows forced) value %.1f ms/step %.4f e2e %.1f stages %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], {k: round(v,4) for k,v in l["roofline"]["stage_ms_per_block"].items()}))
except Exception as ex:
    print("no line", ex)
PY
for w in cfg3 cfg2; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/launches_${tag}_$w.csv \
    python bench.py --workload $w --steps 4 --warmup 3 --no-configs --no-cpu-baseline --min-seconds 0 --noise-inputs > gpurun_out/ncu_launch_${tag}_$w.log 2>&1
  echo "ncu $w rc=$?"
  python profiles/launch_summary.py gpurun_out/launches_${tag}_$w.csv | head -16
done
