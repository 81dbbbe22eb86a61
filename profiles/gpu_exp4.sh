#!/bin/bash
tag=${1:-exp4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "dc_ or cfg4_all or cfg5_full or five_modems" > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; grep -E "re-run|passed|failed|Error" gpurun_out/pytest_$tag.log | tail -6
one() {
  name=$1; shift
  timeout 300 "$@" > gpurun_out/bench_${tag}_$name.json 2> gpurun_out/bench_${tag}_$name.err
  python - <<PY
import json
try:
    l=json.load(open("gpurun_out/bench_${tag}_$name.json"))
    print("$name value %.1f ms/step %.4f e2e %.1f stages %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], {k: round(v,4) for k,v in l["roofline"]["stage_ms_per_block"].items()}))
except Exception as ex:
    print("$name no line", ex)
PY
}
B="python bench.py --workload cfg5 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4"
one default $B
one ch256 $B --channels 256
one ch128 $B --channels 128
SDR_B200_DC_WINDOW=1 one ch128_win $B --channels 128
one ch64 $B --channels 64
SDR_B200_DC_WINDOW=1 one ch64_win $B --channels 64
one cfg4 python bench.py --workload cfg4 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4
one cfg3 python bench.py --workload cfg3 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/launches_${tag}_ch128.csv \
    python bench.py --workload cfg5 --channels 128 --steps 4 --warmup 3 --no-configs --no-cpu-baseline --min-seconds 0 --noise-inputs > gpurun_out/ncu_launch_${tag}.log 2>&1
python profiles/launch_summary.py gpurun_out/launches_${tag}_ch128.csv | head -18
