#!/bin/bash
tag=${1:-exp3}
mkdir -p gpurun_out
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
one ch256_old $B --channels 256
SDR_B200_DC_WINDOW=1 one ch256_win $B --channels 256
one cfg4_old python bench.py --workload cfg4 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4
SDR_B200_DC_WINDOW=1 one cfg4_win python bench.py --workload cfg4 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4
one cfg2 python bench.py --workload cfg2 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4
one chz python bench.py --workload cfg5 --channelizer --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4
