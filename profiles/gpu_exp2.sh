#!/bin/bash
# experiment: what delays the warm-up DC replay next to the spectrum (cfg5), tail phases
tag=${1:-exp2}
mkdir -p gpurun_out
run() { # name, env..., -- args
  name=$1; shift
  env "$@" > /dev/null 2>&1
}
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
  grep -A3 "device timeline" gpurun_out/bench_${tag}_$name.err | tail -2
}
B="python bench.py --workload cfg5 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.4"
one default $B
one nospec $B --fft 0
one fft4096 $B --fft 4096
SDR_B200_DC_SMS=8 one sms8 $B
SDR_B200_DC_TMA=0 one cpasync $B
SDR_B200_DC_WINDOW=1 one window $B
one ch256 $B --channels 256
one ch128 $B --channels 128
one ch64 $B --channels 64
timeout 300 python profiles/tail_phases.py > gpurun_out/tail_phases_$tag.txt 2>&1; tail -30 gpurun_out/tail_phases_$tag.txt
