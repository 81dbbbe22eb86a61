#!/bin/bash
tag=${1:-exp5}
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
B="python bench.py --workload cfg5 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.3"
for sms in 8 24 48 74; do
  SDR_B200_DC_SMS=$sms one ch64_s$sms $B --channels 64
done
for sms in 16 32; do
  SDR_B200_DC_SMS=$sms one ch128_s$sms $B --channels 128
  SDR_B200_DC_SMS=$sms one ch256_s$sms $B --channels 256
done
SDR_B200_DC_SMS=16 one cfg3_s16 python bench.py --workload cfg3 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.3
SDR_B200_DC_SMS=32 one cfg3_s32 python bench.py --workload cfg3 --steps 50 --no-configs --no-cpu-baseline --min-seconds 0.3
