#!/bin/bash
# quick DC-replay iteration: bit-exact tests + stage timing for both staging variants
tag=${1:-dc}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "dc_blocker or five_modems or tiny_blocks or ragged or cfg5_full" > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$tag.log
./profiles/micro/dc_chain_forms.bin > gpurun_out/chain_forms_$tag.txt 2>&1; cat gpurun_out/chain_forms_$tag.txt
for v in 1 0; do
  SDR_B200_DC_FORM=$v timeout 300 python bench.py --steps 50 --no-configs --no-cpu-baseline --noise-inputs --min-seconds 0.3 > gpurun_out/bench_${tag}_tma$v.json 2> gpurun_out/bench_${tag}_tma$v.err
  echo "bench form=$v rc=$?"; python - <<PY
import json
try:
    l=json.load(open("gpurun_out/bench_${tag}_tma$v.json"))
    print("value", l["value"], "ms/step", l["ms_per_step"], "e2e", l["e2e"]["value"], "stages", l["roofline"]["stage_ms_per_block"])
except Exception as ex:
    print("no line", ex)
PY
  tail -12 gpurun_out/bench_${tag}_tma$v.err
done
for w in cfg2 cfg3; do
  timeout 300 python bench.py --workload $w --steps 50 --no-configs --no-cpu-baseline --noise-inputs --min-seconds 0.3 > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err
  python - <<PY
import json
try:
    l=json.load(open("gpurun_out/bench_${tag}_$w.json"))
    print("$w value", l["value"], "ms/step", l["ms_per_step"], "e2e", l["e2e"]["value"], "stages", l["roofline"]["stage_ms_per_block"])
except Exception as ex:
    print("$w no line", ex)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_dc_lanes' -s 8 -c 2 \
  -o gpurun_out/prof_$tag python bench.py --steps 4 --warmup 3 --no-configs --no-cpu-baseline --min-seconds 0 --noise-inputs > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu rc=$?"
