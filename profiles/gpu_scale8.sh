#!/bin/bash
# the driver's scaling run, N = 8 only: torch.distributed.run, one rank per GPU, bench.py --gpus 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_scale_n8.json 2> gpurun_out/bench_scale_n8.err
echo "rc=$?"
python - <<PY
import json
try:
    l=json.load(open("gpurun_out/bench_scale_n8.json"))
    print("N=8 value", l["value"], "ms/step", l["ms_per_step"], "e2e", l["e2e"]["value"], "stages", l["roofline"]["stage_ms_per_block"])
except Exception as ex:
    print("no line", ex)
PY
tail -8 gpurun_out/bench_scale_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_scale_n4.json 2> gpurun_out/bench_scale_n4.err
python - <<PY
import json
try:
    l=json.load(open("gpurun_out/bench_scale_n4.json"))
    print("N=4 value", l["value"], "ms/step", l["ms_per_step"], "e2e", l["e2e"]["value"], "stages", l["roofline"]["stage_ms_per_block"])
except Exception as ex:
    print("no line", ex)
PY
