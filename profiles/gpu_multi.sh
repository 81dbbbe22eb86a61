#!/bin/bash
# multi-GPU call: group parity on hardware + the bench line at N ranks (launched exactly like the driver does)
N=${1:-2}; tag=${2:-mg}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_group.py -m gpu -q -x > gpurun_out/pytest_group_$tag.log 2>&1; echo "group tests rc=$?"; tail -8 gpurun_out/pytest_group_$tag.log
for n in $(seq 2 $N); do
  if [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_n$n.json 2> gpurun_out/bench_${tag}_n$n.err
    echo "bench N=$n rc=$?"; python - <<PY
import json
try:
    l=json.load(open("gpurun_out/bench_${tag}_n$n.json"))
    print("N=$n value", l["value"], "ms/step", l["ms_per_step"], "e2e", l["e2e"]["value"], "stages", l["roofline"]["stage_ms_per_block"])
except Exception as ex:
    print("no line", ex)
PY
    tail -5 gpurun_out/bench_${tag}_n$n.err
  fi
done
