#!/bin/bash
# One multi-GPU measurement session on an N-GPU box:  tools/multi_session.sh N TAG [sweep args]
# bench (weak + strong scaling), polygon overlay over N ranks (BASELINE.json configs[3]) and the
# scalability sweep (configs[4]); everything lands in gpurun_out/.
N=$1; TAG=$2; shift 2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
$RUN --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
tail -n 9 gpurun_out/bench_${TAG}_${N}gpu.err | grep "rank"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_${N}gpu.json"))
print({k: d.get(k) for k in ["n_gpus", "value", "ms_per_step", "step_ms", "e2e", "strong_scaling", "parity_vs_oracle"]})
PY
$RUN --master-port 29512 tools/overlay_multi.py > gpurun_out/overlay_${TAG}_${N}gpu.json 2> gpurun_out/overlay_${TAG}_${N}gpu.err
cat gpurun_out/overlay_${TAG}_${N}gpu.json
$RUN --master-port 29513 tools/sweep_multi.py "$@" > gpurun_out/sweep_${TAG}_${N}gpu.jsonl 2> gpurun_out/sweep_${TAG}_${N}gpu.err
tail -n 3 gpurun_out/sweep_${TAG}_${N}gpu.err
cat gpurun_out/sweep_${TAG}_${N}gpu.jsonl | cut -c1-330
