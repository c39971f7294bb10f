#!/bin/bash
# One proof split over N GPUs + the point-split MSM at fixed total sizes (BASELINE config #3 / #4) on an N-GPU box.
# usage: bash tools/multi_gpu_sweep.sh N [tag]     (run under `gpurun --gpus N`); writes gpurun_out/<tag>_*_<N>gpu.json
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/${TAG}_dist_tests_${N}gpu.log 2>&1; tail -2 gpurun_out/${TAG}_dist_tests_${N}gpu.log)
$TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_prove_${N}gpu.json 2> gpurun_out/${TAG}_prove_${N}gpu.err
for L in 22 24 26; do
  $TR --master-port $((29530 + L)) bench.py --gpus $N --workload msm --total-log-n $L --steps 5 --warmup 3 > gpurun_out/${TAG}_msm_total${L}_${N}gpu.json 2> gpurun_out/${TAG}_msm_total${L}_${N}gpu.err
done
$TR --master-port 29560 bench.py --gpus $N --workload msm --log-n 18 --steps 20 --warmup 3 > gpurun_out/${TAG}_msm_weak18_${N}gpu.json 2> gpurun_out/${TAG}_msm_weak18_${N}gpu.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*_${N}gpu.json")):
    try:
        d = json.load(open(f)); print(f, d["metric"], round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 3), d.get("scaling"))
    except Exception as e:
        print(f, "unreadable", e)
PY
