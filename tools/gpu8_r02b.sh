#!/bin/bash
# Round-2 final 8-GPU run: headline bench with extras (tape retention, grouped collectives, all-gather overlap), and the overlap switched off for comparison.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench8_r02_final.json 2> gpurun_out/bench8_r02_final.err
TNE_COMM_OVERLAP=0 timeout 300 $TR --master-port 29612 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/bench8_r02_nooverlap.json 2> gpurun_out/bench8_r02_nooverlap.err
tail -c 400 gpurun_out/bench8_r02_final.err
python - <<'PY'
import json
for f in ("bench8_r02_final.json", "bench8_r02_nooverlap.json"):
    d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["recompute_flops_per_step_per_gpu"], sum(v["ms"] for v in d["kernels"].values()))
PY
