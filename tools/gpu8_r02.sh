#!/bin/bash
# Round-2 run on 8 B200s: multi-GPU parity, the big-lattice norms, the headline bench with its extras, all-reduce vs reduce-scatter.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_multi_r02.log
TNE_OPTS=rs_comm=1 timeout 300 $TR --master-port 29601 tools/norm_big.py check > gpurun_out/norm_check8_r02.log 2>&1
timeout 900 $TR --master-port 29602 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench8_r02.json 2> gpurun_out/bench8_r02.err
TNE_RS_COMM=0 timeout 300 $TR --master-port 29603 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/bench8_r02_allreduce.json 2> gpurun_out/bench8_r02_allreduce.err
TNE_OPTS=rs_comm=1 timeout 300 $TR --master-port 29604 tools/norm_big.py time 8 4 1 > gpurun_out/norm8_time8_r02.log 2>&1
cat gpurun_out/pytest_multi_r02.log; grep -v "^\*\|OMP_NUM" gpurun_out/norm_check8_r02.log | tail -8; tail -3 gpurun_out/norm8_time8_r02.log
tail -c 300 gpurun_out/bench8_r02.err
