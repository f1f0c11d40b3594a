#!/bin/bash
# Round-2 evidence run on ONE B200: per-bucket kernel profile, launch list of the bench command, ncu --set full of the new kernels.
mkdir -p gpurun_out
TNE_FINE=1 python tools/prof_step.py 6 4 2 > gpurun_out/prof_step_fine_r02.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err || exit 1
# launch list of the same command (serialised, cold-cache: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 6000 --csv --log-file gpurun_out/launches_bench_r02.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-extras > gpurun_out/ncu_r02_list.log 2>&1
# full captures: fused reverse site kernel, fused forward site kernel, dense absorb (one launch each from the middle of the step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gett_site2_bwd --launch-skip 200 -c 2 -o gpurun_out/prof_site2b_full_r02 -f \
    python tools/prof_step.py 6 4 1 > gpurun_out/ncu_r02_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gett_gemm -c 3 -o gpurun_out/prof_gemm_r02 -f \
    python tools/gemm_time.py > gpurun_out/ncu_r02_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:permute -c 5 -o gpurun_out/prof_permute_r02 -f \
    python tools/permute_time.py > gpurun_out/ncu_r02_c.log 2>&1
ls -la gpurun_out/*r02*
