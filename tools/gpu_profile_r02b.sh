#!/bin/bash
# Round-2 evidence, part 2 (one B200): ncu --set full of the dominant absorb kernel inside a step and of the 16-warp GEMM tile.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gett_absorb --launch-skip 3000 -c 3 -o gpurun_out/prof_absorb_r02 -f \
    python tools/prof_step.py 6 4 1 > gpurun_out/ncu_r02_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gett_gemm --launch-skip 100 -c 2 -o gpurun_out/prof_gemm16_r02 -f \
    python tools/gemm_time.py > gpurun_out/ncu_r02_e.log 2>&1
ls -la gpurun_out/prof_absorb_r02.ncu-rep gpurun_out/prof_gemm16_r02.ncu-rep
