#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tests/gpu_probe.py attention 2>&1 | tail -9
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 2 -c 1 -o gpurun_out/prof_gemm2 -f python tests/ncu_target.py gemm2 > gpurun_out/ncu_gemm2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 3 -c 4 -o gpurun_out/prof_attn -f python tests/ncu_target.py attn > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
