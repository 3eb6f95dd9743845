#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tests/gpu_probe.py attention 2>&1 | tail -9
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python tests/gpu_probe.py rowwise 2>&1 | grep -E "colsum|ln "
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
timeout 600 python tests/gpu_probe.py precision 2>&1 | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 5 -o gpurun_out/prof_attn -f python tests/ncu_target.py attn > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
