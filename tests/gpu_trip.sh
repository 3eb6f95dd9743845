#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/gpu_probe.py rowwise 2>&1 | tail -14
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_ -c 6 -o gpurun_out/prof_ln -f python tests/ncu_target.py ln > gpurun_out/ncu_ln.log 2>&1; tail -2 gpurun_out/ncu_ln.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
