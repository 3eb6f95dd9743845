#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/gpu_probe.py attention 2>&1 | grep -v "^=====" | tail -8
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
