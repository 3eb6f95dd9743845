#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tests/gpu_probe.py attention 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
