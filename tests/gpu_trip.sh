#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/gpu_probe.py precision large base_rq4 2>&1 | tail -4
timeout 600 python bench.py --config large --batch 32 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_large.json | cut -c1-330
timeout 600 python bench.py --config base_rq4 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_rq4.json | cut -c1-330
timeout 600 python bench.py --config small --batch 128 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_small.json | cut -c1-330
