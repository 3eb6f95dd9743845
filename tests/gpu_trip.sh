#!/bin/bash
# one GPU-box visit: parity tests, smoke, probes, bench
mkdir -p gpurun_out
timeout 300 python tests/gpu_probe.py gemm_basic gemm_cg2 gemm_major 2>&1 | grep -vE "^=====.*exit 0" | head -80
B200VQ_ATTN_FWD=mma timeout 200 python tests/gpu_probe.py attention 2>&1 | tail -8
timeout 200 python tests/gpu_probe.py attention 2>&1 | tail -9
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 200 python tests/gpu_probe.py gemm_perf 2>&1 | grep -E "bn=256|cuBLAS|dgrad|wgrad"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
