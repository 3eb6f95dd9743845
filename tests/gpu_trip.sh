#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/dbg_colsum.py 2>&1 | tail -3
B200VQ_GEMM_COLSUM=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_nocs.json | cut -c1-160
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-160
B200VQ_GEMM_COLSUM=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_nocs2.json | cut -c1-160
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench2.json | cut -c1-160
