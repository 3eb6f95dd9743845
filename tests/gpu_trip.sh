#!/bin/bash
# round-end evidence run: parity tests, smoke, bench, launch list, full ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2800 -c 1000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 2 -c 1 -o gpurun_out/prof_gemm2 -f python tests/ncu_target.py gemm2 > gpurun_out/ncu_gemm2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 5 -o gpurun_out/prof_attn -f python tests/ncu_target.py attn > gpurun_out/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_fwd -s 1 -c 1 -o gpurun_out/prof_vq -f python tests/ncu_target.py vq > gpurun_out/ncu_vq.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:ln_ -s 2 -c 2 -o gpurun_out/prof_ln -f python tests/ncu_target.py ln > gpurun_out/ncu_ln.log 2>&1
ls -la gpurun_out | head -30
