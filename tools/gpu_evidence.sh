#!/bin/bash
# Round-end GPU evidence, run from the repo root on a B200 (e.g. `gpurun --timeout 900 -- 'bash tools/gpu_evidence.sh'`).
# Everything lands in gpurun_out/ (kept under 64 MiB so it is copied back); then, on the build host:
#     python tools/summarize_profiles.py        # -> profiles/r01_*
# ~9 minutes of box time: bench with CPU baseline ~70 s, launch list ~110 s, each --set full capture 40-150 s.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv,noheader
timeout 240 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err > gpurun_out/bench.json; cut -c1-160 gpurun_out/bench.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 2800 -c 1000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none -k regex:gemm_tf32 -c 3 -o gpurun_out/prof_gemm2 -f python tests/ncu_target.py gemm2 > gpurun_out/ncu_gemm2.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:attn_ -c 6 -o gpurun_out/prof_attn -f python tests/ncu_target.py attn > gpurun_out/ncu_attn.log 2>&1
timeout 150 ncu --set full --clock-control none -k regex:ln_ -c 6 -o gpurun_out/prof_ln -f python tests/ncu_target.py ln > gpurun_out/ncu_ln.log 2>&1
timeout 150 ncu --set full --clock-control none -k regex:vq_fwd -c 2 -o gpurun_out/prof_vq -f python tests/ncu_target.py vq > gpurun_out/ncu_vq.log 2>&1
du -sh gpurun_out
