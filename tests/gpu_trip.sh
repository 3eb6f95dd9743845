#!/bin/bash
# round-end evidence (outputs must stay < 64 MiB to be copied back)
mkdir -p gpurun_out
timeout 240 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err > gpurun_out/bench.json; cut -c1-120 gpurun_out/bench.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 2800 -c 1000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 100 ncu --set full --clock-control none -k regex:gemm_tf32 -c 3 -o gpurun_out/prof_gemm2 -f python tests/ncu_target.py gemm2 > gpurun_out/ncu_gemm2.log 2>&1
timeout 120 ncu --set full --clock-control none -k regex:attn_ -c 6 -o gpurun_out/prof_attn -f python tests/ncu_target.py attn > gpurun_out/ncu_attn.log 2>&1
timeout 100 ncu --set full --clock-control none -k regex:ln_ -c 6 -o gpurun_out/prof_ln -f python tests/ncu_target.py ln > gpurun_out/ncu_ln.log 2>&1
timeout 100 ncu --set full --clock-control none -k regex:vq_fwd -c 2 -o gpurun_out/prof_vq -f python tests/ncu_target.py vq > gpurun_out/ncu_vq.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | awk '{print $5, $9}' | tail -12
