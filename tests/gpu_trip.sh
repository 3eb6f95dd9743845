#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 60 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "(attention and (16 or 24 or 130)) or (gemm_nt and 64) or vq_matches or gemm_epilogues" > gpurun_out/sanitizer_racecheck_full.txt 2>&1
grep -E "Race reported|Error:|Warning:|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck_full.txt | sed 's/=========//' | cut -c1-260 | sort | uniq -c | sort -rn | head -30
