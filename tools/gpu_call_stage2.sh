# development trip: stage-2 tests, the attention / GEMM regression subset, stage-2 timings, a short bench A/B
O=gpurun_out/call2
mkdir -p $O
timeout 600 python -m pytest tests/test_stage2.py -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | tail -60 > $O/stage2_tests.log
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_kernels_r2.py -m gpu -q --no-header -p no:cacheprovider -k "attention or gemm" 2>&1 | tail -15 > $O/kernels_regress.log
timeout 300 python tools/stage2_probe.py 1024 8 8 > $O/stage2_probe_B8.json 2> $O/stage2_probe.err
timeout 300 python tools/stage2_probe.py 1024 8 32 > $O/stage2_probe_B32.json 2>> $O/stage2_probe.err
timeout 400 python bench.py --steps 5 --warmup 3 --extras "" > $O/bench_short.json 2> $O/bench_short.err
tail -4 $O/stage2_tests.log; tail -2 $O/kernels_regress.log; cat $O/stage2_probe_B8.json $O/stage2_probe_B32.json; cut -c1-400 $O/bench_short.json; tail -3 $O/stage2_probe.err $O/bench_short.err
