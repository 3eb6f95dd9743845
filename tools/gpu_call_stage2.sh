# development trip: stage-2 tests and timings
O=gpurun_out/call3
mkdir -p $O
timeout 600 python -m pytest tests/test_stage2.py -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | tail -60 > $O/stage2_tests.log
timeout 300 python tools/stage2_probe.py 1024 8 32 > $O/stage2_probe_B32.json 2> $O/stage2_probe.err
tail -12 $O/stage2_tests.log; cat $O/stage2_probe_B32.json; tail -n 3 $O/stage2_probe.err
