# round-end sanity trip on one B200: the whole -m gpu suite, smoke(), and a bench line with the cheap extras
O=gpurun_out/fulltests
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x --durations=10 ) > $O/pytest_gpu.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --extras vq,secondary > $O/bench.json 2> $O/bench.err
tail -6 $O/pytest_gpu.log; tail -2 $O/smoke.log; cut -c1-300 $O/bench.json; tail -n 3 $O/bench.err
