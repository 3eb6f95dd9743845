mkdir -p gpurun_out/fulltests
( time timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x --durations=15 ) > gpurun_out/fulltests/pytest_gpu.log 2>&1
tail -30 gpurun_out/fulltests/pytest_gpu.log
