mkdir -p gpurun_out/call1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/call1/gpu.txt
timeout 600 python -m pytest tests/test_stage2.py -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | tail -120 > gpurun_out/call1/stage2_tests.log
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_kernels_r2.py -m gpu -q --no-header -p no:cacheprovider -k "attention" 2>&1 | tail -30 > gpurun_out/call1/attention_regress.log
timeout 200 python -m pytest tests/test_gpu_model.py -m gpu -q --no-header -p no:cacheprovider -k "post_quant_with_positional or reference_golden" 2>&1 | tail -30 > gpurun_out/call1/model_fused.log
timeout 300 python tools/stage2_probe.py 1024 8 8 > gpurun_out/call1/stage2_probe.json 2> gpurun_out/call1/stage2_probe.err
tail -5 gpurun_out/call1/stage2_tests.log; tail -3 gpurun_out/call1/model_fused.log; tail -3 gpurun_out/call1/attention_regress.log; cat gpurun_out/call1/stage2_probe.json; tail -3 gpurun_out/call1/stage2_probe.err
