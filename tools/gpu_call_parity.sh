# the BASELINE-config parity table with its printed error figures
mkdir -p gpurun_out/parity
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --no-header -p no:cacheprovider -s -k "baseline_config_parity" 2>&1 | grep "parity\]\|passed\|failed\|Error" > gpurun_out/parity/parity_table.txt
cat gpurun_out/parity/parity_table.txt
