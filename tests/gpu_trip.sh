#!/bin/bash
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread|Core" | head -6
python - <<'PY'
import time, torch, sys, os
sys.path.insert(0, os.getcwd())
from bench import oracle_step
for thr in (8, 16, 32, 64):
    step = oracle_step("base", 2, thr)
    t0 = time.time(); s = step(); 
    print(f"threads={thr}: base B=2 fwd+bwd {s:.1f}s -> {2/s:.3f} img/s", flush=True)
PY
