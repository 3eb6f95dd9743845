#!/bin/bash
B200VQ_LIB=$PWD/enhancing-transformers_b200/libb200vq_trace.so timeout 200 python tests/gpu_probe.py attn_trace 2>&1 | tail -58
