#!/bin/bash
# NLHE session: parity tests (per-test timeout), timing probe for both expansion-kernel builds
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nlhe_gpu.py -x -q --timeout 120 > gpurun_out/pytest_nlhe.log 2>&1
tail -5 gpurun_out/pytest_nlhe.log
echo "compact (unroll 1)"; timeout 300 python tools/nlhe_probe.py 16384 65536 2>&1 | tee gpurun_out/nlhe_probe.log
echo "unroll 10"; RBP_NLHE_UNROLL=10 timeout 300 python tools/nlhe_probe.py 16384 2>&1 | tee -a gpurun_out/nlhe_probe.log
