#!/bin/bash
# NLHE session: parity tests (per-test timeout), timing probe, per-kernel launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nlhe_gpu.py -x -q --timeout 120 > gpurun_out/pytest_nlhe.log 2>&1
tail -15 gpurun_out/pytest_nlhe.log
timeout 300 python tools/nlhe_probe.py > gpurun_out/nlhe_probe.log 2>&1
cat gpurun_out/nlhe_probe.log
