#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nlhe_gpu.py -x -q --timeout 300 --durations=5 > gpurun_out/pytest_nlhe.log 2>&1
tail -12 gpurun_out/pytest_nlhe.log
timeout 300 python tools/nlhe_probe.py 16384 2>&1 | tee gpurun_out/nlhe_probe.log
