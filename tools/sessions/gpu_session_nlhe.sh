#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nlhe_gpu.py tests/test_mccfr_gpu.py -x -q --timeout 300 > gpurun_out/pytest_nlhe.log 2>&1
tail -4 gpurun_out/pytest_nlhe.log
timeout 300 python tools/nlhe_probe.py 1024 16384 65536 2>&1 | tee gpurun_out/nlhe_probe.log
