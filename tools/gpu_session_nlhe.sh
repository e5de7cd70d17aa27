#!/bin/bash
# NLHE bring-up session: parity tests, then a timing probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nlhe_gpu.py -x -q > gpurun_out/pytest_nlhe.log 2>&1
tail -30 gpurun_out/pytest_nlhe.log
timeout 600 python - > gpurun_out/nlhe_probe.log 2>&1 <<'P'
import time
from robopoker_b200.nlhe import Nlhe
for batch in (1024, 16384, 65536):
    g = Nlhe(batch=batch, seed=1, table_slots=1 << 24)
    g.step(3)
    c0 = g.counters()
    ms = g.step_timed(5, flush_l2=True)
    c1 = g.counters()
    up = c1["updates"] - c0["updates"]
    print(batch, [round(x / 5, 3) for x in ms], "updates/epoch", up // 5, "updates/s %.3e" % (up / (ms[0] * 1e-3)), c1, flush=True)
P
cat gpurun_out/nlhe_probe.log
