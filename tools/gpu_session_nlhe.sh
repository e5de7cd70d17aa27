#!/bin/bash
# NLHE session: parity tests (per-test timeout), timing probe, per-kernel launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nlhe_gpu.py -x -q --timeout 120 > gpurun_out/pytest_nlhe.log 2>&1
tail -15 gpurun_out/pytest_nlhe.log
timeout 300 python tools/nlhe_probe.py > gpurun_out/nlhe_probe.log 2>&1
cat gpurun_out/nlhe_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nlhe_launches.csv python tools/nlhe_probe.py 16384 > gpurun_out/nlhe_ncu.log 2>&1
python - <<'P'
import csv, collections
rows = list(csv.reader(open("gpurun_out/nlhe_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
k, v = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) > v:
        name = r[k].split("(")[0][:60]
        agg[name][0] += 1
        agg[name][1] += float(r[v].replace(",", "")) / 1e3
for name, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print(f"{name:60s} launches {n:6d}  total {us / 1e3:9.3f} ms")
P
