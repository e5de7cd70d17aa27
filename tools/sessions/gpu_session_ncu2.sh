#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r1j_nlhe_launches64k.csv python tools/nlhe_probe.py 65536 > $O/ncu_launch_r1j.log 2>&1
# classify and expand at level 10 of epoch 4 (3 launches per level, 48 levels per epoch, root kernel first)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nlhe_classify_kernel --launch-skip 154 -c 1 -o $O/r1j_nlhe_classify -f python tools/nlhe_probe.py 65536 > $O/ncu_classify_r1j.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nlhe_expand_kernel --launch-skip 154 -c 1 -o $O/r1j_nlhe_expand -f python tools/nlhe_probe.py 65536 > $O/ncu_expand_r1j.log 2>&1
python - <<'P'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r1j_nlhe_launches64k.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
k, v = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) > v:
        name = r[k].split("(")[0][:60]
        agg[name][0] += 1
        agg[name][1] += float(r[v].replace(",", "")) / 1e3
for name, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
    print(f"{name:60s} launches {n:6d}  per-epoch {us / 1e3 / 8:8.3f} ms")
P
