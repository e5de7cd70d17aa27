#!/bin/bash
# r2ae: launch list of the turn-layer Elkan iteration (which kernel is the k-independent part of the step?)
O=gpurun_out
TAG=${1:-r2ae}
for K in 100 500; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_lloyd_k${K}_launches.csv python bench.py --workload lloyd_turn --k $K --points 3000000 --steps 3 --warmup 3 --skip-cpu-baseline > $O/${TAG}_k$K.log 2>&1; tail -1 $O/${TAG}_k$K.log | cut -c1-200
python - $O/${TAG}_lloyd_k${K}_launches.csv <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
t=collections.OrderedDict(); n=collections.Counter()
for r in rows[1:]:
    v=float(r[vi].replace(',','')); v = v/1e6 if r[ui]=='ns' else (v/1e3 if r[ui] in('us','usecond') else v)
    k=r[ki][:70]; t[k]=t.get(k,0)+v; n[k]+=1
for k,v in sorted(t.items(), key=lambda x:-x[1])[:12]: print(f"{v:9.3f} ms total {n[k]:4d} launches  {k}")
PY
done
