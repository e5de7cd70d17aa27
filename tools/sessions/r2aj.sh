#!/bin/bash
# r2aj: configs[1] (Leduc, reference ordered fold) through the round-2 bench on the final code
O=gpurun_out
timeout 100 python bench.py --workload leduc --steps 20 --warmup 5 --skip-cpu-baseline > $O/bench_r2aj_leduc_ordered_n1.json 2> $O/bench_r2aj_leduc.err; tail -1 $O/bench_r2aj_leduc.err; cut -c1-400 $O/bench_r2aj_leduc_ordered_n1.json
