#!/bin/bash
# r2o: first run of the tcgen05 / TMA screen — error against the exact solver, survivors, speed
O=gpurun_out
TAG=${1:-r2o}
timeout 60 python tests/measure/screen_probe.py --n 2000 --k 200 --alpha 0.02 > $O/${TAG}_screen_a002.json 2> $O/${TAG}_screen.err; tail -5 $O/${TAG}_screen.err; cat $O/${TAG}_screen_a002.json
timeout 60 python tests/measure/screen_probe.py --n 2000 --k 200 --alpha 0.3 > $O/${TAG}_screen_a03.json 2>> $O/${TAG}_screen.err; tail -5 $O/${TAG}_screen.err; cat $O/${TAG}_screen_a03.json
