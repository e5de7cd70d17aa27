#!/bin/bash
# Leduc exploitability curve (GPU vs oracle in lockstep) + full abstraction pipeline with the rewritten Sinkhorn layer.
O=gpurun_out
TAG=${1:-r1p}
mkdir -p $O
timeout 900 python tests/measure/leduc_curve.py --trees 1048576 --batch 1 1024 16384 > $O/${TAG}_leduc_curve.jsonl 2> $O/${TAG}_leduc_curve.err
tail -4 $O/${TAG}_leduc_curve.jsonl | cut -c1-400
timeout 600 python tests/measure/leduc_curve.py --trees 268435456 --batch 262144 --fold batched --no-oracle > $O/${TAG}_leduc_curve_batched.jsonl 2>> $O/${TAG}_leduc_curve.err
tail -2 $O/${TAG}_leduc_curve_batched.jsonl | cut -c1-300
timeout 900 python tools/abstraction_pipeline.py --blueprint-epochs 20 > $O/${TAG}_pipeline_blueprint.json 2> $O/${TAG}_pipeline.err
cat $O/${TAG}_pipeline_blueprint.json | cut -c1-250
tail -3 $O/${TAG}_pipeline.err $O/${TAG}_leduc_curve.err
