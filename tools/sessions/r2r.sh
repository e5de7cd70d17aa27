#!/bin/bash
# r2r: bench-size parity tests (NLHE 16384 x 2, 65536 x 1; Leduc batched 262144, batch-1 lock-step), contract-vs-libm flip rate at 20000 x 200
O=gpurun_out
TAG=${1:-r2r}
timeout 900 python -m pytest tests/test_nlhe_gpu.py tests/test_mccfr_gpu.py -x -q -m gpu --timeout 600 -k "bench_sizes or old_bench_size or lock_step" 2>&1 | tail -4
timeout 600 python tests/measure/libm_flip_rate_gpu.py --n 20000 --k 200 > $O/${TAG}_contract_vs_libm_flips_n20000_k200.json 2> $O/${TAG}_flips.err; tail -2 $O/${TAG}_flips.err; cat $O/${TAG}_contract_vs_libm_flips_n20000_k200.json
