#!/bin/bash
TAG=${1:-r01z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== rw tests"; timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -p no:cacheprovider -k "rw or layerwise" 2>&1 | tail -15 | tee $OUT/pytest_rw.txt
echo "== gemm bench"; timeout 300 python scripts/gemm_bench.py 2>&1 | tail -8 | tee $OUT/gemm_bench.txt
echo "== direct accuracy"; timeout 300 python scripts/umnn_direct.py 100 63 1.0 2>&1 | tail -16 | cut -c1-140 | tee $OUT/direct.txt
echo "== bench default"; timeout 600 python bench.py --no-cpu-baseline --no-eval 2>&1 | tail -1 | tee $OUT/bench_default.json
echo "== bench B=1024"; timeout 600 python bench.py --batch 1024 --steps 10 --no-cpu-baseline --no-eval 2>&1 | tail -1 | tee $OUT/bench_b1024.json
