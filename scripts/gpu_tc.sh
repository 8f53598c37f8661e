#!/bin/bash
TAG=${1:-r01zd}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== tc tests"; timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee $OUT/pytest_tc.txt
echo "== direct probe"; timeout 300 python scripts/umnn_direct.py 100 63 1.0 2>&1 | tail -16 | cut -c1-120 | tee $OUT/direct.txt
for a in "138600 160 160 fwd 3" "138600 160 160 dgrad 3"; do timeout 120 python scripts/gemm_trace.py $a; done 2>&1 | cut -c1-400 | tee $OUT/trace.txt
echo "== gemm bench"; timeout 600 python scripts/gemm_bench.py 2>&1 | tail -40 | tee $OUT/gemm_bench.txt
echo "== bench cfg4 train (auto)"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eval 2>&1 | tail -1 | tee $OUT/bench_cfg4.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval > $OUT/ncu_launches.log 2>&1
