#!/bin/bash
# ncu launch list of the default bench command + one full capture of the dominant kernel of the layer-wise UMNN backward.
TAG=${1:-r01x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== ncu launch list (eager, 1 timed step region)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval --cuda-graph off > $OUT/ncu_launches.log 2>&1
tail -2 $OUT/ncu_launches.log
echo "== ncu full: tc_gemm in the UMNN layer-wise path"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 40 -c 6 -o $OUT/prof_tc_gemm -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval --cuda-graph off > $OUT/ncu_gemm.log 2>&1
tail -2 $OUT/ncu_gemm.log
ls -la $OUT
