#!/bin/bash
# One short gpurun call for the round's closing evidence (most important first): GPU parity suite, smoke, both bench arms,
# ncu launch list of the default bench command, then the other configs' bench lines.
TAG=${1:-r01zzz}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench default"; timeout 150 python bench.py 2>&1 | tail -1 | tee $OUT/bench_default.json | cut -c1-120
echo "== bench reference"; timeout 100 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-120
echo "== ncu launch list"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval --cuda-graph off > $OUT/ncu_launches.log 2>&1
echo "== bench cfg5/cfg3/cfg2/cfg1"
timeout 100 python bench.py --config cfg5 --steps 5 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg5.json | cut -c1-120
timeout 100 python bench.py --config cfg3 --steps 5 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg3.json | cut -c1-120
timeout 60 python bench.py --config cfg2 --steps 20 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg2.json | cut -c1-120
timeout 60 python bench.py --config cfg1 --steps 30 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg1.json | cut -c1-120
ls $OUT
