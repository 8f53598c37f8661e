#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines, ncu launch list and one full capture of the top kernel.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench cfg4 train"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | tee $OUT/bench_cfg4_train.json
echo "== bench cfg4 eval"; timeout 600 python bench.py --steps 30 --warmup 5 --mode eval 2>&1 | tail -1 | tee $OUT/bench_cfg4_eval.json
echo "== bench cfg4 train B=1024"; timeout 600 python bench.py --steps 10 --warmup 3 --batch 1024 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg4_train_b1024.json
echo "== bench cfg2 train"; timeout 600 python bench.py --steps 20 --warmup 5 --config cfg2 2>&1 | tail -1 | tee $OUT/bench_cfg2_train.json
echo "== bench cfg3 eval"; timeout 600 python bench.py --steps 5 --warmup 3 --config cfg3 --mode eval --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg3_eval.json
echo "== bench cfg1 train"; timeout 600 python bench.py --steps 30 --warmup 5 --config cfg1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg1_train.json
echo "== bench cfg5 train"; timeout 600 python bench.py --steps 3 --warmup 3 --config cfg5 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg5_train.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "== ncu full: umnn bwd / fwd"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umnn_bwd -s 3 -c 1 -o $OUT/prof_umnn_bwd -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umnn_fwd -s 3 -c 1 -o $OUT/prof_umnn_fwd -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_fwd.log 2>&1
ls -la $OUT
