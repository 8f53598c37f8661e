#!/bin/bash
# One gpurun call: the evidence set of the round (GPU tests, smoke, bench lines of every config + the reference arm, randomised S, large
# batch, K4 roofline, ncu launch list of the default bench command, ncu --set full captures of the hot kernels for profiles/ncu_traffic.json).
OUT=gpurun_out/${1:-r02x}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench default"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench_default.json | cut -c1-200
echo "== bench reference arm"; timeout 900 python bench.py --impl reference 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-300
for cfg in cfg1 cfg2 cfg3 cfg5; do
  echo "== bench $cfg"; timeout 900 python bench.py --config $cfg --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_$cfg.json | cut -c1-160
done
echo "== bench cfg4 random S"; timeout 600 python bench.py --random-steps --no-cpu-baseline --no-eval 2>&1 | tail -1 | tee $OUT/bench_cfg4_randomS.json | cut -c1-160
echo "== bench cfg4 B=1024"; timeout 600 python bench.py --batch 1024 --steps 10 --no-cpu-baseline --no-eval 2>&1 | tail -1 | tee $OUT/bench_cfg4_b1024.json | cut -c1-160
echo "== k4 roofline"; timeout 300 python scripts/k4_roofline.py 2>&1 | tail -1 | tee $OUT/k4_roofline.json | cut -c1-300
echo "== ncu launch list cfg4 (eager)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_cfg4_train_eager.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eval --cuda-graph off > $OUT/ncu_launches.log 2>&1; tail -1 $OUT/ncu_launches.log | cut -c1-200
echo "== ncu full cfg4"; timeout 900 ncu --set full --clock-control none -k regex:"tc_gemm|tc_wgrad2|rw_wgrad_kernel|umnn_fwd_tc3|umnn_bwd_tc3|dag_l1" -s 52 -c 13 -o $OUT/prof_cfg4 python bench.py --steps 3 --warmup 3 --no-eval --no-cpu-baseline --cuda-graph off > $OUT/ncu_cfg4.log 2>&1; tail -2 $OUT/ncu_cfg4.log | cut -c1-200
ncu -i $OUT/prof_cfg4.ncu-rep --page raw --csv > $OUT/prof_cfg4_raw.csv 2>/dev/null; rm -f $OUT/prof_cfg4.ncu-rep   # gpurun_out travels back only below 64 MiB
echo "== ncu full cfg5"; timeout 900 ncu --set full --clock-control none -k regex:"tc_gemm|tc_wgrad2|dag_embed" -s 44 -c 11 -o $OUT/prof_cfg5 python bench.py --config cfg5 --steps 3 --warmup 3 --no-eval --no-cpu-baseline --cuda-graph off > $OUT/ncu_cfg5.log 2>&1; tail -2 $OUT/ncu_cfg5.log | cut -c1-200
ncu -i $OUT/prof_cfg5.ncu-rep --page raw --csv > $OUT/prof_cfg5_raw.csv 2>/dev/null; rm -f $OUT/prof_cfg5.ncu-rep
ls -la $OUT
