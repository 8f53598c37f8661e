#!/bin/bash
# One gpurun call: full GPU parity suite, smoke, default bench line, engine comparison (fused FFMA vs layer-wise tcgen05).
TAG=${1:-r01w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench default"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench_default.json
echo "== bench fused"; timeout 600 python bench.py --umnn-engine fused --no-cpu-baseline --no-eval 2>&1 | tail -1 | tee $OUT/bench_fused.json
echo "== bench layerwise eager"; timeout 600 python bench.py --umnn-engine layerwise --cuda-graph off --no-cpu-baseline --no-eval 2>&1 | tail -1 | tee $OUT/bench_lw_eager.json
echo "== bench B=1024 auto"; timeout 600 python bench.py --batch 1024 --steps 10 --no-cpu-baseline --no-eval 2>&1 | tail -1 | tee $OUT/bench_b1024.json
echo "== bench B=1024 fused"; timeout 600 python bench.py --batch 1024 --steps 10 --umnn-engine fused --no-cpu-baseline --no-eval 2>&1 | tail -1 | tee $OUT/bench_b1024_fused.json
echo "== direct accuracy"; timeout 300 python scripts/umnn_direct.py 100 63 1.0 2>&1 | tail -16 | cut -c1-140 | tee $OUT/direct.txt
