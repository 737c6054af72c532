#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench, launch list, one full ncu capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -5
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tee $OUT/pytest_gpu.log | tail -15
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-1500
tail -5 $OUT/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_reference.json | cut -c1-600
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 > $OUT/bench_under_ncu.log 2>&1
grep -c . $OUT/launches.csv
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tracePersistent -s 4 -c 2 -o $OUT/prof_trace -f python bench.py --steps 3 --warmup 3 > $OUT/ncu_full.log 2>&1
ls -la $OUT
