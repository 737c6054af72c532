#!/bin/bash
# One GPU-box visit. Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh <tag> [steps...]
# steps: smoke tests bench ref launches ncu sweep sweeprandom   (default: smoke tests bench ref launches ncu)
TAG=${1:-r01}; shift
STEPS=${@:-smoke tests bench ref launches ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
for s in $STEPS; do
case $s in
smoke) echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -5;;
tests) echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tee $OUT/pytest_gpu.log | tail -15;;
bench) echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-1200; tail -5 $OUT/bench.err;;
ref) echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_reference.json | cut -c1-400;;
launches) echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 > $OUT/bench_under_ncu.log 2>&1
  grep -c . $OUT/launches.csv;;
ncu) echo "== ncu full capture"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tracePersistent -s 4 -c 1 -o $OUT/prof_trace -f python bench.py --steps 3 --warmup 3 > $OUT/ncu_full.log 2>&1; tail -2 $OUT/ncu_full.log | cut -c1-300;;
baketests) echo "== bake tests"; timeout 900 python -m pytest tests/test_gpu_bake.py -q --timeout 600 2>&1 | tee $OUT/pytest_bake.log | tail -25;;
bakesan) echo "== bake under compute-sanitizer"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_bake.py -q --timeout 800 -k "unbaked or cycles or uniform or deterministic" > $OUT/bake_memcheck.log 2>&1; echo "exit $?"; tail -8 $OUT/bake_memcheck.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_bake.py -q --timeout 800 -k "unbaked" > $OUT/bake_racecheck.log 2>&1; echo "exit $?"; tail -5 $OUT/bake_racecheck.log;;
bake) echo "== bake bench"
  timeout 600 python scripts/bake_bench.py --scene terrain --log2 12 --edits 200 --out $OUT/bake.jsonl 2>&1 | tail -3
  timeout 600 python scripts/bake_bench.py --scene city --log2 16 --edits 200 --out $OUT/bake.jsonl 2>&1 | tail -3
  timeout 600 python scripts/bake_bench.py --scene soup --log2 14 --edits 50 --out $OUT/bake.jsonl 2>&1 | tail -3;;
build) echo "== build bench"; timeout 900 python scripts/build_bench.py --out $OUT/build.jsonl 2>&1 | tail -8;;
bakelaunch) echo "== ncu launch list of the bake kernels (time + DRAM bytes)"
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"bake|scan|subdag|dense" --csv --log-file $OUT/bake_launches.csv python scripts/bake_bench.py --scene soup --log2 14 --edits 50 --reps 0 --no-reference > $OUT/ncu_bake.log 2>&1; grep -c . $OUT/bake_launches.csv;;
edit) echo "== edit bench"
  timeout 600 python scripts/edit_bench.py --scene city --log2 16 --out $OUT/edit.jsonl 2>&1 | tail -3
  timeout 600 python scripts/edit_bench.py --scene terrain --log2 12 --out $OUT/edit.jsonl 2>&1 | tail -3
  timeout 900 python bench.py --workload pathtrace --scene city --scene-log2 16 --width 3840 --height 2160 --spp 16 --edits --device-edits --steps 3 --warmup 3 2> $OUT/config5_device_edits.err | tee $OUT/config5_device_edits.json | cut -c1-900;;
sweep) echo "== sweep"; timeout 900 python scripts/sweep.py primary 2>&1 | tee $OUT/sweep_primary.jsonl;;
sweeprandom) echo "== sweep random"; timeout 900 python scripts/sweep.py random 2>&1 | tee $OUT/sweep_random.jsonl;;
esac
done
ls $OUT
