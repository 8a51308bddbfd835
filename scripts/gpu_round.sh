#!/bin/bash
# One gpurun call: parity tests, smoke, bench, per-shape timings, ncu launch list + full captures.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag] [steps...]'   steps: test bench micro ncu ncufull
set -u
TAG=${1:-r01}; shift || true
STEPS=${*:-test bench micro ncu ncufull}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
cp MEASURED_PEAKS.json $OUT/ 2>/dev/null || true
for s in $STEPS; do
  case $s in
    test)
      timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
      tail -5 $OUT/pytest_gpu.log
      timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps 32 --warmup 4 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
      cat $OUT/bench.json | head -c 3000; echo ;;
    micro)
      timeout 300 python scripts/bench_read.py > $OUT/bench_read.txt 2>&1; cat $OUT/bench_read.txt
      timeout 300 python scripts/bench_conv.py > $OUT/bench_conv.txt 2>&1
      timeout 300 python scripts/profile_frame.py bf16 > $OUT/profile_frame.txt 2>&1; head -40 $OUT/profile_frame.txt ;;
    ncu)
      timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file $OUT/launches.csv python scripts/one_frame.py bf16 1 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?" ;;
    ncufull)
      timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:memory_read_tc -c 1 -f -o $OUT/read_full python scripts/one_frame.py bf16 1 > $OUT/ncu_read.log 2>&1; echo "ncu read rc=$?"
      timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:conv_tc_kernel -c 180 -f -o $OUT/conv_full python scripts/one_frame.py bf16 1 > $OUT/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
      ls -la $OUT ;;
  esac
done
