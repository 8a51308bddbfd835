#!/bin/bash
# One gpurun call: parity tests, smoke, bench, per-shape timings, ncu launch list + captures.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh TAG step...'   steps: test bench ab micro ncu ncufull
# gpurun_out/ is capped at 64 MiB: .ncu-rep files are exported to CSV here and dropped when large.
set -u
TAG=${1:-r01}; shift || true
STEPS=${*:-test bench micro ncu ncufull}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
cp MEASURED_PEAKS.json $OUT/ 2>/dev/null || true
export_rep() {  # $1 = basename without extension
  ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null
  ncu -i $1.ncu-rep --page details --csv > $1.details.csv 2>/dev/null
  if [ "${2:-}" = "source" ]; then ncu -i $1.ncu-rep --page source --csv > $1.source.csv 2>/dev/null; fi
  sz=$(stat -c %s $1.ncu-rep); if [ $sz -gt 12000000 ]; then rm -f $1.ncu-rep; fi
}
for s in $STEPS; do
  case $s in
    test)
      timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
      tail -15 $OUT/pytest_gpu.log
      timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps 32 --warmup 4 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
      tail -3 $OUT/bench.err; cat $OUT/bench.json | head -c 3500; echo ;;
    cfg3)  # BASELINE configs[2]: 1024x1024 clip, T=16 growing memory bank
      timeout 600 python bench.py --size 1024 --memory 16 --steps 16 --warmup 3 --no-cpu-baseline --no-profile \
        > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; echo "cfg3 rc=$?"; tail -2 $OUT/bench_cfg3.err; cat $OUT/bench_cfg3.json ;;
    ab)   # scheduling A/B: side-stream memorize and PDL on/off (device-resident value only)
      for v in "0 0" "1 0" "0 1"; do set -- $v
        OTVM_OVERLAP=$1 OTVM_PDL=$2 timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-profile \
          > $OUT/bench_overlap$1_pdl$2.json 2> $OUT/bench_overlap$1_pdl$2.err
        echo "overlap=$1 pdl=$2: $(python -c "import json,sys; d=json.load(open('$OUT/bench_overlap$1_pdl$2.json')); print(d['value'], d['e2e']['value'])" 2>&1 | tail -1)"
      done ;;
    micro)
      timeout 300 python scripts/bench_read.py > $OUT/bench_read.txt 2>&1; cat $OUT/bench_read.txt
      timeout 300 python scripts/bench_conv.py > $OUT/bench_conv.txt 2>&1
      OTVM_OVERLAP=0 timeout 300 python scripts/profile_frame.py bf16 > $OUT/profile_frame.txt 2>&1; head -12 $OUT/profile_frame.txt ;;
    ncu)
      OTVM_PDL=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file $OUT/launches.csv python scripts/one_frame.py ${PREC:-bf16x2} 1 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
      OTVM_PDL=0 timeout 900 ncu --profile-from-start off --clock-control none --csv \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
        --log-file $OUT/launch_metrics.csv python scripts/one_frame.py ${PREC:-bf16x2} 1 > $OUT/ncu_metrics.log 2>&1; echo "ncu metrics rc=$?" ;;
    ncufull)
      OTVM_PDL=0 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:memory_read_tc -c 1 -f -o $OUT/read_full python scripts/one_frame.py ${PREC:-bf16x2} 1 > $OUT/ncu_read.log 2>&1; echo "ncu read rc=$?"
      export_rep $OUT/read_full source
      OTVM_PDL=0 timeout 600 ncu --set full --clock-control none --import-source on \
        -k regex:conv_tc -c 12 -f -o $OUT/conv_full python scripts/conv_one.py ${PREC:-bf16x2} > $OUT/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
      export_rep $OUT/conv_full source
      OTVM_PDL=0 timeout 600 ncu --set full --clock-control none \
        -k regex:gn_apply -c 4 -f -o $OUT/gn_full python scripts/conv_one.py ${PREC:-bf16x2} gn > $OUT/ncu_gn.log 2>&1; echo "ncu gn rc=$?"
      export_rep $OUT/gn_full
      ls -la $OUT; du -sh gpurun_out ;;
  esac
done
