#!/bin/bash
# Runs on the GPU box: GPU suite, both bench arms, then the launch lists and the mask-scan capture.
set -u
OUT=gpurun_out/check
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1; echo "pytest rc $?"; tail -3 $OUT/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc $?"; tail -1 $OUT/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc $?"
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc $?"
C="python tools/profile_all.py 1000000 200000"
timeout 300 $C > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv $C > $OUT/launches.log 2>&1
echo "launch list rc $?"
for K in mask_scan_fp4_kernel mask_scan_fp4_multi_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o /tmp/$K $C > $OUT/$K.log 2>&1
  echo "$K rc $?"
  ncu -i /tmp/$K.ncu-rep --page raw --csv > $OUT/${K}_raw.csv 2>/dev/null
done
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $B > $OUT/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv $B > $OUT/launches_bench.log 2>&1
echo "bench launch list rc $?"
