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
ncu --set full --clock-control none --import-source on -k regex:mask_scan_fp4_kernel -s 1 -c 1 -o /tmp/m4 $C > $OUT/m4.log 2>&1
echo "m4 rc $?"
ncu -i /tmp/m4.ncu-rep --page raw --csv > $OUT/mask_scan_fp4_raw.csv 2>/dev/null
ncu -i /tmp/m4.ncu-rep --page source --csv > $OUT/mask_scan_fp4_source.csv 2>/dev/null
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $B > $OUT/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv $B > $OUT/launches_bench.log 2>&1
echo "bench launch list rc $?"
