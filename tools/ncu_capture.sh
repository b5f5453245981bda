#!/bin/bash
# Runs on the GPU box: one ncu --set full capture per product kernel, exported to CSV (raw page) so that only
# small files travel back.  Usage: tools/ncu_capture.sh <rows> <batch_rows>
set -u
C="python tools/profile_all.py ${1:-1000000} ${2:-200000}"
OUT=gpurun_out/ncu
mkdir -p $OUT
timeout 300 $C > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv $C > $OUT/launches.log 2>&1
echo "launch list rc $?"
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/$1 $C > $OUT/$1.log 2>&1
  echo "$1 rc $?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > $OUT/$1_source.csv 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
cap scan_fused 'scan_kernel' 1
cap scan_distances 'scan_kernel' 4
cap scan_search 'scan_kernel' 7
cap mask_scan_fp4 'mask_scan_fp4_kernel' 1
cap batch_distances_s8 'batch_distances_kernel' 1
cap batch_distances_u8 'batch_distances_kernel' 4
cap mask_scan_fp4_multi 'mask_scan_fp4_multi_kernel' 1
cap combine_decode 'combine_decode_kernel' 0
du -sh $OUT
# launch list of the bench command itself (the timed region is the scan_kernel launches, one per step)
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $B > $OUT/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv $B > $OUT/launches_bench.log 2>&1
echo "bench launch list rc $?"
