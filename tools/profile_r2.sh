#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE B200): launch list of the default workload (shortened
# sequence) and one `ncu --set full` capture per hot kernel.  Every ncu run is preceded by the same
# command without ncu.  Outputs: gpurun_out/r2_*.  Summaries are made from them by tools/summarize_profiles.py.
set -u
OUT=gpurun_out
CMD="python bench.py --frames 640 --steps 1 --warmup 1 --no-aux --no-cpu-baseline"
$CMD > $OUT/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:^(void )?k_' -c 800 \
    --csv --log-file $OUT/r2_launches.csv $CMD > $OUT/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_integrate<\(bool\)0, \(bool\)1, \(int\)128, \(int\)8, \(int\)2' -s 7 -c 2 -f -o $OUT/r2_integrate_5mm $CMD > $OUT/r2_ncu_integ.log 2>&1
echo "integrate rc=$?"
$CMD > $OUT/r2_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:^(void )?(k_touch|k_depth_prepare|k_list_slots|k_sort_slots)' -s 8 -c 6 -f -o $OUT/r2_others_5mm $CMD > $OUT/r2_ncu_others.log 2>&1
echo "others rc=$?"
CMDC="python bench.py --workload quest300_rgb_v10mm --steps 1 --warmup 1 --no-aux --no-cpu-baseline"
$CMDC > $OUT/r2_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_integrate<\(bool\)1, \(bool\)1, \(int\)128, \(int\)7, \(int\)2' -s 3 -c 1 -f -o $OUT/r2_integrate_rgb $CMDC > $OUT/r2_ncu_integ_rgb.log 2>&1
echo "integrate rgb rc=$?"
$CMDC > $OUT/r2_plain5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:^(void )?k_' -c 400 \
    --csv --log-file $OUT/r2_launches_rgb.csv $CMDC > $OUT/r2_ncu_launches_rgb.log 2>&1
echo "launch list rgb rc=$?"
$CMD > $OUT/r2_plain6.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:^(void )?(k_mc_|k_scan_)' -s 6 -c 6 -f -o $OUT/r2_mc_5mm $CMD > $OUT/r2_ncu_mc.log 2>&1
echo "mc rc=$?"
