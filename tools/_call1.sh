export GV_COOPERATIVE=0
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-int8-pass --operand int8"
$CMD > gpurun_out/r02_prof_plain.log 2>&1 && ncu --set full --clock-control none --kernel-name-base demangled -k regex:MiEpi -c 1 -o gpurun_out/r02_mi_c4_int8_v2 $CMD > gpurun_out/r02_prof_ncu2.log 2>&1; echo "mi int8 rc=$?"
CMD="python tools/disc_bench.py C4 1 0"
$CMD > gpurun_out/r02_prof_plain.log 2>&1 && ncu --set full --clock-control none -k "regex:k_transpose_counts|k_rows_to_bins" -s 2 -c 2 -o gpurun_out/r02_disc_c4_v4 $CMD > gpurun_out/r02_prof_ncu3.log 2>&1; echo "disc rc=$?"
ls -la gpurun_out/r02_mi_c4_int8_v2.ncu-rep gpurun_out/r02_disc_c4_v4.ncu-rep
