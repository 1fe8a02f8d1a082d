export GV_COOPERATIVE=0
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-int8-pass"
$CMD > gpurun_out/r02_prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(gram_tiles_kernel|k_.*)$" -c 40 --csv --log-file gpurun_out/r02_ncu_launches_c4_v2.csv $CMD > gpurun_out/r02_prof_ncu1.log 2>&1; echo "launch list rc=$?"
$CMD > gpurun_out/r02_prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:MiEpi -c 1 -o gpurun_out/r02_mi_c4_fp4_v2 $CMD > gpurun_out/r02_prof_ncu2.log 2>&1; echo "mi full rc=$?"
$CMD > gpurun_out/r02_prof_plain.log 2>&1 && ncu --set full --clock-control none -k "regex:k_transpose_counts|k_rows_to_bins|k_gene_luts" -c 3 -o gpurun_out/r02_disc_c4_v3 $CMD > gpurun_out/r02_prof_ncu3.log 2>&1; echo "disc rc=$?"
ls -la gpurun_out/r02*_v2.ncu-rep gpurun_out/r02_disc_c4_v3.ncu-rep
