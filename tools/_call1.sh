timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests10.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gpu_tests10.log
python tools/disc_bench.py C4 3 0 | tail -1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-int8-pass > gpurun_out/r02_bench_C4_e.json 2> gpurun_out/r02_bench_C4_e.err; python tools/show_bench.py gpurun_out/r02_bench_C4_e.json | head -2
python bench.py --target pearson --steps 5 --warmup 3 > gpurun_out/r02_bench_C5_pearson_e.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02_bench_C5_pearson_e.json | head -2
