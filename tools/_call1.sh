timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests12.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gpu_tests12.log
python tools/disc_bench.py C4 3 0 | tail -2
python tools/disc_bench.py C5 3 0 | tail -1
python tools/disc_bench.py C2 3 0 | tail -1
python bench.py --target pearson --steps 5 --warmup 3 > gpurun_out/r02_bench_C5_pearson_f.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02_bench_C5_pearson_f.json | head -2
