timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests6.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gpu_tests6.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-int8-pass > gpurun_out/r02_bench_C4_c.json 2> gpurun_out/r02_bench_C4_c.err; echo rc=$?
python tools/show_bench.py gpurun_out/r02_bench_C4_c.json
GV_PACK_UPLOADS=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-int8-pass > gpurun_out/r02_bench_C4_c_nopack.json 2> gpurun_out/r02_bench_C4_c_nopack.err; echo rc=$?
python tools/show_bench.py gpurun_out/r02_bench_C4_c_nopack.json | tail -1
python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline --no-int8-pass > gpurun_out/r02_bench_C2_c.json 2> gpurun_out/r02_bench_C2_c.err; python tools/show_bench.py gpurun_out/r02_bench_C2_c.json
