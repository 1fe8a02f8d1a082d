timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests9.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gpu_tests9.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_C4_d.json 2> gpurun_out/r02_bench_C4_d.err; echo rc=$?
python tools/show_bench.py gpurun_out/r02_bench_C4_d.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_C4_d.json')); r=d['roofline_int8']; print('int8', r['kernel_ms'], r['achieved'], r['frac'], r['frac_of_nominal'])"
python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline --no-int8-pass > gpurun_out/r02_bench_C3_d.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02_bench_C3_d.json | head -2
python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline --no-int8-pass > gpurun_out/r02_bench_C2_d.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02_bench_C2_d.json | head -2
