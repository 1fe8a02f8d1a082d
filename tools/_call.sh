python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-int8-pass > gpurun_out/r02_bench_C4_b.json 2> gpurun_out/r02_bench_C4_b.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_C4_b.json')); print(d['ms_per_step'], json.dumps(d['e2e'], indent=1))"
