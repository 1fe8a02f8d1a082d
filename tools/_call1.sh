timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r02_bench_default_final.json
python bench.py --impl reference > gpurun_out/r02_bench_reference_final.json 2>/dev/null; head -c 700 gpurun_out/r02_bench_reference_final.json
