timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runtime.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r02_gpu_tests13.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gpu_tests13.log
python tools/disc_bench.py C4 3 0 | tail -2
python tools/disc_bench.py C5 3 0 | tail -1
