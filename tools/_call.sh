timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests8.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02_gpu_tests8.log
