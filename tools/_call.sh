nvidia-smi topo -m 2>/dev/null | head -6
timeout 600 python -m pytest tests/test_gpu_runtime.py -m gpu -x -q > gpurun_out/r02_gpu_tests3.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_gpu_tests3.log
for f in broadcast sharded; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --front $f --no-int8-pass > gpurun_out/r02_bench_C4_n2_$f.json 2> gpurun_out/r02_bench_C4_n2_$f.err; echo "bench $f rc=$?"
tail -3 gpurun_out/r02_bench_C4_n2_$f.err
python tools/show_bench.py gpurun_out/r02_bench_C4_n2_$f.json 2>/dev/null || head -c 1500 gpurun_out/r02_bench_C4_n2_$f.json
done
