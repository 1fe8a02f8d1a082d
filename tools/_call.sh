NG=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/r02_bench_C4_n8_driver_style.json 2> gpurun_out/r02_bench_C4_n8_driver_style.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r02_bench_C4_n8_driver_style.json
timeout 300 $TR bench.py --impl reference --gpus $NG --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_n8_torchrun.json 2>/dev/null; head -c 900 gpurun_out/r02_bench_reference_n8_torchrun.json | tail -c 500
timeout 600 $TR bench.py --gpus $NG --target pearson --steps 5 --warmup 3 > gpurun_out/r02_bench_C5_pearson_n8_v3.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02_bench_C5_pearson_n8_v3.json | head -1
