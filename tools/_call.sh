NG=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_check.py 2600 4000 > gpurun_out/dist_check_world$NG.log 2>&1; echo "dist_check rc=$?"; grep -c "identical to single rank: True" gpurun_out/dist_check_world$NG.log; grep "False\|DIST CHECK" gpurun_out/dist_check_world$NG.log | head -5; grep -i "error\|Traceback" gpurun_out/dist_check_world$NG.log | head -5
timeout 600 $TR bench.py --gpus $NG --target pearson --steps 5 --warmup 3 > gpurun_out/r02_bench_C5_pearson_n2.json 2>gpurun_out/r02_bench_C5_pearson_n2.err; echo rc=$?; python tools/show_bench.py gpurun_out/r02_bench_C5_pearson_n2.json | head -2
timeout 600 $TR bench.py --gpus $NG --target cosine --steps 5 --warmup 3 > gpurun_out/r02_bench_C5_cosine_n2.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02_bench_C5_cosine_n2.json | head -1
