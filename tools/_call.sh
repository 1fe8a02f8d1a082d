NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gpu_tests5.log
grep -c "identical to single rank: True" gpurun_out/dist_check_world$NG.log; grep "False\|DIST CHECK" gpurun_out/dist_check_world$NG.log | head
run() { name=$1; shift
  timeout 600 env $GVENV $TR bench.py --gpus $NG --steps 3 --warmup 3 --no-int8-pass "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "bench $name rc=$?"
  python tools/show_bench.py gpurun_out/r02_bench_$name.json | grep -v "host phases"
}
GVENV="GV_SHARD_PUSH=dma" run C4_n${NG}_sharded_dma --front sharded
GVENV="GV_SHARD_PUSH=stores" run C4_n${NG}_sharded_stores --front sharded
