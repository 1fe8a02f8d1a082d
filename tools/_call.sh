NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_gpu_runtime.py -m gpu -x -q > gpurun_out/r02_gpu_tests11.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gpu_tests11.log
grep -c "identical to single rank: True" gpurun_out/dist_check_world$NG.log; grep "False\|DIST CHECK" gpurun_out/dist_check_world$NG.log | head
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
run() { name=$1; shift
  timeout 600 env $GVENV $TR bench.py --gpus $NG --steps 3 --warmup 3 --no-int8-pass "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "bench $name rc=$?"
  python tools/show_bench.py gpurun_out/r02_bench_$name.json | grep -v "clocks"
}
GVENV="GV_FRONT_CHUNKS=4" run C4_n${NG}_sharded_chunks4 --front sharded
GVENV="GV_FRONT_CHUNKS=1" run C4_n${NG}_sharded_chunks1 --front sharded
GVENV="GV_X=1" run C4_n${NG}_bcast_v3 --front broadcast
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-int8-pass > gpurun_out/r02_bench_C4_f.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02_bench_C4_f.json | grep -v clocks
