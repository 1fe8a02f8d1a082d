NG=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_check.py 2600 4000 > gpurun_out/dist_check_world$NG.log 2>&1; echo "dist_check rc=$?"; grep -c "identical to single rank: True" gpurun_out/dist_check_world$NG.log; grep "False\|DIST CHECK" gpurun_out/dist_check_world$NG.log | head -5
run() { name=$1; shift
  timeout 600 env $GVENV $TR bench.py --gpus $NG --steps 3 --warmup 3 "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "bench $name rc=$?"
  python tools/show_bench.py gpurun_out/r02_bench_$name.json | grep -v clocks
}
GVENV="GV_X=1" run C4_n${NG}_sharded_v2 --front sharded
GVENV="GV_X=1" run C4_n${NG}_bcast_v2 --front broadcast --no-int8-pass
GVENV="GV_X=1" run C3_n${NG}_sharded_v2 --front sharded --workload C3 --no-int8-pass
GVENV="GV_X=1" run C5_pearson_n${NG}_v2 --target pearson
timeout 300 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_n8box.json 2>/dev/null; head -c 600 gpurun_out/r02_bench_reference_n8box.json
