NG=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
tools/gpu_env_probe.sh > gpurun_out/r02_env_n$NG.txt 2>&1; head -3 gpurun_out/r02_env_n$NG.txt
timeout 300 $TR tools/upload_bench_ranks.py > gpurun_out/r02_upload_ranks_n$NG.txt 2>&1; echo rc=$?
GV_STAGE_MEMCPY=1 GV_REG_CHUNK_MB=64 timeout 300 $TR tools/upload_bench_ranks.py >> gpurun_out/r02_upload_ranks_n$NG.txt 2>&1; echo rc=$?
grep -v "OMP_NUM\|\*\*\*\|NCCL version" gpurun_out/r02_upload_ranks_n$NG.txt
timeout 600 $TR tools/dist_check.py 2600 4000 > gpurun_out/dist_check_world$NG.log 2>&1; echo "dist_check rc=$?"; grep -c "identical to single rank: True" gpurun_out/dist_check_world$NG.log; grep "False\|DIST CHECK" gpurun_out/dist_check_world$NG.log | head -5
run() { name=$1; shift
  timeout 600 env $GVENV $TR bench.py --gpus $NG --steps 3 --warmup 3 --no-int8-pass "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "bench $name rc=$?"
  python tools/show_bench.py gpurun_out/r02_bench_$name.json | grep -v clocks
}
GVENV="GV_X=1" run C4_n${NG}_sharded --front sharded
GVENV="GV_X=1" run C4_n${NG}_bcast --front broadcast
GVENV="GV_X=1" run C3_n${NG}_sharded --front sharded --workload C3
GVENV="GV_X=1" run C3_n${NG}_bcast --front broadcast --workload C3
