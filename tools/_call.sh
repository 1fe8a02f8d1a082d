NG=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
GV_FRONT_CHUNK_NNZ=20000 timeout 600 $TR tools/dist_check.py 2600 4000 > gpurun_out/dist_check_world8.log 2>&1; echo "dist_check8 rc=$?"; grep -c "identical to single rank: True" gpurun_out/dist_check_world8.log; grep "False\|DIST CHECK" gpurun_out/dist_check_world8.log | head -3
run() { name=$1; shift
  timeout 600 env $GVENV $TR bench.py --gpus $NG --steps 3 --warmup 3 "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "bench $name rc=$?"
  python tools/show_bench.py gpurun_out/r02_bench_$name.json | grep -v "clocks\|host phases"
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$name.json')); print('   phases', {k: round(v,2) for k,v in d['phases_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],1), 'enqueue', {k: round(v,1) for k,v in d['e2e']['host_phases_ms'].items() if 'enqueue' in k or 'wait' in k})"
}
GVENV="GV_FRONT_CHUNKS=1" run C4_n8_chunks1 --no-int8-pass
GVENV="GV_FRONT_CHUNKS=4" run C4_n8_chunks4 --no-int8-pass
GVENV="GV_FRONT_CHUNKS=4" run C3_n8_v3 --workload C3 --no-int8-pass
export CUDA_VISIBLE_DEVICES=0,1,2,3
NG=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR tools/dist_check.py 2600 4000 > gpurun_out/dist_check_world4.log 2>&1; echo "dist_check4 rc=$?"; grep -c "identical to single rank: True" gpurun_out/dist_check_world4.log; grep "False\|DIST CHECK" gpurun_out/dist_check_world4.log | head -3
GVENV="GV_FRONT_CHUNKS=1" run C4_n4_default
