mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N1 ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
port=29820
run() { port=$((port+1)); env "$@" timeout 200 $B --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 --no-parity 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$*', '| ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'k1 dev', round(k['k1']['ms_per_launch']*1e3,1), 'ev', round(k['k1']['ms_per_launch_events']*1e3,1), 'k2 dev', round(k['k2']['ms_per_launch']*1e3,1), 'ev', round(k['k2'].get('ms_per_launch_events',0)*1e3,1), d['config']['mem_mode'], d['config']['multicast_bound'], d['config']['grad_buckets'])"; }
run A=1
run STK_MEM=ipc
run STK_MULTICAST=0
run STK_BUCKET_MB=1000
run STK_MEM=ipc STK_BUCKET_MB=1000
run STK_K2_AG=mc
run A=2
STK_K2_AG=mc python -m pytest tests/test_gpu_multi.py -q -x -k "bulk" > gpurun_out/t_r2h_multi.log 2>&1; tail -2 gpurun_out/t_r2h_multi.log | cut -c1-300
