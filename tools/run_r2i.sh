mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29840
run() { wl=$1; shift; port=$((port+1)); env "$@" timeout 300 $B --master-port $port bench.py --gpus 2 --workload $wl --steps 30 --warmup 5 --no-parity 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$wl $*', '| ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'k1 dev', round(k['k1']['ms_per_launch']*1e3,1), 'ev', round(k['k1']['ms_per_launch_events']*1e3,1), 'k2 dev', round(k['k2']['ms_per_launch']*1e3,1), d['config']['grad_buckets'])"; }
run resnet50 A=1
run bert STK_OVERLAP=off
run bert STK_OVERLAP=on
run bert STK_OVERLAP=on STK_K1_OVERLAP_BLOCKS=32
run bert STK_OVERLAP=on STK_K1_OVERLAP_BLOCKS=32 STK_K1_ALGO=ldg
run bert STK_OVERLAP=off A=2
python bench.py --workload bert --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bert N1 ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
