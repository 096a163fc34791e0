mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for keep in 0 1; do STK_NORM_L2_KEEP=$keep python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('keep $keep N1 k1n us', round(k['k1']['ms_per_launch']*1e3,2), 'k2 us', round(k['k2']['ms_per_launch']*1e3,2), 'value', round(d['value'],1), 'engine', round(d['roofline']['engine']['frac'],3))"; done
port=29800
for cfg in "1 64" "1 32" "1 16" "0 32"; do set -- $cfg; port=$((port+1))
STK_COOP_LAUNCH=$1 STK_K1_MAX_BLOCKS=$2 timeout 200 $B --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --no-parity 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('coop $1 maxblk $2', 'k1 dev us', round(k['k1']['ms_per_launch']*1e3,1), 'ev', round(k['k1']['ms_per_launch_events']*1e3,1), 'k2 us', round(k['k2']['ms_per_launch']*1e3,1), 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1))"
done
timeout 300 python bench.py --workload bert --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bert_n1_r2b.err | grep '^{' | tail -1 > gpurun_out/bert_n1_r2b.json; tail -c 300 gpurun_out/bert_n1_r2b.err; python -c "
import json
d=json.loads(open('gpurun_out/bert_n1_r2b.json').read()); print('bert n1', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
for mb in 0 32; do port=$((port+1)); STK_K1_MAX_BLOCKS=$mb timeout 300 $B --master-port $port bench.py --gpus 2 --workload bert --steps 20 --warmup 3 --no-parity 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('bert n2 maxblk $mb', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], 'k1 dev us', round(k['k1']['ms_per_launch']*1e3,1), 'ev', round(k['k1']['ms_per_launch_events']*1e3,1), 'k2', round(k['k2']['ms_per_launch']*1e3,1))"; done
