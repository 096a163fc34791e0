mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_n1_r2b.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r2b.json').read()); k=d['roofline']['kernels']
print('N1 k1n us', round(k['k1']['ms_per_launch']*1e3,2), 'k2 us', round(k['k2']['ms_per_launch']*1e3,2), 'value', round(d['value'],1), 'engine', d['roofline']['engine']['frac'])
PY
for cfg in "1 0" "1 64" "1 32" "1 16" "0 0" "0 32"; do set -- $cfg
STK_COOP_LAUNCH=$1 STK_K1_MAX_BLOCKS=$2 timeout 200 $B --master-port 297$2$1 bench.py --gpus 2 --steps 20 --warmup 5 --no-parity 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('coop $1 maxblk $2', 'k1 dev us', round(k['k1']['ms_per_launch']*1e3,1), 'ev', round(k['k1']['ms_per_launch_events']*1e3,1), 'k2 us', round(k['k2']['ms_per_launch']*1e3,1), 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1))"
done
timeout 300 python bench.py --workload bert --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bert_n1_r2a.err | grep '^{' | tail -1 > gpurun_out/bert_n1_r2a.json; tail -c 400 gpurun_out/bert_n1_r2a.err; cut -c1-400 gpurun_out/bert_n1_r2a.json
timeout 300 $B --master-port 29790 bench.py --gpus 2 --workload bert --steps 10 --warmup 3 2> gpurun_out/bert_n2_r2a.err | grep '^{' | tail -1 > gpurun_out/bert_n2_r2a.json; tail -c 400 gpurun_out/bert_n2_r2a.err; cut -c1-400 gpurun_out/bert_n2_r2a.json
timeout 200 $B --master-port 29791 bench.py --gpus 2 --workload allreduce_sweep 2> gpurun_out/sweep_n2_r2a.err | grep '^{' | tail -1 | cut -c1-300; tail -c 300 gpurun_out/sweep_n2_r2a.err
python -m pytest tests/test_gpu_multi.py -q -x -k bulk > gpurun_out/t_r2d_multi.log 2>&1; tail -3 gpurun_out/t_r2d_multi.log | cut -c1-300
