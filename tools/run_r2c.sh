mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -x -k nvls > gpurun_out/t_r2c_multi.log 2>&1; tail -3 gpurun_out/t_r2c_multi.log | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/mgpu_w2_nvls.json')); print(d['caps']); print({k:(v.get('rel_err'),v.get('mc'),v.get('buckets')) for k,v in d.items() if isinstance(v,dict) and 'rel_err' in v})"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_r2a.json 2> gpurun_out/bench_n2_r2a.err; echo bench2 rc=$?; tail -c 1500 gpurun_out/bench_n2_r2a.err; cut -c1-600 gpurun_out/bench_n2_r2a.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench_allreduce.py --max-mb 256 --out gpurun_out/allreduce_w2_r2a.json > gpurun_out/ar2_r2a.log 2>&1; echo sweep rc=$?; tail -3 gpurun_out/ar2_r2a.log | cut -c1-300
for h in 0 1; do for b in 0 2 4; do STK_K2_STREAM_HINT=$h STK_NORM_BLOCKS_PER_SM=$b python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('hint $h blocks $b', 'k1n us', round(k['k1']['ms_per_launch']*1e3,2), 'k2 us', round(k['k2']['ms_per_launch']*1e3,2), 'value', round(d['value'],1))"; done; done
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_optim_step|k_grad_norm" -o gpurun_out/prof_r2a python bench.py --ncu-step > gpurun_out/ncu_r2a.log 2>&1; echo ncu rc=$?
