mkdir -p gpurun_out
python -m pytest tests -m gpu -q --ignore=tests/test_gpu_multi.py > gpurun_out/t_r2k.log 2>&1; tail -3 gpurun_out/t_r2k.log | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 2> gpurun_out/bench_n1_r2c.err | grep '^{' | tail -1 > gpurun_out/bench_n1_r2c.json; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_r2c.json').read()); k=d['roofline']['kernels']; print('N1', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {n:(round(v['ms_per_launch']*1e3,2), round(v['frac'],3)) for n,v in k.items()}, d['roofline']['engine']['frac'], d['cpu_baseline'])"
python bench_kernels.py --out gpurun_out/kernels_r2b.json > gpurun_out/bk_r2b.log 2>&1; tail -1 gpurun_out/bk_r2b.log | cut -c1-200
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_optim_step|k_grad_norm" -o gpurun_out/prof_r2b python bench.py --ncu-step > gpurun_out/ncu_r2b.log 2>&1; echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r2.log 2>&1; echo launches rc=$? $(wc -l < gpurun_out/launches_r2.csv)
python bench_sampler.py > gpurun_out/sampler_r2.log 2>&1; tail -3 gpurun_out/sampler_r2.log | cut -c1-300
