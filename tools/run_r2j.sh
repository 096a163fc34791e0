mkdir -p gpurun_out
export STK_SPIN_TIMEOUT_S=60
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $B --master-port 29931 bench.py --gpus 8 --steps 50 --warmup 5 2> gpurun_out/bench_n8_r2b.err | grep '^{' | tail -1 > gpurun_out/bench_n8_r2b.json; echo bench8 rc=$? $(wc -c < gpurun_out/bench_n8_r2b.json); tail -c 200 gpurun_out/bench_n8_r2b.err
port=29940
run() { port=$((port+1)); env "$@" timeout 200 $B --master-port $port bench.py --gpus 8 --steps 25 --warmup 5 --no-parity 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_n8_var_$port.json; python -c "
import json
d=json.loads(open('gpurun_out/bench_n8_var_$port.json').read()); k=d['roofline']['kernels']; print('$*', '| ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'k1 dev', round(k['k1']['ms_per_launch']*1e3,1), 'ev', round(k['k1']['ms_per_launch_events']*1e3,1), 'k2 dev', round(k['k2']['ms_per_launch']*1e3,1), 'ev', round(k['k2'].get('ms_per_launch_events',0)*1e3,1), d['config']['grad_buckets'])"; }
run STK_K2_AG=mc
run STK_K2_PAIR=0
timeout 300 $B --master-port 29939 bench_allreduce.py --study --max-mb 256 --out gpurun_out/study_w8_r2a.json > gpurun_out/study_w8_r2a.log 2>&1; echo study rc=$?; grep -c busbw gpurun_out/study_w8_r2a.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n8_r2b.json").read()); k=d["roofline"]["kernels"]
print("default", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {n:(round(v["ms_per_launch"]*1e3,1), round(v.get("ms_per_launch_events",0)*1e3,1), round(v["frac"],3)) for n,v in k.items()}, (d.get("parity_check") or {}).get("ok"), d["config"]["grad_buckets"])
s=json.load(open("gpurun_out/study_w8_r2a.json"))
for r in s["nvls_grid"]: print("nvls", r["bytes"]>>20, "MiB blocks", r["blocks"], round(r["us"],1), "us", round(r["busbw"]))
for r in s["reduce_scatter"]: print("rs", r["algo"], r["blocks"], round(r["us"],1), "us", round(r["wire_gbs"]))
for r in s["allreduce_grid"]: print("ar bulk 64MiB", r["blocks"], round(r["us"],1), round(r["busbw"]))
PY
