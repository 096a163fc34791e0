mkdir -p gpurun_out
export STK_SPIN_TIMEOUT_S=60
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_multi.py -q -x -k "8-bulk or 8-nvls" > gpurun_out/t_r2g_multi.log 2>&1; tail -3 gpurun_out/t_r2g_multi.log | cut -c1-400
timeout 300 $B --master-port 29901 bench.py --gpus 8 --steps 50 --warmup 5 2> gpurun_out/bench_n8_r2a.err | grep '^{' | tail -1 > gpurun_out/bench_n8_r2a.json; echo bench8 rc=$? $(wc -c < gpurun_out/bench_n8_r2a.json); tail -c 300 gpurun_out/bench_n8_r2a.err
timeout 200 $B --master-port 29902 bench.py --gpus 8 --steps 20 --warmup 5 --oss --no-parity 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_n8_oss_r2a.json; echo oss rc=$?
STK_DDP_SHARD=0 timeout 200 $B --master-port 29903 bench.py --gpus 8 --steps 20 --warmup 5 --no-parity 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_n8_allreduce_r2a.json; echo ar rc=$?
timeout 300 $B --master-port 29904 bench.py --gpus 8 --workload bert --steps 20 --warmup 3 --no-parity 2> gpurun_out/bert_n8_r2a.err | grep '^{' | tail -1 > gpurun_out/bert_n8_r2a.json; echo bert rc=$?; tail -c 300 gpurun_out/bert_n8_r2a.err
timeout 400 $B --master-port 29905 bench_allreduce.py --out gpurun_out/allreduce_w8_r2a.json > gpurun_out/ar8_r2a.log 2>&1; echo sweep rc=$?; tail -2 gpurun_out/ar8_r2a.log | cut -c1-200
python - <<'PY'
import json
for f in ("bench_n8_r2a","bench_n8_oss_r2a","bench_n8_allreduce_r2a","bert_n8_r2a"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read()); k=d["roofline"]["kernels"]
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {n:(round(v["ms_per_launch"]*1e3,1), round(v.get("ms_per_launch_events",0)*1e3,1), round(v["frac"],3)) for n,v in k.items()}, (d.get("parity_check") or {}).get("ok"))
    except Exception as e: print(f, "ERR", e)
try:
    d=json.load(open("gpurun_out/allreduce_w8_r2a.json"))
    for r in d["rows"]: print(r["bytes"]>>10, "bulk", round(r["ours_bf16_busbw"]), round(r["ours_fp32_busbw"]), "nvls", round(r.get("ours_nvls_bf16_busbw",0)), round(r.get("ours_nvls_fp32_busbw",0)), "nccl", round(r.get("nccl_busbw",0)), "mm", round(r.get("symm_multimem_busbw",0)), "| us bulk/nvls/nccl/mm", round(r["ours_bf16_us"],1), round(r.get("ours_nvls_bf16_us",0),1), round(r.get("nccl_us",0),1), round(r.get("symm_multimem_us",0),1))
    print(d["spot_check_ok"], d["nvls_spot_check_ok"])
except Exception as e: print("sweep ERR", e)
PY
