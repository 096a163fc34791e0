# -*- coding: utf-8 -*-
"""bench.py -- the hot path's headline benchmark (BASELINE.json: samples/sec, ResNet-50 synthetic, DDP-mode, bf16 mixed
precision, grad_clip=1.0), measured through the ``Stoke`` API.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference] [--oss]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One step = ``out = s.model(x); l = s.loss(out, y); s.backward(l); s.step()`` on one synthetic batch per GPU.
Prints ONE JSON line on rank 0:
  value      samples/sec over all N GPUs with the batch already resident in HBM (device-timed, max over ranks)
  e2e        the same loop with the batch copied from pinned host memory every step and the loss read back to the host
  roofline   the dominant kernel of the engine (K2, the fused optimizer step): algorithmic bytes / CUDA-event duration,
             measured live over the timed region, against MEASURED_PEAKS.json (hbm_gbs); ``k1`` is reported beside it
  cpu_baseline  the oracle port of the reference's CPU call order (oracle/stoke_port.py) on a bounded sample (rank 0, N=1)
``--impl reference`` times that CPU port alone (the reference is pure Python and does not travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "samples/sec"
WORKLOAD = "resnet50_synthetic_3x224x224_ddp_bf16_adam_clipnorm1.0"
ADAM = {"lr": 1e-3}
CPU_SAMPLE_BATCH = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--oss", action="store_true", help="configs[2]: sharded (ZeRO-1) optimizer")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling helper: warm up, then run ONE step between cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints no bench line")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_arm(steps: int, warmup: int, batch: int = CPU_SAMPLE_BATCH):
    """The reference's CPU path (port, see oracle/stoke_port.py) on ResNet-50 fp32: samples/sec on the host cores."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from stoke_port import StokePortCPU

    from stoke_b200 import synthetic

    model = synthetic.resnet50()
    s = StokePortCPU(model, torch.optim.Adam, ADAM, torch.nn.CrossEntropyLoss(), grad_accum_steps=1,
                     clip=("norm", 1.0, 2.0))
    x, y = synthetic.resnet50_batch(batch)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = s.loss(s.model(x), y)
        s.backward(loss)
        s.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": METRIC, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"resnet50 fp32, batch {batch}, {len(times)} steps after {warmup} warm-up "
                      f"(median {statistics.median(times) * 1e3:.0f} ms/step, os.cpu_count()={os.cpu_count()})",
            "ms_per_step": total / len(times) * 1e3}


class ClockSampler:
    """nvidia-smi sampled in the background during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        base = cpu_reference_arm(max(2, min(args.steps, 6)), max(1, min(args.warmup, 2)))
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": METRIC, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "note": "reference CPU call order (oracle port), fp32, bounded sample",
                           "per_step_batch": CPU_SAMPLE_BATCH},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch

    import stoke_b200 as sb
    from stoke_b200 import synthetic

    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", local_rank)

    model = synthetic.resnet50().to(memory_format=torch.channels_last)
    configs = [sb.DDPConfig(local_rank=local_rank)] if world > 1 else None
    s = sb.Stoke(model=model, optimizer=sb.StokeOptimizer(optimizer=torch.optim.Adam, optimizer_kwargs=ADAM),
                 loss=torch.nn.CrossEntropyLoss(), batch_size_per_device=args.batch,
                 grad_clip=sb.ClipGradNormConfig(max_norm=1.0, norm_type=2.0), gpu=True, fp16="bf16",
                 distributed="ddp" if world > 1 else None, fairscale_oss=bool(args.oss and world > 1), configs=configs,
                 verbose=False)
    eng = s.engine
    path = s.optimizer.path
    x_host, y_host = synthetic.resnet50_batch(args.batch, rank)
    x_host = x_host.contiguous(memory_format=torch.channels_last).pin_memory()
    y_host = y_host.pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step_resident():
        s.backward(s.loss(s.model(x_dev), y_dev))
        s.step()

    from stoke_b200.data import DevicePrefetcher

    def host_batches():
        while True:
            yield x_host, y_host   # the same pinned batch every step: the H2D copy is real, the data is synthetic

    feed = iter(DevicePrefetcher(host_batches()))  # what StokeDataLoader uses: batch i+1 is copied while i computes

    def step_e2e():
        x, y = next(feed)                   # 77 MB host -> device copy per step, inside the timed region
        s.backward(s.loss(s.model(x), y))   # s.loss reads the synced loss back to the host every step
        s.step()

    def fence():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        fence()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(3, args.warmup)):
        step_resident()
    if args.ncu_step:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        return

    # ---- timed region 1: inputs resident; kernel events recorded live for the roofline ----
    eng.profile(True)  # CUDA events recorded inside the library, immediately around each K1 / K2 launch
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = eng.launches
    ms_total = timed(step_resident, args.steps)
    launches = eng.launches - launches0
    clock_info = clocks.stop() if rank == 0 else None
    k_ms = {}
    for key, kind in (("k1", 0), ("k2", 1)):
        tot, cnt = eng.profile_read(kind)
        k_ms[key] = tot / max(cnt, 1)
    k1_dev_ms, k1_dev_n, k1_zero_ms = eng.profile_read_k1_device()
    # the device timer brackets the data phase between K1's barriers; at W = 1 there are no barriers (block 0 would only
    # time its own chunk), so the event pair recorded around the launch is the kernel time there
    k_ms["k1_device"] = k1_dev_ms / max(k1_dev_n, 1) if world > 1 else k_ms["k1"]
    eng.profile(False)

    # ---- timed region 2: end to end (H2D of the batch + D2H of the loss inside the timed region) ----
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    if rank != 0:
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    n_local, n = path.n_local, path.n
    k2_bytes = n_local * 30  # g,p,m,v read (16) + p,m,v write (12) + bf16 param write (2), per element
    # K1 at W = 1: HBM bytes (read bf16 grad, write fp32 main grad, zero the bucket).  At W > 1: bytes crossing each
    # direction of this GPU's NVLink = (W-1)/W * n * (b_in + b_out), b_in = 2 (bf16), b_out = 4 (fp32; 0 when sharded)
    if world > 1:
        k1_bytes = (world - 1) / world * n * (2 + (0 if path.sharded else 4))
    else:
        k1_bytes = n * (2 + 4 + 2)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except OSError:
        pass
    k2_bound, k2_peak, k2_name = "hbm", hbm_peak, "k_optim_step (K2 fused Adam + clip + bf16 param write)"
    if path.sharded:
        # sharded step: 1/W of the elements, and the updated bf16 shard is pushed to every rank from inside the kernel
        # (parameter all-gather): NVLink-bound, (W-1)/W * n * 2 B per direction; timed by events, so rank skew is included
        k2_bytes = (world - 1) / world * n * 2
        k2_bound, k2_peak, peak_src = "nvlink", 900.0, "NVLink 5 nominal per direction"
        k2_name = "k_optim_step (K2 sharded step + in-kernel bf16 parameter all-gather)"
        traffic = None
    roofline = {"kernel": k2_name, "bound": k2_bound,
                "achieved": k2_bytes / (k_ms["k2"] * 1e-3) / 1e9, "peak": k2_peak, "unit": "GB/s",
                "frac": k2_bytes / (k_ms["k2"] * 1e-3) / 1e9 / k2_peak, "traffic": traffic,
                "bytes_per_launch": k2_bytes, "ms_per_launch": k_ms["k2"], "peak_source": peak_src,
                "k1": {"kernel": "k_grad_reduce (K1)", "bytes_per_launch": k1_bytes,
                       "ms_per_launch_events": k_ms["k1"], "ms_per_launch": k_ms["k1_device"],
                       "ms_zero_tail": (k1_zero_ms / max(k1_dev_n, 1)) if world > 1 else 0.0,
                       "achieved": k1_bytes / (k_ms["k1_device"] * 1e-3) / 1e9, "unit": "GB/s",
                       "bound": "hbm" if world == 1 else "nvlink",
                       "peak": hbm_peak if world == 1 else 900.0,
                       "frac": k1_bytes / (k_ms["k1_device"] * 1e-3) / 1e9 / (hbm_peak if world == 1 else 900.0),
                       "note": "W>1: ms_per_launch = device timer between K1's start and end barriers = the NVLink phase (peer reads "
                               "+ peer writes); excludes the wait for the slowest rank's launch (in ms_per_launch_events) and the "
                               "local HBM zeroing of the bucket after the end barrier (ms_zero_tail); bytes per NVLink direction "
                               "(W-1)/W*n*(2+4), peak 900 GB/s nominal (770 measured peer copy).  W=1: event-timed, HBM bytes 8 B/elem"}}
    samples = args.batch * world * args.steps
    line = {"metric": METRIC, "value": samples / (ms_total * 1e-3), "unit": METRIC, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD + ("_oss" if path.sharded else ""), "per_gpu_batch": args.batch,
                       "global_batch": args.batch * world, "parallelism": f"dp{world}", "params": n,
                       "l2": "per-step working set (activations, 0.9 GB of optimizer state) exceeds the 126 MB L2; no flush"},
            "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": METRIC,
                    "h2d_bytes_per_step": x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size(),
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clock_info, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        base = cpu_reference_arm(4, 1)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))


if __name__ == "__main__":
    main()
