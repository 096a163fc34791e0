# -*- coding: utf-8 -*-
"""BASELINE.json configs[3]: BERT-base synthetic (seq_len <= 512) with ``BucketedDistributedSampler`` (bucket by length),
DDP bf16, through the ``Stoke`` API.  Secondary benchmark (bench.py carries the headline line): exercises the sampler
feeding length-bucketed batches and the engine at 109.5 M parameters.

    python bench_bert.py [--steps 20] [--warmup 5]                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_bert.py
"""
import argparse
import faulthandler
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    faulthandler.enable()
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--n", type=int, default=1_000_003)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from transformers import BertConfig, BertForSequenceClassification

    import stoke_b200 as sb
    from stoke_b200 import synthetic
    from stoke_b200.data import DevicePrefetcher

    torch.manual_seed(0)
    model = BertForSequenceClassification(BertConfig())
    bs = 32
    s = sb.Stoke(model=model, optimizer=sb.StokeOptimizer(optimizer=torch.optim.AdamW, optimizer_kwargs={"lr": 1e-4}),
                 loss=torch.nn.CrossEntropyLoss(), batch_size_per_device=bs,
                 grad_clip=sb.ClipGradNormConfig(max_norm=1.0, norm_type=2.0), gpu=True, fp16="bf16",
                 distributed="ddp" if world > 1 else None,
                 configs=[sb.DDPConfig(local_rank=local)] if world > 1 else None, verbose=False)
    eng, path = s.engine, s.optimizer.path

    # dataset: lengths -> device argsort -> bucketed sampler (this replica's epoch indices)
    lens = synthetic.sampler_lengths(args.n, 16, 513)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    sorted_idx = sb.argsort_lengths(lens)
    smp = sb.BucketedDistributedSampler(range(args.n), buckets=16, batch_size=bs, sorted_idx=sorted_idx,
                                        num_replicas=world, rank=rank if world > 1 else 0, shuffle=True, seed=0,
                                        info_rank=-1)
    idx = smp.indices_tensor().cpu().numpy()
    t1.record(); torch.cuda.synchronize()
    sampler_ms = t0.elapsed_time(t1)
    rng = np.random.default_rng(1234 + rank)

    def host_batches():
        for b in range(len(idx) // bs):
            ii = idx[b * bs:(b + 1) * bs]
            ll = lens[ii]
            L = int((ll.max() + 31) // 32 * 32)  # few distinct shapes: cuBLASLt / SDPA heuristics are cached per shape
            ids = torch.from_numpy(rng.integers(0, 30522, size=(bs, L))).pin_memory()
            mask = torch.from_numpy((np.arange(L)[None, :] < ll[:, None]).astype(np.int64)).pin_memory()
            y = torch.from_numpy(rng.integers(0, 2, size=(bs,))).pin_memory()
            yield ids, mask, y, int(ll.sum()), bs * L

    feed = iter(DevicePrefetcher(({"ids": a, "mask": m, "y": y, "tok": t, "pad": p} for a, m, y, t, p in host_batches())))
    tokens = padded = 0

    def step():
        nonlocal tokens, padded
        b = next(feed)
        out = s.model(input_ids=b["ids"], attention_mask=b["mask"])
        s.backward(s.loss(out.logits, b["y"]))
        s.step()
        tokens += b["tok"]; padded += b["pad"]

    for _ in range(args.warmup):
        step()
    tokens = padded = 0
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    tok = torch.tensor([float(tokens), float(padded)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(tok)
    k2_ms, k2_n = eng.profile_read(1)
    k1_ms, k1_n, _ = eng.profile_read_k1_device()
    eng.profile(False)
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        peak = 6650.0
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except OSError:
            pass
        k2 = k2_ms / max(k2_n, 1)
        line = {"metric": "samples/sec", "value": bs * world * args.steps / sec, "unit": "samples/sec", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 / args.steps,
                "tokens_per_sec": float(tok[0].item()) / sec, "padding_efficiency": float(tok[0].item() / tok[1].item()),
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "bert_base_synthetic_bucketed_sampler_ddp_bf16_adamw_clipnorm1.0",
                           "per_gpu_batch": bs, "dataset_items": args.n, "buckets": 16, "params": path.n,
                           "e2e": "batches built on the host, copied by the prefetcher inside the timed region"},
                "sampler_setup_ms": sampler_ms,
                "roofline": {"kernel": "k_optim_step (AdamW, bf16 copy)", "bound": "hbm", "bytes_per_launch": path.n_local * 30,
                             "ms_per_launch": k2, "achieved": path.n_local * 30 / (k2 * 1e-3) / 1e9, "peak": peak,
                             "frac": path.n_local * 30 / (k2 * 1e-3) / 1e9 / peak,
                             "k1_ms_device": (k1_ms / max(k1_n, 1)) if world > 1 else None}}
        print(json.dumps(line))
        if args.out:
            with open(args.out, "w") as f:
                json.dump(line, f, indent=1)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
