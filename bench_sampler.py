# -*- coding: utf-8 -*-
"""BASELINE configs[3], sampler part: per-epoch, per-replica latency of ``BucketedDistributedSampler`` at N = 1,000,003
(16 buckets, batch 32, W = 8, shuffle) and of the stable argsort that produces ``sorted_idx``; device path vs the CPU
oracle port (numpy restatement of the reference; the reference's own list-based ``__iter__`` measured 0.52 s per replica
in the build container).

    python bench_sampler.py [--out profiles/sampler_rNN.json]
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from sampler_oracle import oracle_indices

    import stoke_b200 as sb
    from stoke_b200 import synthetic

    n, buckets, bs, w, rank = 1_000_003, 16, 32, 8, 3
    lens = synthetic.sampler_lengths(n, 16, 513)
    torch.cuda.set_device(0)
    res = {"n": n, "buckets": buckets, "batch_size": bs, "world": w}

    def med(fn, reps):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts) * 1e3

    # argsort
    res["argsort_cpu_numpy_stable_ms"] = med(lambda: np.argsort(lens, kind="stable"), 5)
    keys_dev = torch.as_tensor(lens).cuda()
    sb.argsort_lengths(keys_dev)
    res["argsort_gpu_ms"] = med(lambda: sb.argsort_lengths(keys_dev), 10)
    sorted_idx = sb.argsort_lengths(keys_dev)
    assert np.array_equal(sorted_idx.cpu().numpy(), np.argsort(lens, kind="stable"))
    # epoch indices
    smp = sb.BucketedDistributedSampler(range(n), buckets, bs, sorted_idx, num_replicas=w, rank=rank, shuffle=True,
                                        seed=0, info_rank=-1)
    smp.indices_tensor()
    res["iter_gpu_indices_tensor_ms"] = med(lambda: smp.indices_tensor(), 10)
    res["iter_gpu_python_list_ms"] = med(lambda: list(iter(smp)), 5)
    host_sorted = sorted_idx.cpu().numpy()
    ref = oracle_indices(host_sorted, buckets, bs, w, rank, True, 0, 0)
    assert list(iter(smp)) == ref
    res["iter_cpu_oracle_port_ms"] = med(lambda: oracle_indices(host_sorted, buckets, bs, w, rank, True, 0, 0), 3)
    res["iter_cpu_reference_ms_build_container"] = 520.0
    # where the device path spends its time
    import ctypes as C
    from stoke_b200 import _lib
    lib = _lib.load()
    plan = smp._plan
    lens_arr = (C.c_int64 * buckets)(*smp._bucket_lens)
    host = np.empty(plan.n + plan.n_batches, dtype=np.int32)
    res["host_randperm_mt19937_ms"] = med(lambda: lib.stk_randperm(0, lens_arr, buckets, host.ctypes.data), 5)
    g = torch.Generator()

    def torch_perm():
        g.manual_seed(0)
        for ln in smp._bucket_lens:
            torch.randperm(ln, generator=g)

    res["host_torch_randperm_ms"] = med(torch_perm, 5)
    res["cores"] = torch.get_num_threads()
    print(json.dumps(res))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
