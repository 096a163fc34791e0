# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*`` by running the UNMODIFIED reference in this container.

    python oracle/make_golden.py            # needs /root/reference (build container only)

Fixtures written (all small, committed):

  tests/golden/sampler_golden.json   reference ``BucketedDistributedSampler`` (/root/reference/stoke/data.py:111-516) over a
                                     grid of (N, buckets, bs, W, drop_last, overlap, shuffle, seed, epoch): per case the
                                     per-replica length, rank-0 head, sha256 over all replicas' int64 lists; small cases
                                     also carry the full lists.  Includes SURVEY.md Appendix B's four cases.
  tests/golden/cfg1_*.npz            BASELINE.json configs[0]: BasicNN (128-256-256-1) + BCEWithLogitsLoss + Adam, CPU,
                                     fp32, grad_accum=2, 50 optimizer steps through the reference ``Stoke`` object
                                     (``DistributedNullCPU + NullFP16 + BaseOptimizer``): initial weights, final weights,
                                     per-micro-step losses, counter trace; variants without clip / clip-by-norm /
                                     clip-by-value.

The recipe for the data is in ``stoke_b200/synthetic.py`` so the GPU-side tests can rebuild the identical inputs.
"""
import hashlib
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
sys.path.insert(0, os.path.dirname(HERE))
from stoke_b200 import synthetic as workloads  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sampler_cases():
    cases = [
        # SURVEY.md Appendix B
        dict(n=1000, buckets=4, bs=8, w=2, drop_last=False, overlap=False, shuffle=True, seed=0, epoch=0),
        dict(n=1003, buckets=4, bs=8, w=2, drop_last=True, overlap=True, shuffle=True, seed=0, epoch=3),
        dict(n=1000, buckets=4, bs=8, w=2, drop_last=False, overlap=False, shuffle=False, seed=0, epoch=0),
        dict(n=5000, buckets=5, bs=16, w=8, drop_last=False, overlap=False, shuffle=True, seed=7, epoch=1),
    ]
    grid = [
        (1001, 3, 8, 4), (1237, 4, 5, 3), (2048, 2, 16, 8), (3001, 7, 4, 2), (997, 1, 16, 4), (4099, 8, 8, 1),
        (1500, 3, 7, 5), (10007, 10, 32, 8), (1024, 4, 16, 2), (1025, 4, 16, 2), (1279, 4, 16, 2),
    ]
    for (n, b, bs, w) in grid:
        for drop_last, overlap in ((False, False), (True, False), (True, True)):
            for shuffle in (True, False):
                cases.append(dict(n=n, buckets=b, bs=bs, w=w, drop_last=drop_last, overlap=overlap,
                                  shuffle=shuffle, seed=n % 11, epoch=n % 5))
    cases.append(dict(n=100003, buckets=16, bs=32, w=8, drop_last=False, overlap=False, shuffle=True, seed=0, epoch=0))
    cases.append(dict(n=100003, buckets=16, bs=32, w=8, drop_last=True, overlap=True, shuffle=True, seed=3, epoch=2))
    return cases


def run_reference_sampler(stoke, case):
    n = case["n"]
    sorted_idx = workloads.sampler_sorted_idx(n).tolist()
    ds = list(range(n))
    out = []
    for r in range(case["w"]):
        with redirect_stdout(io.StringIO()):
            s = stoke.BucketedDistributedSampler(
                ds, buckets=case["buckets"], batch_size=case["bs"], sorted_idx=sorted_idx,
                backend=stoke.DistributedOptions.ddp, allow_bucket_overlap=case["overlap"],
                num_replicas=case["w"], rank=r, shuffle=case["shuffle"], seed=case["seed"],
                drop_last=case["drop_last"], info_rank=-1,
            )
        s.set_epoch(case["epoch"])
        out.append([int(v) for v in iter(s)])
    return out


def make_sampler(stoke):
    rows = []
    for case in sampler_cases():
        row = dict(case)
        try:
            lists = run_reference_sampler(stoke, case)
        except (ValueError, AssertionError) as e:
            row["raises"] = type(e).__name__
            rows.append(row)
            continue
        h = hashlib.sha256()
        for lst in lists:
            h.update(np.asarray(lst, dtype="<i8").tobytes())
        row["len_per_replica"] = len(lists[0])
        row["rank0_head"] = lists[0][:16]
        row["sha256"] = h.hexdigest()
        if case["n"] <= 1300:
            row["lists"] = lists
        rows.append(row)
    with open(os.path.join(GOLD, "sampler_golden.json"), "w") as f:
        json.dump({"torch": torch.__version__, "numpy": np.__version__, "cases": rows}, f)
    print(f"sampler: {len(rows)} cases ({sum('raises' in r for r in rows)} raising)")


def make_cfg1(stoke):
    variants = {
        "noclip": None,
        "clipnorm": stoke.ClipGradNormConfig(max_norm=0.05, norm_type=2.0),
        "clipvalue": stoke.ClipGradConfig(clip_value=0.002),
    }
    for name, clip in variants.items():
        model = workloads.basic_nn()
        init = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy().copy()
        opt = stoke.StokeOptimizer(optimizer=torch.optim.Adam, optimizer_kwargs=workloads.CFG1_ADAM)
        with redirect_stdout(io.StringIO()):
            s = stoke.Stoke(model=model, optimizer=opt, loss=torch.nn.BCEWithLogitsLoss(),
                            batch_size_per_device=workloads.CFG1_BATCH, grad_accum_steps=workloads.CFG1_ACCUM,
                            grad_clip=clip, gpu=False, verbose=False)
        losses, trace = [], []
        for x, y in workloads.cfg1_batches(workloads.CFG1_OPT_STEPS * workloads.CFG1_ACCUM):
            out = s.model(x)
            l = s.loss(out, y)
            losses.append(s.step_loss)
            s.backward(l)
            s.step()
            trace.append((s._grad_accum_counter, s._backward_steps, s._optimizer_steps))
        final = torch.cat([p.detach().reshape(-1) for p in s.model_access.parameters()]).numpy()
        np.savez(os.path.join(GOLD, f"cfg1_{name}.npz"), init=init, final=final,
                 losses=np.asarray(losses, dtype=np.float64), trace=np.asarray(trace, dtype=np.int64))
        print(f"cfg1/{name}: final |w|={np.linalg.norm(final):.6f} last loss={losses[-1]:.6f} trace tail={trace[-1]}")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shim.import_reference()
    make_sampler(ref)
    make_cfg1(ref)
