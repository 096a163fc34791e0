"""GPU parity tests for the sampler: bit-exact against the reference fixtures, the oracle, and size-independent
properties at BASELINE's full size (N = 1,000,003)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sha(lists):
    h = hashlib.sha256()
    for lst in lists:
        h.update(np.asarray(lst, dtype="<i8").tobytes())
    return h.hexdigest()


def _gpu_lists(case, sorted_idx=None):
    import stoke_b200 as sb
    from stoke_b200 import synthetic

    n = case["n"]
    if sorted_idx is None:
        sorted_idx = synthetic.sampler_sorted_idx(n)
    out = []
    for r in range(case["w"]):
        s = sb.BucketedDistributedSampler(range(n), buckets=case["buckets"], batch_size=case["bs"],
                                          sorted_idx=sorted_idx, allow_bucket_overlap=case["overlap"],
                                          num_replicas=case["w"], rank=r, shuffle=case["shuffle"], seed=case["seed"],
                                          drop_last=case["drop_last"], info_rank=-1)
        s.set_epoch(case["epoch"])
        assert len(s) == s.rounded_num_samples_per_replica
        out.append(list(iter(s)))
    return out


def test_sampler_matches_reference_fixtures(golden_dir):
    with open(os.path.join(golden_dir, "sampler_golden.json")) as f:
        cases = json.load(f)["cases"]
    for case in cases:
        lists = _gpu_lists(case)
        assert len(lists[0]) == case["len_per_replica"], case
        assert lists[0][:16] == case["rank0_head"], case
        assert _sha(lists) == case["sha256"], case
        if "lists" in case:
            assert lists == case["lists"]


def test_sampler_random_configs_vs_oracle():
    from sampler_oracle import oracle_indices
    import stoke_b200 as sb
    from stoke_b200 import synthetic

    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(40):
        w = int(rng.integers(1, 9)); bs = int(rng.integers(2, 33)); buckets = int(rng.integers(1, 9))
        lo = max(100, 2 * bs * w) * buckets
        n = int(rng.integers(lo + 1, 5 * lo))
        drop_last, overlap, shuffle = (bool(rng.integers(0, 2)) for _ in range(3))
        seed, epoch = int(rng.integers(0, 1000)), int(rng.integers(0, 10))
        sorted_idx = synthetic.sampler_sorted_idx(n)
        r = int(rng.integers(0, w))
        try:
            ref = oracle_indices(sorted_idx, buckets, bs, w, r, shuffle, seed, epoch, drop_last, overlap)
        except (ValueError, AssertionError) as e:
            with pytest.raises(type(e)):
                s = sb.BucketedDistributedSampler(range(n), buckets, bs, sorted_idx, allow_bucket_overlap=overlap,
                                                  num_replicas=w, rank=r, shuffle=shuffle, seed=seed,
                                                  drop_last=drop_last, info_rank=-1)
                s.set_epoch(epoch)
                list(iter(s))
            continue
        s = sb.BucketedDistributedSampler(range(n), buckets, bs, sorted_idx, allow_bucket_overlap=overlap,
                                          num_replicas=w, rank=r, shuffle=shuffle, seed=seed, drop_last=drop_last,
                                          info_rank=-1)
        s.set_epoch(epoch)
        assert list(iter(s)) == ref
        checked += 1
    assert checked >= 20


def test_argsort_matches_numpy_stable():
    import stoke_b200 as sb
    from stoke_b200 import synthetic

    for n in (1, 7, 2048, 2049, 100_003, 1_000_003):
        lens = synthetic.sampler_lengths(n, 16, 513)
        got = sb.argsort_lengths(lens).cpu().numpy()
        assert np.array_equal(got, np.argsort(lens, kind="stable"))
    keys = np.random.default_rng(1).integers(0, 2**32, size=300_001, dtype=np.uint64).astype(np.int64)
    assert np.array_equal(sb.argsort_lengths(keys).cpu().numpy(), np.argsort(keys, kind="stable"))


def test_full_size_properties_cfg4():
    """BASELINE configs[3] sampler workload: N = 1,000,003, 16 buckets, bs 32, W 8.  Properties that do not need the
    oracle: per-replica length, every replica's batches stay inside one length bucket, the W replicas of a slice are
    disjoint and together cover exactly the bucketed samples once (padding only repeats in-bucket items), determinism per
    epoch and change across epochs.  Plus the oracle itself on two replicas (it finishes in ~1 s each)."""
    from sampler_oracle import oracle_indices
    import stoke_b200 as sb
    from stoke_b200 import synthetic

    n, buckets, bs, w = 1_000_003, 16, 32, 8
    lens = synthetic.sampler_lengths(n, 16, 513)
    sorted_idx = sb.argsort_lengths(lens)
    assert np.array_equal(sorted_idx.cpu().numpy(), np.argsort(lens, kind="stable"))
    samplers = [sb.BucketedDistributedSampler(range(n), buckets, bs, sorted_idx, num_replicas=w, rank=r, shuffle=True,
                                              seed=0, info_rank=-1) for r in range(w)]
    outs = [s.indices_tensor() for s in samplers]
    assert all(o.numel() == 125_440 for o in outs)
    bucket_of = torch.empty(n, dtype=torch.int64, device="cuda")
    bounds = np.cumsum([0] + [len(p) for p in np.array_split(np.arange(n), buckets)])
    for b in range(buckets):
        bucket_of[sorted_idx[bounds[b]: bounds[b + 1]]] = b
    for o in outs:
        per_batch = bucket_of[o].view(-1, bs)
        assert bool((per_batch == per_batch[:, :1]).all())
    allidx = torch.cat(outs)
    counts = torch.bincount(allidx, minlength=n)
    assert int(counts.min()) >= 1 and int(counts.max()) <= 2  # padded slices repeat a few items, nothing is dropped
    assert int((counts == 2).sum()) == w * 125_440 - n
    assert torch.equal(samplers[0].indices_tensor(), outs[0])
    samplers[0].set_epoch(1)
    assert not torch.equal(samplers[0].indices_tensor(), outs[0])
    for r in (0, 5):
        assert outs[r].tolist() == oracle_indices(sorted_idx.cpu().numpy(), buckets, bs, w, r, True, 0, 0)
