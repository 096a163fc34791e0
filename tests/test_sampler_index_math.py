"""CPU test of the sampler's device index arithmetic: a numpy transcription of ``k_sampler_indices``
(stoke_b200/csrc/sampler.cu) fed with the library's own host-side products (``stk_sampler_plan``, ``stk_randperm``,
``stk_sampler_last_slice``) must reproduce the oracle's index lists -- so the plan, the padding tables, the permutation
layout and the per-output-index formula are all checked without a GPU (the CUDA kernel itself is compared bit-for-bit on
the GPU box in tests/test_gpu_sampler.py)."""
import ctypes as C

import numpy as np
import pytest

from sampler_oracle import oracle_indices
from stoke_b200 import _lib, synthetic


def emulate_device_indices(lib, n, buckets, bs, w, rank, shuffle, seed, epoch, drop_last, overlap, sorted_idx):
    p = _lib.SamplerPlan()
    p.n, p.buckets, p.batch_size, p.world, p.rank = n, buckets, bs, w, rank
    p.drop_last, p.allow_bucket_overlap, p.shuffle = int(drop_last), int(overlap), int(shuffle)
    if lib.stk_sampler_plan(C.byref(p)) != 0:
        raise ValueError(lib.stk_last_error(None).decode())
    if p.n_batches * p.batch_size != p.rounded_per_replica:
        raise AssertionError  # what BucketedDistributedSampler.indices_tensor raises (reference data.py:447)
    bucket_lens = [p.bucket_base + (1 if b < p.bucket_rem else 0) for b in range(buckets)]
    last_slice = None
    if p.needs_padding:
        last_slice = np.zeros((2, p.slice_size), dtype=np.int32)
        for row, ln in enumerate((p.bucket_base + 1, p.bucket_base)):
            if (row == 0 and p.bucket_rem == 0) or p.rounded_per_bucket <= ln:
                continue
            _lib.check(lib.stk_sampler_last_slice(C.byref(p), ln, last_slice[row].ctypes.data))
    bucket_perm = batch_perm = None
    if shuffle:
        lens = (C.c_int64 * buckets)(*bucket_lens)
        bucket_perm = np.zeros(n, dtype=np.int32)
        _lib.check(lib.stk_randperm(seed + epoch, lens, buckets, bucket_perm.ctypes.data))
        nb = (C.c_int64 * 1)(p.n_batches)
        batch_perm = np.zeros(p.n_batches, dtype=np.int32)
        _lib.check(lib.stk_randperm(seed + epoch, nb, 1, batch_perm.ctypes.data))
    # ---- transcription of k_sampler_indices, vectorised over j ----
    S, W, ns = p.slice_size, p.world, p.slices_per_bucket
    j = np.arange(p.rounded_per_replica, dtype=np.int64)
    bdst, t = j // bs, j % bs
    b = batch_perm[bdst].astype(np.int64) if batch_perm is not None else bdst
    in_slice = rank + t * W
    bucket = np.zeros_like(j)
    pos = np.zeros_like(j)
    main = b < p.n_bucket_batches
    bk = b[main] // ns
    k = b[main] % ns
    ln = p.bucket_base + (bk < p.bucket_rem)
    ps = k * S + in_slice[main]
    padded = (k == ns - 1) & (ns * S > ln)
    if padded.any():
        tab = np.where(ln[padded] == p.bucket_base, 1, 0)
        ps[padded] = last_slice[tab, in_slice[main][padded]]
    bucket[main], pos[main] = bk, ps
    if (~main).any():
        q = (b[~main] - p.n_bucket_batches) * S + in_slice[~main]
        r0 = p.bucket_base - p.rounded_per_bucket
        r1 = r0 + 1
        first = p.bucket_rem * r1
        early = q < first
        bk2 = np.where(early, q // max(r1, 1), p.bucket_rem + (q - first) // max(r0, 1))
        off = np.where(early, q % max(r1, 1), (q - first) % max(r0, 1))
        bucket[~main], pos[~main] = bk2, p.rounded_per_bucket + off
    start = bucket * p.bucket_base + np.minimum(bucket, p.bucket_rem)
    src = bucket_perm[start + pos].astype(np.int64) if bucket_perm is not None else pos
    return np.asarray(sorted_idx, dtype=np.int64)[start + src].tolist()


def test_index_math_matches_oracle_random_configs():
    lib = _lib.load()
    rng = np.random.default_rng(2024)
    ok = raised = 0
    for _ in range(150):
        w = int(rng.integers(1, 9)); bs = int(rng.integers(1, 33)); buckets = int(rng.integers(1, 9))
        lo = max(100, 2 * bs * w) * buckets
        n = int(rng.integers(max(lo - 50, 1), 4 * lo))
        drop_last, overlap, shuffle = (bool(rng.integers(0, 2)) for _ in range(3))
        seed, epoch, rank = int(rng.integers(0, 1000)), int(rng.integers(0, 10)), int(rng.integers(0, w))
        sorted_idx = synthetic.sampler_sorted_idx(n)
        args = (n, buckets, bs, w, rank, shuffle, seed, epoch, drop_last, overlap)
        try:
            ref = oracle_indices(sorted_idx, buckets, bs, w, rank, shuffle, seed, epoch, drop_last, overlap)
        except (ValueError, AssertionError) as e:
            with pytest.raises(type(e)):
                emulate_device_indices(lib, *args, sorted_idx)
            raised += 1
            continue
        assert emulate_device_indices(lib, *args, sorted_idx) == ref, args
        ok += 1
    assert ok > 60 and raised > 5


def test_index_math_full_size_cfg4():
    lib = _lib.load()
    n, buckets, bs, w = 1_000_003, 16, 32, 8
    sorted_idx = np.argsort(synthetic.sampler_lengths(n, 16, 513), kind="stable")
    for rank in (0, 7):
        got = emulate_device_indices(lib, n, buckets, bs, w, rank, True, 0, 0, False, False, sorted_idx)
        assert len(got) == 125_440
        assert got == oracle_indices(sorted_idx, buckets, bs, w, rank, True, 0, 0)
