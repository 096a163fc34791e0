"""Property tests (hypothesis) on the sampler, CPU only: invariants that hold for every valid configuration, checked on the
oracle and on the numpy transcription of the device index arithmetic fed by the library's host functions
(tests/test_sampler_index_math.py::emulate_device_indices).  SURVEY.md section 4 names these properties: per-replica length
== len(sampler), batches stay inside one length bucket, the W replicas of a slice are disjoint, nothing is dropped without
drop_last, determinism per (seed, epoch)."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from sampler_oracle import SamplerPlan, oracle_indices
from stoke_b200 import _lib
from test_sampler_index_math import emulate_device_indices


@st.composite
def configs(draw):
    w = draw(st.integers(1, 8))
    bs = draw(st.integers(1, 16))
    buckets = draw(st.integers(1, 6))
    lo = max(100, 2 * bs * w) * buckets
    n = draw(st.integers(lo + 1, lo * 3))
    return dict(n=n, buckets=buckets, bs=bs, w=w, shuffle=draw(st.booleans()), seed=draw(st.integers(0, 10_000)),
                epoch=draw(st.integers(0, 50)), drop_last=draw(st.booleans()))


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(configs())
def test_sampler_invariants(cfg):
    n, buckets, bs, w = cfg["n"], cfg["buckets"], cfg["bs"], cfg["w"]
    rng = np.random.default_rng(cfg["seed"])
    lengths = rng.integers(1, 513, size=n)
    sorted_idx = np.argsort(lengths, kind="stable")
    try:
        plan = SamplerPlan(n, buckets, bs, w, cfg["drop_last"], False)
    except ValueError:
        return
    lib = _lib.load()
    replicas = []
    for r in range(w):
        ref = oracle_indices(sorted_idx, buckets, bs, w, r, cfg["shuffle"], cfg["seed"], cfg["epoch"], cfg["drop_last"])
        got = emulate_device_indices(lib, n, buckets, bs, w, r, cfg["shuffle"], cfg["seed"], cfg["epoch"], cfg["drop_last"],
                                     False, sorted_idx)
        assert got == ref
        assert len(ref) == plan.rounded_per_replica
        replicas.append(np.asarray(ref))
    # batches stay inside one bucket of the length-sorted order
    bounds = np.cumsum([0] + [len(p) for p in np.array_split(np.arange(n), buckets)])
    bucket_of = np.empty(n, dtype=np.int64)
    for b in range(buckets):
        bucket_of[sorted_idx[bounds[b]: bounds[b + 1]]] = b
    for rep in replicas:
        per_batch = bucket_of[rep].reshape(-1, bs)
        assert (per_batch == per_batch[:, :1]).all()
    counts = np.bincount(np.concatenate(replicas), minlength=n)
    if not cfg["drop_last"]:
        assert counts.min() >= 1                      # nothing dropped; padding only repeats in-bucket items
        assert counts.sum() - n == w * plan.rounded_per_replica - n
    else:
        assert counts.max() <= 1                      # no padding with drop_last: every index at most once
    # determinism for the same (seed, epoch)
    again = oracle_indices(sorted_idx, buckets, bs, w, 0, cfg["shuffle"], cfg["seed"], cfg["epoch"], cfg["drop_last"])
    assert again == replicas[0].tolist()
