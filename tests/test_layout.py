"""CPU tests of the flat-buffer planning (stoke_b200/layout.py): parameter offsets, gradient buckets, per-bucket shards and
the segment tables -- the logic that must come out identical on every rank.  ``shard_range`` is the Python twin of the
library's ``stk_shard_range`` (checked against the C function, which needs no GPU)."""
import ctypes as C
import random

import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from stoke_b200 import _lib, layout


def _c_shard_range(n, world, rank):
    lib = _lib.load()
    b, e = C.c_size_t(), C.c_size_t()
    assert lib.stk_shard_range(n, world, rank, C.byref(b), C.byref(e)) == 0
    return b.value, e.value


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 5_000_000).map(lambda v: v // 8 * 8), world=st.sampled_from([1, 2, 3, 4, 8]))
def test_shard_range_matches_the_library_and_partitions(n, world):
    prev = 0
    for r in range(world):
        b, e = layout.shard_range(n, world, r)
        assert (b, e) == _c_shard_range(n, world, r)
        assert b == prev and e >= b and (b % 16 == 0 or b == n) and (e % 16 == 0 or e == n)
        prev = e
    assert prev == n


@settings(max_examples=100, deadline=None)
@given(numels=st.lists(st.integers(1, 300_000), min_size=1, max_size=40), world=st.sampled_from([1, 2, 4, 8]),
       cap=st.integers(16, 400_000), sharded=st.booleans())
def test_buckets_and_segments_tile_the_flat_buffer(numels, world, cap, sharded):
    offsets, padded, n = layout.param_offsets(numels)
    assert all(o % layout.ALIGN_ELEMS == 0 for o in offsets) and n % layout.ALIGN_ELEMS == 0
    assert all(p >= k and p - k < layout.ALIGN_ELEMS for p, k in zip(padded, numels))
    buckets = layout.plan_buckets(offsets, n, cap)
    # launch order = reverse registration order; contiguous; cut at parameter boundaries; tile [0, n)
    assert buckets[0][1] == n and buckets[-1][0] == 0
    for (b0, b1), (c0, c1) in zip(buckets, buckets[1:]):
        assert c1 == b0 and c0 < c1
    assert all(b0 in offsets for b0, _ in buckets)
    assert len(buckets) <= layout.MAX_SEGMENTS
    # every bucket but the last reaches the cap (the last takes what is left)
    assert all(b1 - b0 >= min(cap, n) for b0, b1 in buckets[:-1]) or len(buckets) == layout.MAX_SEGMENTS
    segs, n_local = layout.plan_segments(buckets, world, sharded)
    assert len(segs) == world
    covered = []
    for r in range(world):
        lo = 0
        last_g = -1
        for g0, g1, l0, k in segs[r]:
            assert l0 == lo and g1 > g0 and g0 > last_g           # local offsets contiguous, ascending global order
            assert g0 % 16 == 0 and (g1 % 16 == 0 or g1 == n)
            if sharded:
                b0, b1 = buckets[k]
                assert b0 <= g0 and g1 <= b1                        # a segment never straddles a bucket
            lo += g1 - g0
            last_g = g0
            covered.append((g0, g1))
        assert lo == n_local[r]
    if sharded:
        covered.sort()
        pos = 0
        for g0, g1 in covered:   # the ranks' segments partition [0, n)
            assert g0 == pos
            pos = g1
        assert pos == n
    else:
        assert all(s == [(0, n, 0, 0)] for s in segs) and n_local == [n] * world


def test_resnet50_sized_plan():
    random.seed(0)
    numels = [random.choice([64, 256, 512, 2048, 36864, 147456, 589824, 2359296]) for _ in range(161)]
    offsets, _, n = layout.param_offsets(numels)
    buckets = layout.plan_buckets(offsets, n, 25 * (1 << 20) // 2)   # 25 MiB of bf16
    assert sum(b1 - b0 for b0, b1 in buckets) == n
    segs, n_local = layout.plan_segments(buckets, 8, True)
    assert sum(n_local) == n and max(n_local) - min(n_local) <= 16 * len(buckets) * 8


def test_parameter_group_layout_matches_torch_numbering():
    """``optim._group_layout`` + ``state_dict_index``: torch-style parameter groups given through ``optimizer_kwargs["params"]``
    map every trainable parameter (registration order) to its group, and the fused optimizer's ``state_dict()`` numbers the
    parameters exactly like the torch optimizer built from the same groups."""
    import torch

    from stoke_b200.optim import _group_layout, fused_supported, state_dict_index

    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.LayerNorm(8), torch.nn.Linear(8, 2))
    net[2].bias.requires_grad_(False)                       # frozen parameters are not part of the layout
    decay = [p for n, p in net.named_parameters() if p.requires_grad and p.ndim > 1]
    no_decay = [p for n, p in net.named_parameters() if p.requires_grad and p.ndim == 1]
    kw = {"lr": 1e-3, "weight_decay": 0.1, "params": [{"params": no_decay, "weight_decay": 0.0}, {"params": decay}]}
    params, group_of, extra, rest = _group_layout(net, kw)
    assert [p.shape for p in params] == [p.shape for p in net.parameters() if p.requires_grad]
    assert group_of == [1, 0, 0, 0, 1] and extra == [{"weight_decay": 0.0}, {}] and rest == {"lr": 1e-3, "weight_decay": 0.1}
    # (torch writes the defaults INTO the group dicts it is given: hand it copies)
    ref = torch.optim.AdamW([dict(g) for g in kw["params"]], lr=1e-3, weight_decay=0.1)
    ids = {id(p): i for g in ref.param_groups for p, i in zip(g["params"], ref.state_dict()["param_groups"][ref.param_groups.index(g)]["params"])}
    index = state_dict_index(group_of, len(extra))
    assert [index[i] for i in range(len(params))] == [ids[id(p)] for p in params]
    assert fused_supported(torch.optim.AdamW, kw, net)
    assert not fused_supported(torch.optim.Adam, dict(kw, amsgrad=True), net)       # stock-optimizer route
    assert not fused_supported(torch.optim.RMSprop, {"lr": 1e-3}, net)
    with pytest.raises(ValueError):
        _group_layout(net, {"params": [{"params": decay}]})                          # a trainable parameter left out
