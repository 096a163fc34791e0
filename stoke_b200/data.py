# -*- coding: utf-8 -*-
"""Data side of the path: ``BucketedDistributedSampler`` (/root/reference/stoke/data.py:111-516) with the index pipeline
on the device, and the ``StokeDataLoader`` device-placement shim (:24-108).

Per epoch and replica the sampler (1) draws the reference's permutations on the host with the library's bit-exact
mt19937 ``randperm`` (one generator across the buckets, a second one -- same seed -- for the batch shuffle; the swap
chain is serial by construction), (2) uploads them, (3) launches one gather kernel that maps every output position through
batch shuffle -> bucket/slice -> replica stride -> padding table or residual batches -> bucket permutation ->
``sorted_idx``, and (4) reads the int64 indices back.  ``argsort_lengths`` is the device radix argsort that produces
``sorted_idx`` from raw lengths (the user-side ``np.argsort(kind="stable")`` of the reference workflow).
"""
import ctypes as C
from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import DataLoader as DL
from torch.utils.data import Dataset, Sampler

from . import _lib
from .engine import get_engine
from .utils import place_data_on_gpu


class DevicePrefetcher:
    """Wraps an iterator of host batches: batch i+1 is copied to the GPU on a side stream while batch i is being consumed
    (the reference's ``place_data_on_gpu`` is a blocking ``.to("cuda")`` per tensor, stoke/utils.py:39-80).  Copies overlap
    compute when the host tensors are pinned (``DataLoader(pin_memory=True)``)."""

    def __init__(self, host_iter, fp16=None, device=None):
        self._it = iter(host_iter)
        self._fp16 = fp16
        self._device = torch.cuda.current_device() if device is None else device
        self._stream = torch.cuda.Stream(self._device)

    def _stage(self):
        try:
            batch = next(self._it)
        except StopIteration:
            return None
        with torch.cuda.stream(self._stream):
            return place_data_on_gpu(batch, self._fp16)

    @staticmethod
    def _record(batch, stream):
        if isinstance(batch, torch.Tensor):
            batch.record_stream(stream)
        elif isinstance(batch, (list, tuple)):
            for b in batch:
                DevicePrefetcher._record(b, stream)
        elif isinstance(batch, dict):
            for b in batch.values():
                DevicePrefetcher._record(b, stream)

    def __iter__(self):
        nxt = self._stage()
        while nxt is not None:
            cur_stream = torch.cuda.current_stream(self._device)
            cur_stream.wait_stream(self._stream)
            cur = nxt
            self._record(cur, cur_stream)
            nxt = self._stage()
            yield cur


class StokeDataLoader(DL):
    def __init__(self, dataset, gpu: bool, fp16=None, **kwargs):
        super().__init__(dataset, **kwargs)
        self._gpu = gpu
        self._fp16 = fp16

    def __iter__(self):
        if not self._gpu:
            yield from super().__iter__()
        else:
            yield from DevicePrefetcher(super().__iter__(), self._fp16)


def argsort_lengths(lengths, device: Optional[int] = None) -> torch.Tensor:
    """Stable argsort of non-negative integer lengths on the device (LSD radix sort); returns int64 indices (CUDA)."""
    eng = get_engine(device)
    dev = torch.device("cuda", eng.device)
    keys = lengths.to(dev) if isinstance(lengths, torch.Tensor) else torch.as_tensor(np.asarray(lengths)).to(dev)
    if keys.numel() == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    if keys.dtype not in (torch.int32, torch.int64, torch.uint8, torch.int16) or int(keys.min()) < 0 \
            or int(keys.max()) >= 2**32:
        raise ValueError("Stoke -- argsort_lengths needs integer keys in [0, 2**32)")
    k64 = keys.to(torch.int64)
    keys32 = torch.where(k64 >= 2**31, k64 - 2**32, k64).to(torch.int32).contiguous()  # same 32-bit pattern
    n = keys32.numel()
    out = torch.empty(n, dtype=torch.int64, device=dev)
    tmp = torch.empty(eng.lib.stk_argsort_tmp_bytes(n), dtype=torch.uint8, device=dev)
    _lib.check(eng.lib.stk_argsort_u32(eng.ctx, keys32.data_ptr(), n, out.data_ptr(), tmp.data_ptr(), eng._stream()),
               eng.ctx)
    eng.launches += 13
    return out


class BucketedDistributedSampler(Sampler):
    """Same constructor and results as the reference class (bit-identical index lists)."""

    def __init__(self, dataset: Dataset, buckets: int, batch_size: int, sorted_idx: Sequence[int], backend=None,
                 allow_bucket_overlap: bool = False, num_replicas: Optional[int] = None, rank: Optional[int] = None,
                 shuffle: bool = True, seed: int = 0, drop_last: bool = False, info_rank: int = 0) -> None:
        if num_replicas is None or rank is None:
            if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
                raise RuntimeError("Requires distributed package (torch.dist or hvd) to be available")
            num_replicas = torch.distributed.get_world_size() if num_replicas is None else num_replicas
            rank = torch.distributed.get_rank() if rank is None else rank
        self.num_replicas, self.rank = int(num_replicas), int(rank)
        self.epoch = 0
        self.drop_last, self.shuffle, self.seed = bool(drop_last), bool(shuffle), int(seed)
        self.buckets, self.batch_size = int(buckets), int(batch_size)
        self.sorted_n_samples = sorted_idx
        self.allow_bucket_overlap = bool(allow_bucket_overlap)
        lib = _lib.load()
        plan = _lib.SamplerPlan()
        plan.n, plan.buckets, plan.batch_size = len(dataset), self.buckets, self.batch_size
        plan.world, plan.rank = self.num_replicas, self.rank
        plan.drop_last, plan.allow_bucket_overlap, plan.shuffle = int(self.drop_last), int(self.allow_bucket_overlap), int(self.shuffle)
        code = lib.stk_sampler_plan(C.byref(plan))
        if code != 0:
            raise ValueError(lib.stk_last_error(None).decode())
        self._plan = plan
        self.slice_size = plan.slice_size
        self.num_samples_per_bucket = plan.per_bucket
        self.num_slices_per_bucket = plan.slices_per_bucket
        self.rounded_num_samples_per_bucket = plan.rounded_per_bucket
        self.rounded_num_samples_per_replica = plan.rounded_per_replica
        if len(sorted_idx) != len(dataset):
            raise ValueError("Stoke -- sorted_idx must hold one entry per dataset item")
        # host-side tables that do not change between epochs
        self._bucket_lens = [plan.bucket_base + (1 if b < plan.bucket_rem else 0) for b in range(self.buckets)]
        self._last_slice = None
        if plan.needs_padding:
            tab = np.zeros((2, plan.slice_size), dtype=np.int32)
            for row, ln in enumerate((plan.bucket_base + 1, plan.bucket_base)):
                if (row == 0 and plan.bucket_rem == 0) or plan.rounded_per_bucket <= ln:
                    continue
                _lib.check(lib.stk_sampler_last_slice(C.byref(plan), ln, tab[row].ctypes.data))
            self._last_slice = tab
        self._dev = None
        if self.rank == info_rank:
            print(f"Stoke -- BucketedDistributedSampler -- # Samples Per Bucket: {self.rounded_num_samples_per_bucket}, "
                  f"# of Samples Per Replica: {self.rounded_num_samples_per_replica}")

    def _device_state(self):
        if self._dev is None:
            eng = get_engine()
            dev = torch.device("cuda", eng.device)
            si = self.sorted_n_samples
            si = si.to(device=dev, dtype=torch.int64) if isinstance(si, torch.Tensor) else \
                torch.as_tensor(np.asarray(si, dtype=np.int64)).to(dev)
            ls = torch.as_tensor(self._last_slice).to(dev) if self._last_slice is not None else None
            self._dev = (eng, dev, si.contiguous(), ls)
        return self._dev

    def indices_tensor(self) -> torch.Tensor:
        """This replica's indices for the current epoch as an int64 CUDA tensor."""
        plan = self._plan
        if plan.n_batches * plan.batch_size != plan.rounded_per_replica:
            # the reference's own ``assert len(final_indices) == self.rounded_num_samples_per_replica`` (data.py:447)
            raise AssertionError
        eng, dev, sorted_idx, last_slice = self._device_state()
        bucket_perm = batch_perm = None
        if self.shuffle:
            lens = (C.c_int64 * self.buckets)(*self._bucket_lens)
            host = torch.empty(plan.n + plan.n_batches, dtype=torch.int32).pin_memory()
            base = host.data_ptr()
            _lib.check(eng.lib.stk_randperm(self.seed + self.epoch, lens, self.buckets, base))
            nb = (C.c_int64 * 1)(plan.n_batches)
            _lib.check(eng.lib.stk_randperm(self.seed + self.epoch, nb, 1, base + 4 * plan.n))
            perms = host.to(dev, non_blocking=True)
            bucket_perm, batch_perm = perms[: plan.n], perms[plan.n:]
        out = torch.empty(plan.rounded_per_replica, dtype=torch.int64, device=dev)
        _lib.check(eng.lib.stk_sampler_indices(
            eng.ctx, C.byref(plan), sorted_idx.data_ptr(), bucket_perm.data_ptr() if bucket_perm is not None else None,
            batch_perm.data_ptr() if batch_perm is not None else None,
            last_slice.data_ptr() if last_slice is not None else None, out.data_ptr(), eng._stream()), eng.ctx)
        eng.launches += 1
        return out

    def __iter__(self) -> Iterator[int]:
        return iter(self.indices_tensor().tolist())

    def __len__(self) -> int:
        return self.rounded_num_samples_per_replica

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch
