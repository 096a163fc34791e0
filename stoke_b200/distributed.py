# -*- coding: utf-8 -*-
"""Distributed mixins of the runner -- the method set of /root/reference/stoke/distributed.py:94-295
(``setup_distributed``, ``wrap_distributed``, ``detach_and_sync_loss``, ``grad_accum_context``, ``step_context``,
``barrier``, ``clean``, ``print_device``, ``rank``/``world_size``/``initialized``/``device_id``).

``DistributedB200GPU`` replaces ``DistributedNullGPU`` (:351) and ``DistributedB200DDP`` replaces ``DistributedDDP`` (:404):
torch.distributed/NCCL is used for bring-up only (rank discovery, IPC-handle exchange, initial parameter broadcast);
the per-step loss mean, barrier, buffer broadcast and gradient reduction run in the library's kernels over peer memory.
"""
from contextlib import contextmanager, nullcontext
from enum import Enum
from typing import List, Optional, Union

import torch

from .engine import get_engine
from .extensions import DistributedHandlerEnum
from .utils import unrolled_print


class BaseDistributed:
    def __init__(self, device_id, batch_size_per_device: int, info_rank, name: str, verbose: bool = True):
        self._batch_size_per_device = batch_size_per_device
        self._device_id = device_id
        self._info_rank = info_rank
        self._name = name
        self._verbose = verbose
        self._engine = None
        self._state_id = None     # this runner's device state (scaler + step counters): one per Stoke object
        self._defer_sync = False

    def _print_info(self):
        self._print_device(f"{self._name} Initialized: {self.initialized}")

    def setup_distributed(self):
        self._engine = get_engine(device=torch.cuda.current_device(), rank=0, world=1)
        self._state_id = self._engine.state_create()

    def sync_loss_begin(self, loss):
        """Launches the cross-rank loss mean and returns ticket(s); nothing synchronises until the value is read."""
        if isinstance(loss, (list, tuple)):
            return type(loss)(self._engine.loss_sync_begin(val.detach()) for val in loss)
        return self._engine.loss_sync_begin(loss.detach())

    def sync_loss_wait(self, ticket):
        if isinstance(ticket, (list, tuple)):
            return type(ticket)(self._engine.loss_sync_wait(t) for t in ticket)
        return self._engine.loss_sync_wait(ticket)

    def wrap_distributed(self, model, grad_accum: Optional[int], optimizer=None):
        if self._verbose:
            self._print_info()
        return model, optimizer

    def detach_and_sync_loss(self, loss, device=None):
        if isinstance(loss, (list, tuple)):
            return type(loss)(self._engine.loss_sync(val.detach()) for val in loss)
        return self._engine.loss_sync(loss.detach())

    @contextmanager
    def _deferred(self, model):
        # local accumulation instead of a cross-rank reduce for this backward (DDP.no_sync semantics,
        # reference distributed.py:648-669); the fp16 mixin's backward_call reads the flag
        self._defer_sync = True
        inner = model.no_sync() if hasattr(model, "no_sync") else nullcontext()
        try:
            with inner:
                yield
        finally:
            self._defer_sync = False

    def grad_accum_context(self, model):
        return self._deferred(model)

    def step_context(self, optimizer):
        return nullcontext()

    def clean(self):
        pass

    def _print_device(self, msg: Union[str, List[str]]):
        self.print_device(msg=msg, rank=self._info_rank)

    def print_device(self, msg, rank: Optional[Union[int, List[int]]] = 0, single_line: bool = False):
        if self.rank in ("cpu", "gpu"):
            unrolled_print(msg, single_line=single_line)
        elif isinstance(rank, list) and self.rank in rank:
            unrolled_print(msg, single_line=single_line)
        elif isinstance(rank, int) and rank == self.rank:
            unrolled_print(msg, single_line=single_line)

    def barrier(self):
        pass

    @property
    def device_id(self):
        return self._device_id

    @property
    def engine(self):
        return self._engine


class DistributedB200GPU(BaseDistributed):
    """Single B200, no process group (reference DistributedNullGPU)."""

    def __init__(self, batch_size_per_device: int, info_rank, verbose: bool = True, **kwargs):
        super().__init__(device_id=torch.cuda.current_device(), batch_size_per_device=batch_size_per_device,
                         info_rank=info_rank, name="B200 GPU", verbose=verbose)

    rank = property(lambda self: "gpu")
    world_size = property(lambda self: 1)
    initialized = property(lambda self: True)


class DistributedB200DDP(BaseDistributed):
    """One process per GPU of one NVSwitch node."""

    def __init__(self, batch_size_per_device: int, info_rank, verbose: bool = True, **kwargs):
        self._ddp_config = kwargs["ddp_config"]
        super().__init__(device_id=self._ddp_config.local_rank, batch_size_per_device=batch_size_per_device,
                         info_rank=info_rank, name="B200 DDP", verbose=verbose)
        handler = DistributedHandlerEnum.sddp if kwargs.get("sharded_config") is not None else DistributedHandlerEnum.base
        self._ddp_handler = handler.value(verbose=verbose, sddp_config=kwargs.get("sharded_config"),
                                          ddp_config=self._ddp_config)

    def setup_distributed(self):
        torch.cuda.set_device(self._device_id)
        if not torch.distributed.is_initialized():
            backend = self._ddp_config.backend
            backend = backend.value if isinstance(backend, Enum) else backend
            torch.distributed.init_process_group(backend=backend.strip(), init_method=self._ddp_config.init_method)
        self._engine = get_engine(device=self._device_id, rank=torch.distributed.get_rank(),
                                  world=torch.distributed.get_world_size())
        self._state_id = self._engine.state_create()
        if self._ddp_config.no_sync is False:
            import warnings

            warnings.warn("Stoke -- DDPConfig.no_sync=False: the reference then all-reduces after every micro-step "
                          "(stoke/distributed.py:666-668); this engine accumulates locally and reduces once per optimizer "
                          "step -- the same sum with 1/grad_accum of the bytes")

    def wrap_distributed(self, model, grad_accum: Optional[int], optimizer=None):
        if self._verbose:
            self._print_device(f"{self._name} Class: {type(self._ddp_handler).__name__}")
            self._print_info()
            self._print_device([f"{self._name} -- Device ID: {torch.cuda.current_device()}",
                                f"{self._name} -- Rank: {self.rank}"])
        if self._ddp_config.convert_to_sync_batch_norm:
            self.print_device("Converting all BatchNorm*D layers to torch.nn.SyncBatchNorm layers...")
            model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(module=model)
        return self._ddp_handler.handle_ddp(model=model, optimizer=optimizer, grad_accum=grad_accum, rank=self.rank,
                                            engine=self._engine)

    def barrier(self):
        self._engine.barrier()

    def clean(self):
        torch.distributed.destroy_process_group()

    rank = property(lambda self: torch.distributed.get_rank())
    world_size = property(lambda self: torch.distributed.get_world_size())
    initialized = property(lambda self: torch.distributed.is_initialized())


class RunnerDistEnum(Enum):
    gpu = DistributedB200GPU
    ddp = DistributedB200DDP
