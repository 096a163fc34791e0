# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- CPU port of the reference's training-step call order (world size 1).

The reference (``/root/reference/stoke``) is pure Python and cannot travel to the GPU box, so this port is what
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs time there (``cpu_baseline.kind == "port"``).  It restates,
for the runner the reference builds with ``gpu=False`` (``DistributedNullCPU + NullFP16 + BaseOptimizer``, see
/root/reference/stoke/stoke.py:659-735), exactly which torch calls happen and in which order:

  model()     -> plain forward ........................................ stoke/stoke.py:853-869
  loss()      -> loss fn; ``.item()`` bookkeeping; ``/ grad_accum`` only in training mode  stoke/stoke.py:872-912,
                                                                        stoke/distributed.py:136-160 (base detach)
  backward()  -> counter += 1; ``loss.backward()`` ..................... stoke/stoke.py:960-988, stoke/fp16.py:252-278
  step()      -> only when ``(counter + 1) % (grad_accum + 1) == 0``: clip_grad_norm_/clip_grad_value_, optimizer.step(),
                 zero_grad(set_to_none=True), counter = 0 ............. stoke/stoke.py:326-334, 990-1058,
                                                                        stoke/fp16.py:158-235, 280-298, stoke/utils.py:83-106

``tests/test_oracle_vs_reference.py`` runs this port and the unmodified reference side by side (when the reference tree is
present) and requires bit-identical weights, losses and counters.
"""
from typing import Callable, Dict, Optional, Tuple, Type

import torch


class StokePortCPU:
    def __init__(
        self,
        model: torch.nn.Module,
        optimizer: Type[torch.optim.Optimizer],
        optimizer_kwargs: Dict,
        loss: Callable,
        grad_accum_steps: int = 1,
        clip: Optional[Tuple] = None,
        ema_weight: float = 0.1,
    ):
        self.net = model
        self.loss_fn = loss
        self.accum = grad_accum_steps if grad_accum_steps is not None else 1
        self.clip = clip
        self.opt = optimizer(params=model.parameters(), **optimizer_kwargs)
        self.counter = 0
        self.backward_steps = 0
        self.optimizer_steps = 0
        self.step_loss = 0.0
        self.agg_loss = 0.0
        self.ema_loss = 0.0
        self._ema_n = 0
        self._ema_w = ema_weight

    def _sync_step(self) -> bool:
        return (self.counter + 1) % (self.accum + 1) == 0

    def model(self, *a, **k):
        return self.net(*a, **k)

    def loss(self, *a, **k):
        val = self.loss_fn(*a, **k)
        host = val.item()
        self.step_loss = host
        self.agg_loss += host
        self._ema_n += 1
        self.ema_loss = host if self._ema_n == 1 else self._ema_w * host + (1.0 - self._ema_w) * self.ema_loss
        if self.accum > 1 and self.net.training:
            val = val / self.accum
        return val

    def backward(self, loss: torch.Tensor):
        self.counter += 1
        loss.backward()
        self.backward_steps += 1

    def step(self):
        if not self._sync_step():
            return
        if self.clip is not None:
            if self.clip[0] == "norm":
                torch.nn.utils.clip_grad_norm_(self.net.parameters(), max_norm=self.clip[1], norm_type=self.clip[2])
            else:
                torch.nn.utils.clip_grad_value_(self.net.parameters(), clip_value=self.clip[1])
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        self.counter = 0
        self.agg_loss = 0.0
        self.optimizer_steps += 1
