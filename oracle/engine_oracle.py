# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the post-backward gradient path.

    [W-rank gradient reduce] -> [loss-scaler unscale] -> [grad-accumulation add] -> [grad clip] -> optimizer step
    -> [loss-scaler update]

This is a CPU, fp32 restatement of what the reference (fidelity/stoke) executes on that path.  stoke itself holds no
arithmetic here -- it routes to torch -- so the restatement *calls the same torch CPU ops in the reference's order*:

  * accumulation cadence / no_sync  ........ /root/reference/stoke/stoke.py:326-334, 960-988
                                              /root/reference/stoke/distributed.py:648-669 (DDP ``no_sync``)
  * DDP bucket reduce (pre-divide by W, SUM)  /root/reference/stoke/extensions.py:207-215 -> torch DDP reducer;
                                              python mirror torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:18-32
  * scaler.unscale_ before clipping ......... /root/reference/stoke/fp16.py:180-183, 222-225
  * clip_grad_norm_ / clip_grad_value_ ...... /root/reference/stoke/fp16.py:184, 233
  * scaler.step(optimizer); scaler.update() . /root/reference/stoke/fp16.py:805-806   (no scaler: optimizer.step() :298)
  * zero grads (set_to_none=True) ........... /root/reference/stoke/stoke.py:1042-1058, stoke/utils.py:103-106
  * OSS (fairscale, absent here) ............ final weights equal those of the unsharded optimizer; the oracle therefore
                                              runs the unsharded optimizer ("parity unpinned" for OSS clip)

Pinning: ``tests/test_oracle_vs_reference.py`` checks this engine bit-for-bit against the unmodified reference ``Stoke``
CPU run (world 1) when ``/root/reference`` is present, and ``tests/test_oracle_ddp_gloo.py`` checks the W-rank reduce
against real ``torch.nn.parallel.DistributedDataParallel`` on gloo.  There are no golden vectors in the reference itself
(it ships no tests) -- see DESIGN.md.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this.
"""
from typing import Dict, List, Optional, Sequence, Tuple, Type

import torch


class OracleEngine:
    """Holds ONE replica's fp32 weights (all DDP replicas are identical) and applies W logical ranks' gradients.

    Parameters
    ----------
    params: initial fp32 parameter values (cloned)
    world: number of logical ranks W
    optimizer / optimizer_kwargs: same meaning as ``StokeOptimizer`` (/root/reference/stoke/configs.py:754-770)
    grad_accum: micro-steps per optimizer step
    clip: None | ("norm", max_norm, norm_type) | ("value", clip_value)   (ClipGradNormConfig / ClipGradConfig)
    amp: None | dict(init_scale, growth_factor, backoff_factor, growth_interval)   (AMPConfig, stoke/configs.py:44-65)
    groups: None | [(parameter indices, per-group optimizer kwargs), ...]   (torch parameter groups)

    A gradient entry may be ``None``: the parameter received no gradient on that rank in that micro-step (an unused
    parameter); a parameter whose accumulated gradient is ``None`` is skipped by the optimizer, as torch does after
    ``zero_grad(set_to_none=True)`` (stoke/utils.py:103-106).
    """

    def __init__(
        self,
        params: Sequence[torch.Tensor],
        world: int,
        optimizer: Type[torch.optim.Optimizer],
        optimizer_kwargs: Dict,
        grad_accum: int = 1,
        clip: Optional[Tuple] = None,
        amp: Optional[Dict] = None,
        groups: Optional[Sequence[Tuple[Sequence[int], Dict]]] = None,
    ):
        self.world = int(world)
        self.grad_accum = int(grad_accum)
        self.clip = clip
        self.params = [torch.nn.Parameter(p.detach().to(torch.float32).cpu().clone()) for p in params]
        if groups is None:
            self.optimizer = optimizer(params=self.params, **optimizer_kwargs)
        else:
            # torch-style parameter groups: (indices into params, per-group overrides)
            self.optimizer = optimizer([dict(extra, params=[self.params[i] for i in idx]) for idx, extra in groups],
                                       **optimizer_kwargs)
        self.scaler = None
        if amp is not None:
            # the reference builds torch.cuda.amp.GradScaler (stoke/fp16.py:733-748); the CPU build of the same class
            self.scaler = torch.amp.GradScaler(
                "cpu",
                init_scale=amp.get("init_scale", 2.0**16),
                growth_factor=amp.get("growth_factor", 2.0),
                backoff_factor=amp.get("backoff_factor", 0.5),
                growth_interval=amp.get("growth_interval", 2000),
                enabled=True,
            )
            # ``scaler.scale(loss)`` in backward_call (stoke/fp16.py:782-786) is what lazily creates the scale tensor
            self.scaler.scale(torch.zeros(1))
        # per-rank local accumulators (what ``param.grad`` holds on each rank under ``no_sync``)
        self._local: List[Optional[List[torch.Tensor]]] = [None] * self.world
        self._micro = 0
        self.optimizer_steps = 0
        self.skipped_steps = 0
        self.last_total_norm = None

    # ------------------------------------------------------------------------------------------------------------
    @property
    def loss_scale(self) -> float:
        return 1.0 if self.scaler is None else float(self.scaler.get_scale())

    def micro_step(self, grads_per_rank: Sequence[Sequence[torch.Tensor]]):
        """One backward on every rank.  ``grads_per_rank[r][i]`` is rank r's gradient of parameter i for this
        micro-batch *as produced by autograd*, i.e. already multiplied by the loss scale (if amp) and by
        1/grad_accum (stoke/stoke.py:910-911).  AccumulateGrad adds it to ``param.grad`` in fp32."""
        assert len(grads_per_rank) == self.world
        for r in range(self.world):
            g = [None if x is None else x.detach().to(torch.float32).cpu() for x in grads_per_rank[r]]
            if self._local[r] is None:
                self._local[r] = [None if x is None else x.clone() for x in g]
            else:
                for i, x in enumerate(g):
                    if x is None:
                        continue
                    if self._local[r][i] is None:
                        self._local[r][i] = x.clone()
                    else:
                        self._local[r][i].add_(x)
        self._micro += 1

    def ready(self) -> bool:
        return self._micro == self.grad_accum

    def step(self) -> bool:
        """The sync step: DDP reduce of the accumulated grads, then stoke's clip -> step -> reset order
        (stoke/stoke.py:990-1040).  Returns True if the optimizer stepped (False: skipped on inf/nan)."""
        assert self.ready(), "step() before grad_accum micro-steps"
        W = self.world
        # DDP: bucket = grad / W on every rank, then SUM over ranks (rank order)
        for i, p in enumerate(self.params):
            red = None
            for r in range(W):
                g = self._local[r][i]
                if g is None:
                    continue
                g = g / W if W > 1 else g
                red = g.clone() if red is None else red.add_(g)
            p.grad = None if red is None else red.reshape(p.shape)
        stepped = True
        if self.clip is not None:
            if self.scaler is not None:
                self.scaler.unscale_(self.optimizer)
            if self.clip[0] == "norm":
                self.last_total_norm = torch.nn.utils.clip_grad_norm_(
                    [p for p in self.params if p.grad is not None], max_norm=self.clip[1], norm_type=self.clip[2]
                )
            elif self.clip[0] == "value":
                torch.nn.utils.clip_grad_value_([p for p in self.params if p.grad is not None], clip_value=self.clip[1])
            else:
                raise ValueError(self.clip)
        if self.scaler is not None:
            calls = []
            inner = self.optimizer.step
            self.optimizer.step = lambda *a, **k: (calls.append(1), inner(*a, **k))[1]
            try:
                self.scaler.step(self.optimizer)  # skips ``optimizer.step`` when any grad is inf/nan
            finally:
                self.optimizer.step = inner
            self.scaler.update()
            stepped = bool(calls)
        else:
            self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        self._local = [None] * W
        self._micro = 0
        if stepped:
            self.optimizer_steps += 1
        else:
            self.skipped_steps += 1
        return stepped

    # ------------------------------------------------------------------------------------------------------------
    def weights(self) -> List[torch.Tensor]:
        return [p.detach().clone() for p in self.params]

    def flat_weights(self) -> torch.Tensor:
        return torch.cat([p.detach().reshape(-1) for p in self.params])

    def flat_state(self, key: str) -> torch.Tensor:
        out = []
        for p in self.params:
            st = self.optimizer.state.get(p, {})
            out.append(st[key].reshape(-1) if key in st else torch.zeros(p.numel()))
        return torch.cat(out)
