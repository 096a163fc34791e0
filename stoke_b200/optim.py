# -*- coding: utf-8 -*-
"""The optimizer object ``Stoke.optimizer`` hands back: a ``torch.optim.Optimizer`` (so LR schedulers and
``param_groups`` edits work, README.md:241-250 of the reference) whose ``step()`` is the single fused K2 launch over the
flat fp32 master / moment buffers, and whose sharded flavour is the ZeRO-1 analogue of fairscale ``OSS``
(/root/reference/stoke/extensions.py:109-141): same ``clip_grad_norm`` / ``consolidate_state_dict`` surface.
"""
from typing import Dict, Optional, Type

import torch

from . import _lib
from .engine import ClipSpec, Engine, GradPath

_KINDS = {torch.optim.Adam: _lib.OPT_ADAM, torch.optim.AdamW: _lib.OPT_ADAMW, torch.optim.SGD: _lib.OPT_SGD}
_REJECT_TRUE = ("amsgrad", "capturable", "differentiable")


class B200FusedOptimizer(torch.optim.Optimizer):
    def __init__(self, module: torch.nn.Module, optim_cls: Type[torch.optim.Optimizer], optim_kwargs: Dict,
                 engine: Engine, grad_accum: int = 1, clip: Optional[ClipSpec] = None, sharded: bool = False,
                 lp_dtype: Optional[torch.dtype] = None):
        if optim_cls not in _KINDS:
            raise NotImplementedError(
                f"Stoke -- stoke_b200 fuses torch.optim.Adam, AdamW and SGD; got {getattr(optim_cls, '__name__', optim_cls)}")
        # let torch validate the kwargs exactly as the reference would (optimizer(params=..., **kwargs))
        probe = optim_cls([torch.nn.Parameter(torch.zeros(1))], **optim_kwargs)
        defaults = dict(probe.defaults)
        for k in _REJECT_TRUE:
            if defaults.get(k):
                raise NotImplementedError(f"Stoke -- optimizer option {k}=True is not supported by the fused step")
        self._kind = _KINDS[optim_cls]
        if defaults.get("decoupled_weight_decay"):  # torch >= 2.6: AdamW is Adam(decoupled_weight_decay=True)
            self._kind = _lib.OPT_ADAMW
        self._torch_cls = optim_cls
        params = [p for p in module.parameters() if p.requires_grad]
        sgd = self._kind == _lib.OPT_SGD
        self.path = GradPath(engine, params, grad_accum=grad_accum, clip=clip, sharded=sharded, lp_dtype=lp_dtype,
                             module=module, needs_second_moment=not sgd,
                             needs_first_moment=(not sgd) or defaults.get("momentum", 0) != 0)
        super().__init__(params, defaults)
        # the step counters live in the engine's device state (one live optimizer per engine/process)
        engine.scaler_set(opt_steps=0, skipped_steps=0, found_inf=0, growth_tracker=0)
        if len(self.param_groups) != 1:
            raise NotImplementedError("Stoke -- one parameter group (the reference passes model.parameters())")

    # -- the step -------------------------------------------------------------------------------------------------------
    def _hyper(self) -> _lib.OptimHyper:
        g = self.param_groups[0]
        h = _lib.OptimHyper()
        h.kind = self._kind
        h.lr = float(g["lr"])
        h.weight_decay = float(g.get("weight_decay", 0.0))
        h.maximize = int(bool(g.get("maximize", False)))
        if self._kind == _lib.OPT_SGD:
            h.momentum = float(g.get("momentum", 0.0))
            h.dampening = float(g.get("dampening", 0.0))
            h.nesterov = int(bool(g.get("nesterov", False)))
        else:
            h.beta1, h.beta2 = (float(b) for b in g["betas"])
            h.eps = float(g["eps"])
        return h

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("Stoke -- closures are not supported by the fused step")
        self.path.optimizer_step(self._hyper())

    def zero_grad(self, set_to_none: bool = True):
        """No-op: the gradient bucket is zeroed by the reduce kernel once it has been consumed, and ``param.grad`` must
        stay a view of that bucket (reference: zero_optimizer_grads, stoke/utils.py:83-106)."""
        return None

    # -- fairscale OSS surface (stoke/fp16.py:227-228, stoke/io_ops.py:596-600) ------------------------------------------
    def clip_grad_norm(self, max_norm: float, norm_type: float = 2.0):
        self.path.clip = ClipSpec(_lib.CLIP_NORM, max_norm=max_norm, norm_type=norm_type)

    def consolidate_state_dict(self, recipient_rank: int = 0):
        """fairscale OSS gathers the shards on ``recipient_rank`` before ``state_dict()``; here ``state_dict()`` itself
        all-gathers the sharded buffers on every rank, so this only exists for call-compatibility."""
        return None

    # -- state dict in torch's per-parameter format ----------------------------------------------------------------------
    def _full(self, flat: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if flat is None:
            return None
        return self.path._gather_shards(flat) if self.path.sharded else flat

    def state_dict(self):
        path = self.path
        st = path.engine.scaler_get()
        m, v, master = self._full(path.m_flat), self._full(path.v_flat), self._full(path.master_flat)
        state = {}
        ms = path.unflatten(m) if m is not None else None
        vs = path.unflatten(v) if v is not None else None
        for i in range(len(path.params)):
            if self._kind == _lib.OPT_SGD:
                # torch creates the buffer on the first step; before that (or without momentum) it is None
                state[i] = {"momentum_buffer": ms[i].clone() if ms is not None and st.opt_steps > 0 else None}
            else:
                state[i] = {"step": torch.tensor(float(st.opt_steps)), "exp_avg": ms[i].clone(),
                            "exp_avg_sq": vs[i].clone()}
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        groups[0]["params"] = list(range(len(path.params)))
        return {"state": state, "param_groups": groups, "b200_master": master.clone(), "b200_opt_steps": int(st.opt_steps)}

    def load_state_dict(self, sd):
        path = self.path
        sb, se = path.shard
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
        steps = int(sd.get("b200_opt_steps", 0))

        def put(flat_dst, per_param):
            if flat_dst is None:
                return
            full = torch.zeros(path.n, dtype=torch.float32, device=flat_dst.device)
            for dst, src in zip(path.unflatten(full), per_param):
                if src is not None:
                    dst.copy_(src.to(device=full.device, dtype=torch.float32))
            flat_dst.copy_(full[sb:se])

        n = len(path.params)
        state = sd["state"]
        if self._kind == _lib.OPT_SGD:
            put(path.m_flat, [state.get(i, {}).get("momentum_buffer") for i in range(n)])
        else:
            put(path.m_flat, [state[i]["exp_avg"] for i in range(n)])
            put(path.v_flat, [state[i]["exp_avg_sq"] for i in range(n)])
            if "b200_opt_steps" not in sd and n:
                steps = int(float(state[0]["step"]))
        if "b200_master" in sd:
            master = sd["b200_master"].to(device=path.p_flat.device, dtype=torch.float32)
            path.master_flat.copy_(master[sb:se])
            if path.low_precision:
                path.p_flat.copy_(master)
            elif path.sharded:
                path.p_flat.copy_(master)
        path.engine.scaler_set(opt_steps=steps)
