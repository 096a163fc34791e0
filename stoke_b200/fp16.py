# -*- coding: utf-8 -*-
"""Precision mixins of the runner -- same method set the reference's ``BaseFP16`` family exposes
(/root/reference/stoke/fp16.py:59-298, 694-806: ``wrap_fp16``, ``clip_grad``, ``backward_call``, ``step_call``, ``scaler``,
``model_context``, ``loss_context``), re-targeted at the engine:

  * ``backward_call`` runs autograd into the flat gradient bucket, then launches K1 (or the local accumulate kernel under
    the no-sync context); the scaler's unscale and inf test are fused into K1 instead of ``scaler.unscale_``.
  * ``clip_grad`` only records that this step clips: the norm was reduced inside K1 and the coefficient is applied in
    registers by K2 (no pass over the gradients, no host sync).
  * ``step_call`` = ``optimizer.step()`` = K2 + the one-thread epilogue that does ``scaler.update()`` on the device.
"""
from contextlib import nullcontext
from enum import Enum
from typing import Optional

import torch

from . import _lib
from .configs import ClipGradConfig, ClipGradNormConfig


class DeviceGradScaler:
    """``torch.cuda.amp.GradScaler``-shaped handle on the scaler state that lives on the device inside the engine
    (state_dict keys follow torch/amp/grad_scaler.py: scale, growth_factor, backoff_factor, growth_interval,
    _growth_tracker)."""

    def __init__(self, engine, init_scale=2.0**16, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000,
                 state_id: int = 0):
        self._engine = engine
        self._state = state_id   # the per-optimizer device state this scaler lives in (engine.state_create)
        engine.scaler_set(state=state_id, scale=float(init_scale), growth_factor=float(growth_factor),
                          backoff_factor=float(backoff_factor), growth_interval=int(growth_interval), growth_tracker=0,
                          enabled=1, found_inf=0)
        self._scale = engine.scale_tensor(state_id)

    def scale(self, outputs):
        if isinstance(outputs, (list, tuple)):
            return type(outputs)(self.scale(o) for o in outputs)
        # the scale stays fp32 (torch GradScaler does the same): an fp16 loss times 65536 must not overflow in fp16 --
        # type promotion of two 0-dim tensors gives an fp32 product
        return outputs * self._scale

    def get_scale(self) -> float:
        return float(self._engine.scaler_get(self._state).scale)

    def is_enabled(self) -> bool:
        return True

    def unscale_(self, optimizer):
        """Fused into the gradient reduce; kept so user code calling it does not break."""
        return None

    def step(self, optimizer, *args, **kwargs):
        return optimizer.step(*args, **kwargs)

    def update(self, new_scale=None):
        if new_scale is not None:
            self._engine.scaler_set(state=self._state, scale=float(new_scale))

    def state_dict(self):
        st = self._engine.scaler_get(self._state)
        return {"scale": st.scale, "growth_factor": st.growth_factor, "backoff_factor": st.backoff_factor,
                "growth_interval": st.growth_interval, "_growth_tracker": st.growth_tracker}

    def load_state_dict(self, sd):
        self._engine.scaler_set(state=self._state, scale=float(sd["scale"]), growth_factor=float(sd["growth_factor"]),
                                backoff_factor=float(sd["backoff_factor"]), growth_interval=int(sd["growth_interval"]),
                                growth_tracker=int(sd["_growth_tracker"]))


class BaseFP16:
    """Full precision (reference: NullFP16): fp32 model, fp32 gradients, master weights are the model's own."""

    _lp_dtype: Optional[torch.dtype] = None
    _autocast_dtype: Optional[torch.dtype] = None

    def __init__(self, verbose: bool = True, **kwargs):
        self._scaler = None
        self._verbose = verbose
        self._clip_requested = False

    def _scaler_info(self):
        if self._verbose and self._scaler is not None:
            self._print_device(f"FP16 Mixin: Initialized scaler of type {type(self._scaler).__name__}")

    def wrap_fp16(self, model, optimizer=None):
        # no loss scaling in this mode: this runner's own device state (engine.state_create) starts with the scaler
        # disabled and scale 1, so nothing has to be reset here -- other Stoke objects of the process have their own state
        self._scaler_info()
        return model, optimizer

    def clip_grad(self, grad_clip, model, optimizer, oss: bool, horovod: bool, deepspeed: bool, fsdp: bool):
        if not isinstance(grad_clip, (ClipGradConfig, ClipGradNormConfig)):
            raise ValueError(f"Stoke -- clip_grad received an incorrect instance type of {type(grad_clip)}")
        if self._verbose:
            self._print_device(f'{type(grad_clip).__name__.replace("Config", "")} is automatically clipping '
                               f"calculated/accumulated gradients...")
        # norm already reduced by K1 with the configured norm type; K2 applies the coefficient / clamp
        self._clip_requested = True

    @property
    def scaler(self):
        return self._scaler

    @property
    def loss_context(self):
        return nullcontext() if self._autocast_dtype is None else torch.autocast("cuda", dtype=self._autocast_dtype)

    @property
    def model_context(self):
        return nullcontext() if self._autocast_dtype is None else torch.autocast("cuda", dtype=self._autocast_dtype)

    def backward_call(self, loss, model, optimizer):
        path = optimizer.path
        sync = not getattr(self, "_defer_sync", False)
        unscale = self._scaler is not None
        scaled = loss if self._scaler is None else self._scaler.scale(loss)
        if isinstance(scaled, (list, tuple)):
            for idx, val in enumerate(scaled):
                # the hooks may launch per-bucket reduces only during the LAST backward of the micro-step
                path.begin_backward(sync=sync, unscale=unscale, last=(idx == len(scaled) - 1))
                val.backward(retain_graph=(idx == 0))
        else:
            path.begin_backward(sync=sync, unscale=unscale)
            scaled.backward()
        path.after_backward(sync=sync, unscale=unscale)

    def step_call(self, model, optimizer):
        optimizer.step()
        self._clip_requested = False


class NullFP16(BaseFP16):
    pass


class B200AmpFP16(BaseFP16):
    """``fp16="amp"``: fp16 autocast over an fp32 model + dynamic loss scaling (reference NativeAmpFP16, fp16.py:694-806);
    gradients are fp32, the scaler state lives on the device."""

    _autocast_dtype = torch.float16

    def __init__(self, verbose: bool = True, **kwargs):
        super().__init__(verbose=verbose)
        self._amp_config = kwargs["amp_config"]

    def wrap_fp16(self, model, optimizer=None):
        cfg = self._amp_config
        self._scaler = DeviceGradScaler(self._engine, init_scale=cfg.init_scale, growth_factor=cfg.growth_factor,
                                        backoff_factor=cfg.backoff_factor, growth_interval=cfg.growth_interval,
                                        state_id=self._state_id)
        self._scaler_info()
        return model, optimizer


class B200BF16(BaseFP16):
    """``fp16="bf16"`` (extension): bf16 model and gradients, fp32 master weights and optimizer state, no loss scaling.
    Closest reference semantics: Apex O2 (fp16.py:582-635)."""

    _lp_dtype = torch.bfloat16
    _autocast_dtype = torch.bfloat16


class RunnerFP16Enum(Enum):
    full = NullFP16
    amp = B200AmpFP16
    bf16 = B200BF16
