# -*- coding: utf-8 -*-
"""Host side of the B200 gradient-path engine: one ``Engine`` per process (one process per GPU) owning the C-ABI
context, the peer-visible flat buffers and the launch bookkeeping, and one ``GradPath`` per model that lays the model's
parameters / gradients / optimizer state out in those flat buffers and drives K1 (reduce) and K2 (optimizer step).

torch is used here for what it is good at -- tensors as typed views of device memory, streams, autograd, and
``torch.distributed`` for process-group bring-up (rank/world discovery, the one-time exchange of IPC handles and the
initial parameter broadcast).  Every byte of the per-step gradient path moves through the hand-written kernels in
``csrc/``; if the library is missing or a call fails this module raises -- there is no torch fallback.
"""
import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import StokeB200Error, check

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
_ESZ = {torch.float32: 4, torch.bfloat16: 2, torch.float16: 2}
ALIGN_ELEMS = 8  # every parameter starts on an 8-element boundary (16 B of bf16 / 32 B of fp32)


class DeviceBuffer:
    """A cudaMalloc allocation owned by the library, mapped into every peer (CUDA IPC) when world > 1."""

    def __init__(self, engine: "Engine", ptr: int, nbytes: int, peers: List[int]):
        self.engine, self.ptr, self.nbytes, self.peers = engine, ptr, nbytes, peers
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None,
        }

    def tensor(self, dtype: torch.dtype, numel: Optional[int] = None, offset_elems: int = 0) -> torch.Tensor:
        raw = torch.as_tensor(self, device=torch.device("cuda", self.engine.device))
        t = raw.view(dtype)
        if numel is None:
            numel = t.numel() - offset_elems
        return t[offset_elems: offset_elems + numel]

    def peer_ptrs(self, offset_bytes: int = 0) -> List[int]:
        return [p + offset_bytes for p in self.peers]


class Engine:
    """Process-wide handle on libstoke_b200 (context + signal pads)."""

    def __init__(self, device: int, rank: int = 0, world: int = 1, group=None):
        if not torch.cuda.is_available():
            raise StokeB200Error(_lib.ERR_CUDA, "no CUDA device: stoke_b200 has no CPU path")
        self.lib = _lib.load()
        self.device, self.rank, self.world, self.group = int(device), int(rank), int(world), group
        self._buffers: List[DeviceBuffer] = []
        self.launches = 0  # kernels launched through this engine (bench.py reports it as gpu_launches)
        ctx = C.c_void_p()
        check(self.lib.stk_ctx_create(self.rank, self.world, self.device, 0, C.byref(ctx)))
        self.ctx = ctx
        handle = C.create_string_buffer(_lib.STK_IPC_HANDLE_BYTES)
        self._check(self.lib.stk_comm_local(self.ctx, handle))
        handles = self._exchange(handle.raw)
        self._check(self.lib.stk_comm_connect(self.ctx, handles))
        if self.world > 1:
            torch.distributed.barrier(group=self.group)  # every pad is zeroed and mapped before the first kernel
        caps = _lib.Caps()
        self._check(self.lib.stk_caps(self.ctx, C.byref(caps)))
        self.sm_count = caps.sm_count

    # -- plumbing ---------------------------------------------------------------------------------------------------
    def _check(self, code):
        check(code, self.ctx)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _exchange(self, blob: bytes) -> bytes:
        if self.world == 1:
            return blob
        out = [None] * self.world
        torch.distributed.all_gather_object(out, blob, group=self.group)
        return b"".join(out)

    def alloc(self, nbytes: int) -> DeviceBuffer:
        nbytes = max(int(nbytes), 256)
        ptr = C.c_void_p()
        handle = C.create_string_buffer(_lib.STK_IPC_HANDLE_BYTES)
        self._check(self.lib.stk_mem_alloc_shared(self.ctx, nbytes, C.byref(ptr), handle))
        handles = self._exchange(handle.raw)
        peers = (C.c_void_p * _lib.STK_MAX_WORLD)()
        self._check(self.lib.stk_mem_open_peers(self.ctx, ptr, handles, peers))
        buf = DeviceBuffer(self, ptr.value, nbytes, [peers[r] for r in range(self.world)])
        self._buffers.append(buf)
        return buf

    def shard_range(self, n: int, rank: Optional[int] = None) -> Tuple[int, int]:
        b, e = C.c_size_t(), C.c_size_t()
        check(self.lib.stk_shard_range(n, self.world, self.rank if rank is None else rank, C.byref(b), C.byref(e)))
        return b.value, e.value

    def close(self):
        if getattr(self, "ctx", None):
            torch.cuda.synchronize(self.device)
            self.lib.stk_ctx_destroy(self.ctx)
            self.ctx = None

    # -- scaler / step state ----------------------------------------------------------------------------------------
    def scaler_get(self) -> _lib.ScalerState:
        st = _lib.ScalerState()
        self._check(self.lib.stk_scaler_get(self.ctx, C.byref(st), self._stream()))
        return st

    def scaler_set(self, **fields):
        st = self.scaler_get()
        for k, v in fields.items():
            setattr(st, k, v)
        self._check(self.lib.stk_scaler_set(self.ctx, C.byref(st), self._stream()))

    def scale_tensor(self) -> torch.Tensor:
        """0-dim float32 view of the live loss scale on the device (``scaler.scale(loss)`` multiplies by it)."""
        ptr = self.lib.stk_scaler_scale_ptr(self.ctx)

        class _V:
            pass

        v = _V()
        v.__cuda_array_interface__ = {"shape": (1,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
        v._keep = self
        return torch.as_tensor(v, device=torch.device("cuda", self.device))[0]

    # -- kernels ----------------------------------------------------------------------------------------------------
    def grad_accumulate(self, grad_ptr: int, dtype: torch.dtype, acc_ptr: int, n: int, first: bool, zero_grad: bool):
        self._check(self.lib.stk_grad_accumulate(self.ctx, grad_ptr, _DT[dtype], acc_ptr, n, int(first), int(zero_grad),
                                                 self._stream()))
        self.launches += 1

    def grad_reduce(self, mode: int, grad_ptrs: Sequence[int], grad_dtype: torch.dtype,
                    acc_ptrs: Optional[Sequence[int]], out_ptrs: Sequence[int], out_dtype: torch.dtype, n: int,
                    mul: float, norm_kind: int, norm_p: float, flags: int):
        acc = _lib.ptr_array(acc_ptrs) if acc_ptrs is not None else None
        self._check(self.lib.stk_grad_reduce(self.ctx, mode, _lib.ptr_array(grad_ptrs), _DT[grad_dtype], acc,
                                             _lib.ptr_array(out_ptrs), _DT[out_dtype], n, mul, norm_kind, norm_p, flags,
                                             self._stream()))
        self.launches += 1

    def optim_step(self, hyper: _lib.OptimHyper, master_ptr: int, m_ptr: Optional[int], v_ptr: Optional[int],
                   grad_ptr: int, n_local: int, lp_ptrs: Optional[Sequence[int]], lp_dtype: torch.dtype,
                   lp_offset: int):
        lp = _lib.ptr_array(lp_ptrs) if lp_ptrs is not None else None
        self._check(self.lib.stk_optim_step(self.ctx, C.byref(hyper), master_ptr, m_ptr, v_ptr, grad_ptr, n_local, lp,
                                            len(lp_ptrs) if lp_ptrs is not None else 0, _DT[lp_dtype], lp_offset,
                                            self._stream()))
        self.launches += 1

    def step_epilogue(self):
        self._check(self.lib.stk_step_epilogue(self.ctx, self._stream()))
        self.launches += 1

    def loss_sync(self, loss: torch.Tensor) -> float:
        """Mean over ranks of a scalar device tensor, as a python float (one stream synchronise)."""
        if loss.dtype not in _DT:
            loss = loss.float()
        out = C.c_double()
        self._check(self.lib.stk_loss_sync(self.ctx, loss.data_ptr(), _DT[loss.dtype], C.byref(out), self._stream()))
        self.launches += 1
        return out.value

    def barrier(self):
        self._check(self.lib.stk_barrier(self.ctx, self._stream()))
        self.launches += 1 if self.world > 1 else 0

    def bcast(self, buf: DeviceBuffer, nbytes: int, root: int = 0, offset_bytes: int = 0):
        if self.world == 1:
            return
        self._check(self.lib.stk_bcast(self.ctx, _lib.ptr_array(buf.peer_ptrs(offset_bytes)), nbytes, root, self._stream()))
        self.launches += 1

    def set_k1_algo(self, algo: str):
        """Cross-rank K1 flavour: "ldg" (register-staged 16-byte loads) or "bulk" (bulk-async copies through shared memory)."""
        self._check(self.lib.stk_option_set(self.ctx, _lib.OPT_K1_ALGO, {"ldg": 0, "bulk": 1}[algo]))

    def profile(self, on: bool):
        """Brackets K1 / K2 / accumulate launches with CUDA events inside the library (for bench.py's roofline)."""
        self._check(self.lib.stk_profile_enable(self.ctx, int(on)))

    def profile_read(self, kind: int):
        """(total ms, launches) since the last read for kind 0 = K1, 1 = K2, 2 = accumulate."""
        ms, n = C.c_double(), C.c_int()
        self._check(self.lib.stk_profile_read(self.ctx, kind, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_read_k1_device(self):
        """(ms of K1's barrier-to-barrier NVLink phase, launches, ms of the bucket-zeroing tail) from the device timer."""
        ms, n, z = C.c_double(), C.c_int(), C.c_double()
        self._check(self.lib.stk_profile_read_k1_device(self.ctx, C.byref(ms), C.byref(n), C.byref(z), self._stream()))
        return ms.value, n.value, z.value

    def comm_check(self):
        self._check(self.lib.stk_comm_check(self.ctx, self._stream()))


_ENGINES: Dict[int, Engine] = {}


def get_engine(device: Optional[int] = None, rank: int = 0, world: int = 1, group=None) -> Engine:
    """One engine per (process, device)."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    eng = _ENGINES.get(device)
    if eng is None or eng.ctx is None or eng.world != world or eng.rank != rank:
        eng = Engine(device, rank, world, group)
        _ENGINES[device] = eng
    return eng


# ======================================================================================================================
class ClipSpec:
    """Gradient clipping folded into K1 (norm reduction) and K2 (scaling) -- ClipGradNormConfig / ClipGradConfig."""

    def __init__(self, kind: int = _lib.CLIP_NONE, max_norm: float = 0.0, norm_type: float = 2.0, clip_value: float = 0.0):
        self.kind, self.max_norm, self.norm_type, self.clip_value = kind, float(max_norm), float(norm_type), float(clip_value)

    @property
    def norm_kind(self) -> int:
        if self.kind != _lib.CLIP_NORM:
            return _lib.NORM_NONE
        if self.norm_type == 2.0:
            return _lib.NORM_L2
        if self.norm_type == float("inf"):
            return _lib.NORM_INF
        if self.norm_type <= 0:
            raise ValueError(f"Stoke -- unsupported norm_type {self.norm_type}")
        return _lib.NORM_P


class GradPath:
    """Flat-buffer layout of one model + the per-step launches.

    Layout (all flat, identical element offsets, every parameter 8-element aligned, ``n`` = padded total):

      P      model parameters, model dtype (bf16 | fp32)      peer-visible   what forward/backward read
      G      gradients, model dtype                            peer-visible   ``param.grad`` are views of this
      ACC    fp32 local accumulator (grad_accum > 1 only)      peer-visible
      MAIN   fp32 reduced / unscaled gradients                 peer-visible   all n (DDP) | owned shard (sharded)
      MASTER fp32 master weights                               local          aliases P when the model is fp32
      M, V   fp32 optimizer state                              local          all n (DDP) | owned shard (sharded)
    """

    def __init__(self, engine: Engine, params: Sequence[torch.nn.Parameter], grad_accum: int = 1,
                 clip: Optional[ClipSpec] = None, sharded: bool = False, lp_dtype: Optional[torch.dtype] = None,
                 module: Optional[torch.nn.Module] = None, sync_init: bool = True, needs_second_moment: bool = True,
                 needs_first_moment: bool = True):
        self.engine = engine
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("Stoke -- model has no trainable parameters")
        dev = torch.device("cuda", engine.device)
        for p in self.params:
            if p.device != dev:
                raise ValueError(f"Stoke -- parameter on {p.device}, engine on {dev}")
            if p.dtype != torch.float32:
                raise TypeError("Stoke -- hand the engine an fp32 model; low precision is selected with fp16=...")
        self.grad_accum = max(1, int(grad_accum))
        self.clip = clip or ClipSpec()
        self.sharded = bool(sharded) and engine.world > 1
        self.model_dtype = lp_dtype or torch.float32
        self.low_precision = self.model_dtype != torch.float32
        W = engine.world

        # ---- layout ----
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN_ELEMS - 1) // ALIGN_ELEMS * ALIGN_ELEMS
        self.n = off
        self.shard = engine.shard_range(self.n) if self.sharded else (0, self.n)
        sb, se = self.shard
        self.n_local = se - sb
        esz = _ESZ[self.model_dtype]

        # ---- fp32 source values (rank 0's after the init sync: DDP's _sync_module_states) ----
        full32 = torch.zeros(self.n, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            self._strided(full32, p, o).copy_(p.detach())
        if W > 1 and sync_init:
            torch.distributed.broadcast(full32, src=0, group=engine.group)  # bring-up only (NCCL)

        # ---- buffers ----
        self.P = engine.alloc(self.n * esz)
        self.G = engine.alloc(self.n * esz)
        self.ACC = engine.alloc(self.n * 4) if self.grad_accum > 1 else None
        self.MAIN = engine.alloc(self.n_local * 4)
        self.MASTER = engine.alloc(self.n_local * 4) if self.low_precision else None
        self.M = engine.alloc(self.n_local * 4) if needs_first_moment else None
        self.V = engine.alloc(self.n_local * 4) if needs_second_moment else None

        self.p_flat = self.P.tensor(self.model_dtype, self.n)
        self.g_flat = self.G.tensor(self.model_dtype, self.n)
        self.main_flat = self.MAIN.tensor(torch.float32, self.n_local)
        self.m_flat = self.M.tensor(torch.float32, self.n_local) if self.M else None
        self.v_flat = self.V.tensor(torch.float32, self.n_local) if self.V else None
        if self.low_precision:
            self.master_flat = self.MASTER.tensor(torch.float32, self.n_local)
            self.master_flat.copy_(full32[sb:se])
            self.p_flat.copy_(full32)  # round-to-nearest-even, same as the kernel's cast
            self.master_ptr = self.MASTER.ptr
        else:
            self.p_flat.copy_(full32)
            self.master_flat = self.p_flat[sb:se]
            self.master_ptr = self.P.ptr + sb * 4
        del full32

        # ---- re-point the module at the flat buffers ----
        if self.low_precision and module is not None:
            module.to(self.model_dtype)  # buffers (and any frozen parameter) follow the model dtype
        self.grad_views = []
        for p, o in zip(self.params, self.offsets):
            p.data = self._strided(self.p_flat, p, o)
            gv = self._strided(self.g_flat, p, o)
            p.grad = gv
            self.grad_views.append(gv)
        self._micro = 0
        self._pending_clip = False
        torch.cuda.synchronize(engine.device)
        if W > 1:
            torch.distributed.barrier(group=engine.group)

    @staticmethod
    def _strided(flat: torch.Tensor, like: torch.Tensor, offset: int) -> torch.Tensor:
        if like.is_contiguous():
            return flat[offset: offset + like.numel()].view(like.shape)
        return torch.as_strided(flat, like.shape, like.stride(), storage_offset=flat.storage_offset() + offset)

    # -- per-step API -------------------------------------------------------------------------------------------------
    def ensure_grad_views(self):
        """``zero_grad(set_to_none=True)`` from user code would detach the views; re-attach before backward."""
        for p, gv in zip(self.params, self.grad_views):
            if p.grad is not gv:
                p.grad = gv

    def after_backward(self, sync: bool, unscale: bool):
        """Called once per backward.  ``sync=False``: local accumulation (no_sync); ``sync=True``: K1."""
        e, n = self.engine, self.n
        if not sync:
            if self.ACC is None:
                raise StokeB200Error(_lib.ERR_STATE, "backward without sync but the path was built with grad_accum == 1")
            e.grad_accumulate(self.G.ptr, self.model_dtype, self.ACC.ptr, n, first=(self._micro == 0), zero_grad=True)
            self._micro += 1
            return
        has_acc = self.ACC is not None and self._micro > 0
        flags = _lib.RF_FINAL | _lib.RF_ZERO_INPUT | (_lib.RF_UNSCALE if unscale else 0)
        sb, _ = self.shard
        if self.sharded:
            out_ptrs = [self.MAIN.ptr - sb * 4] * e.world  # only [rank] is used; global element indexing
            mode = _lib.REDUCE_SCATTER
        else:
            out_ptrs = self.MAIN.peer_ptrs()
            mode = _lib.REDUCE_ALL
        e.grad_reduce(mode, self.G.peer_ptrs(), self.model_dtype, self.ACC.peer_ptrs() if has_acc else None, out_ptrs,
                      torch.float32, n, 1.0 / e.world, self.clip.norm_kind, self.clip.norm_type, flags)
        self._micro = 0

    def optimizer_step(self, hyper: _lib.OptimHyper):
        e = self.engine
        hyper.clip_kind = self.clip.kind
        hyper.clip_max_norm = self.clip.max_norm
        hyper.clip_value = self.clip.clip_value
        sb, _ = self.shard
        if self.sharded:
            lp_ptrs, lp_off = self.P.peer_ptrs(), sb
        elif self.low_precision:
            lp_ptrs, lp_off = [self.P.ptr], 0
        else:
            lp_ptrs, lp_off = None, 0
        e.optim_step(hyper, self.master_ptr, self.M.ptr if self.M else None, self.V.ptr if self.V else None,
                     self.MAIN.ptr, self.n_local, lp_ptrs, self.model_dtype, lp_off)
        e.step_epilogue()

    # -- inspection (tests, checkpoints) ----------------------------------------------------------------------------------
    def gather_master(self) -> torch.Tensor:
        """Full fp32 master vector (all-gathered across shards when sharded)."""
        if not self.sharded:
            return self.master_flat.clone()
        return self._gather_shards(self.master_flat)

    def _gather_shards(self, local: torch.Tensor) -> torch.Tensor:
        e = self.engine
        per = e.shard_range(self.n, 0)[1]
        pad = torch.zeros(per, dtype=local.dtype, device=local.device)
        pad[: local.numel()] = local
        out = [torch.empty_like(pad) for _ in range(e.world)]
        torch.distributed.all_gather(out, pad, group=e.group)
        return torch.cat(out)[: self.n]

    def unflatten(self, flat: torch.Tensor) -> List[torch.Tensor]:
        return [self._strided(flat, p, o) for p, o in zip(self.params, self.offsets)]
