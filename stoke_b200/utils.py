# -*- coding: utf-8 -*-
"""Small helpers with the reference's names (/root/reference/stoke/utils.py)."""
import os
from enum import Enum
from typing import Any, Callable, List, Tuple, TypeVar, Union

import torch

T_co = TypeVar("T_co", covariant=True)
T = TypeVar("T")
_worker_init_fn_t = Callable[[int], None]
_collate_fn_t = Callable[[List[T]], Any]


class ParamNormalize(Enum):
    THOUSAND = 1e3
    MILLION = 1e6
    BILLION = 1e9
    TRILLION = 1e12


def place_data_on_gpu(data, fp16=None, non_blocking: bool = True):
    """Moves tensors nested in lists / tuples / dicts to the current CUDA device (utils.py:39-80).  Copies are
    asynchronous when the source is pinned; non-tensor leaves pass through like the reference."""
    if isinstance(data, torch.Tensor):
        return data.to(device="cuda", dtype=data.dtype, non_blocking=non_blocking)
    if isinstance(data, (list, tuple)):
        return type(data)(place_data_on_gpu(v, fp16, non_blocking) for v in data)
    if isinstance(data, dict):
        return {k: place_data_on_gpu(v, fp16, non_blocking) for k, v in data.items()}
    return data


def zero_optimizer_grads(optimizer, apex: bool = False, horovod: bool = False):
    optimizer.zero_grad(set_to_none=True)


def unrolled_print(msg: Union[str, List[str], Tuple[str]], single_line: bool = False):
    if isinstance(msg, (list, tuple)):
        parts = [f"Stoke -- {m}" if (i == 0 or not single_line) else f"{m}" for i, m in enumerate(msg)]
        print(*parts, sep=", " if single_line else "\n")
    else:
        print(f"Stoke -- {msg}")


def make_folder(path: str):
    if not os.path.isdir(path):
        os.makedirs(path, exist_ok=True)
