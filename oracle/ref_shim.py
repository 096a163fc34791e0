# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- import shim for the unmodified reference (fidelity/stoke).

Looks for the reference package at ``/root/reference`` (build container) and, failing that, at ``oracle/_ref`` -- the
verbatim, git-ignored copy that ``oracle/build_ref.py`` makes and that travels to the GPU box with the snapshot.
It is used by ``oracle/make_golden.py`` to generate the fixtures under ``tests/golden/``, by the CPU tests that validate
``oracle/`` against the real reference, and by ``bench.py``'s ``--impl reference`` / ``cpu_baseline`` legs.

Why a shim: the reference imports ``horovod``, ``deepspeed`` and ``fairscale`` at module scope
(/root/reference/stoke/data.py:12, distributed.py:14-18, io_ops.py:12-15, fp16.py:14-15, extensions.py:14-15,
utils.py:13, stoke.py:13-14) and none of them is installable here.  Registering empty stand-in modules lets the
*unmodified* reference import; its CPU path (``DistributedNullCPU + NullFP16 + BaseOptimizer``) and its
``BucketedDistributedSampler`` then run as shipped.  Nothing in the product package imports this file.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("STOKE_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "stoke")) and os.path.isdir(os.path.join(_HERE, "_ref", "stoke")):
    REFERENCE_ROOT = os.path.join(_HERE, "_ref")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "stoke"))


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch

    if "horovod" not in sys.modules:
        hvd = _mod("horovod")
        hvd.torch = _mod("horovod.torch")
    if "deepspeed" not in sys.modules:
        ds = _mod("deepspeed")
        ds.utils = _mod("deepspeed.utils")
        ds.utils.distributed = _mod("deepspeed.utils.distributed")
        ds.utils.distributed.mpi_discovery = lambda *a, **k: None
    if "fairscale" not in sys.modules:
        fs = _mod("fairscale")
        fs.nn = _mod("fairscale.nn")
        fs.nn.data_parallel = _mod("fairscale.nn.data_parallel")

        class FullyShardedDataParallel(torch.nn.Module):
            pass

        class ShardedDataParallel(torch.nn.Module):
            pass

        fs.nn.data_parallel.FullyShardedDataParallel = FullyShardedDataParallel
        fs.nn.data_parallel.ShardedDataParallel = ShardedDataParallel
        fs.optim = _mod("fairscale.optim")
        fs.optim.oss = _mod("fairscale.optim.oss")

        class OSS(torch.optim.Optimizer):
            pass

        fs.optim.oss.OSS = OSS
        fs.optim.grad_scaler = _mod("fairscale.optim.grad_scaler")
        fs.optim.grad_scaler.ShardedGradScaler = object


def import_reference():
    """Returns the unmodified reference package (``import stoke`` from /root/reference)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import stoke  # noqa: E402

    return stoke
