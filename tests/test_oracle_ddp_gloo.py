"""world_size-2 gloo test (CPU): pins the oracle's W-rank reduce against real ``torch.nn.parallel.DistributedDataParallel``
driven in stoke's call order (no_sync on the non-final micro-steps, clip, step, zero_grad) -- the path the reference's
``DistributedDDP`` + ``BaseDDP.handle_ddp`` takes (/root/reference/stoke/distributed.py:648-669, extensions.py:207-215), on
the gloo backend its ``DDPConfig.backend`` admits (configs.py:36-41).  At W = 2 the sum is order-independent and 1/2 is
exact, so the comparison is bit-exact."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _batch(rank, step):
    g = torch.Generator().manual_seed(1000 * step + rank)
    return torch.randn(16, 128, generator=g), (torch.rand(16, 1, generator=g) > 0.5).float()


def _worker(rank, world, port, accum, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from engine_oracle import OracleEngine
    from stoke_b200 import synthetic

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = synthetic.basic_nn(4)
    shadow = synthetic.basic_nn(4)  # plain copy used to obtain every rank's LOCAL gradients for the oracle
    ddp = torch.nn.parallel.DistributedDataParallel(model, bucket_cap_mb=25, broadcast_buffers=True,
                                                    find_unused_parameters=False, gradient_as_bucket_view=False)
    opt = torch.optim.Adam(ddp.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-9)
    oracle = OracleEngine(list(shadow.parameters()), world, torch.optim.Adam,
                          {"lr": 1e-3, "betas": (0.9, 0.98), "eps": 1e-9}, grad_accum=accum, clip=("norm", 0.05, 2.0))
    lossf = torch.nn.BCEWithLogitsLoss()
    step = 0
    for opt_step in range(6):
        for micro in range(accum):
            # the reference: nullcontext on the sync step, model.no_sync() otherwise (stoke.py:978-984)
            ctx = ddp.no_sync() if micro < accum - 1 else torch.autograd.profiler.record_function("sync")
            x, y = _batch(rank, step)
            with ctx:
                (lossf(ddp(x), y) / accum).backward()
            # every rank's local gradient at the current weights, for all ranks (the oracle holds W logical ranks)
            per_rank = []
            for r in range(world):
                xr, yr = _batch(r, step)
                per_rank.append(torch.autograd.grad(lossf(shadow(xr), yr) / accum, list(shadow.parameters())))
            oracle.micro_step(per_rank)
            step += 1
        torch.nn.utils.clip_grad_norm_(ddp.parameters(), max_norm=0.05, norm_type=2.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        oracle.step()
        with torch.no_grad():
            for p, w in zip(shadow.parameters(), oracle.weights()):
                p.copy_(w)
    got = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    ok = torch.equal(got, oracle.flat_weights())
    maxdiff = (got - oracle.flat_weights()).abs().max().item()
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok, maxdiff))
    if rank == 0:
        torch.save(gathered, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("accum", [1, 2])
def test_oracle_reduce_matches_ddp_gloo_world2(tmp_path, accum):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), accum, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert all(ok for ok, _ in res), res


def test_sampler_replicas_world2_gloo(tmp_path):
    """Host side of the sampler under a real 2-rank process group: rank / world are taken from torch.distributed like the
    reference does when num_replicas / rank are None (data.py:299-343); planning runs through the C ABI (no GPU needed)."""
    out = str(tmp_path / "plan.pt")
    mp.spawn(_sampler_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    plans = torch.load(out, weights_only=False)
    assert [p["rank"] for p in plans] == [0, 1]
    assert all(p["world"] == 2 and p["len"] == plans[0]["len"] for p in plans)


def _sampler_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import stoke_b200 as sb
    from stoke_b200 import synthetic

    n = 4099
    s = sb.BucketedDistributedSampler(range(n), buckets=4, batch_size=8, sorted_idx=synthetic.sampler_sorted_idx(n),
                                      info_rank=-1)
    res = [None] * world
    dist.all_gather_object(res, {"rank": s.rank, "world": s.num_replicas, "len": len(s)})
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()
