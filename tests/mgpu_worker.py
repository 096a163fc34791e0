"""Worker for the multi-GPU parity tests: launched by tests/test_gpu_multi.py through torch.distributed.run with one
process per GPU.  Every rank drives the engine with its own seeded gradients; the CPU oracle (W logical ranks) consumes the
same gradients; rank 0 writes a JSON verdict."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main(out_path):
    from engine_oracle import OracleEngine
    from test_gpu_engine import OddNet, _oracle_flat

    from stoke_b200 import _lib, synthetic
    from stoke_b200.engine import ClipSpec, get_engine
    from stoke_b200.optim import B200FusedOptimizer

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = get_engine(local, rank, world)
    results = {}

    def inject(path, step, r, dtype, scale=1.0):
        out = []
        for i, gv in enumerate(path.grad_views):
            g = synthetic.injected_grad(gv.numel(), r, step * 100 + i, dtype=dtype, scale=scale).view(gv.shape)
            if r == rank:
                gv.copy_(g.cuda())
            out.append(g.float())
        return out

    # (name, optimizer, kwargs, model dtype, clip, accum, fairscale_oss, route (None: the default -- sharded), bucket MB)
    cases = [
        ("ddp_adam_bf16_clipnorm", torch.optim.Adam, {"lr": 1e-3}, torch.bfloat16, ("norm", 1.0, 2.0), 1, False, None, 0.008),
        ("ddp_adam_fp32_accum2", torch.optim.Adam, {"lr": 1e-3, "weight_decay": 0.01}, None, ("norm", 0.5, 2.0), 2, False, None, 0.016),
        ("ddp_sgd_bf16_clipvalue", torch.optim.SGD, {"lr": 0.05, "momentum": 0.9}, torch.bfloat16, ("value", 0.2), 1, False, None, 25.0),
        ("allreduce_adam_bf16_clipnorm", torch.optim.Adam, {"lr": 1e-3}, torch.bfloat16, ("norm", 1.0, 2.0), 1, False, "allreduce", 0.008),
        ("allreduce_adamw_fp32_accum2", torch.optim.AdamW, {"lr": 1e-3, "weight_decay": 0.05}, None, ("norm", 1.0, 2.0), 2, False, "allreduce", 25.0),
        ("oss_adam_bf16_clipnorm", torch.optim.Adam, {"lr": 1e-3}, torch.bfloat16, ("norm", 1.0, 2.0), 1, True, None, 0.008),
        ("oss_adamw_fp32_accum3_inf", torch.optim.AdamW, {"lr": 1e-3, "weight_decay": 0.05}, None,
         ("norm", 2.0, float("inf")), 3, True, None, 25.0),
    ]
    nvls = os.environ.get("STK_K1_ALGO") == "nvls"
    # the multimem flavour rounds 16-bit sums to the input type inside the switch (NCCL's NVLS numerics): its bar for the
    # bf16 cases is the bf16 rounding of the gradient sum, not the fp32-exact 1e-5 of the other flavours
    tol_of = {}
    for name, cls, kw, lp, clip, accum, sharded, route, bucket_mb in cases:
        torch.manual_seed(1234 + rank)  # different init per rank: the engine must broadcast rank 0's
        net = OddNet(scale=3).cuda()
        if rank != 0:
            with torch.no_grad():
                for p in net.parameters():
                    p.add_(1.0)
        spec = ClipSpec(_lib.CLIP_NORM, max_norm=clip[1], norm_type=clip[2]) if clip[0] == "norm" else \
            ClipSpec(_lib.CLIP_VALUE, clip_value=clip[1])
        torch.manual_seed(1234)
        init = [p.detach().cpu().clone() for p in OddNet(scale=3).parameters()]
        opt = B200FusedOptimizer(net, cls, kw, engine=eng, grad_accum=accum, clip=spec, sharded=sharded, lp_dtype=lp,
                                 route=route, bucket_mb=bucket_mb)
        path = opt.path
        oracle = OracleEngine(init, world, cls, kw, grad_accum=accum, clip=clip)
        gdtype = lp or torch.float32
        for step in range(8):
            for micro in range(accum):
                grads = [inject(path, step * accum + micro, r, gdtype) for r in range(world)]
                oracle.micro_step(grads)
                path.after_backward(sync=(micro == accum - 1), unscale=False)
            opt.step()
            oracle.step()
        eng.comm_check()
        ref = _oracle_flat(path, oracle.weights())
        got = path.gather_master().cpu()
        err = rel(got, ref)
        # replicas must be bit-identical: compare this rank's model copy with rank 0's
        mine = path.p_flat.float().clone()
        zero = mine.clone()
        torch.distributed.broadcast(zero, src=0)
        same = bool(torch.equal(mine, zero))
        flags = torch.tensor([1.0 if same else 0.0], device="cuda")
        torch.distributed.all_reduce(flags, op=torch.distributed.ReduceOp.MIN)
        norm_err = None
        if clip[0] == "norm":
            norm_err = abs(eng.scaler_get(path.state_id).grad_norm - float(oracle.last_total_norm)) / float(oracle.last_total_norm)
        results[name] = {"rel_err": err, "replicas_identical": bool(flags.item() == 1.0), "norm_rel_err": norm_err,
                         "model_is_rounded_master": bool(torch.equal(path.p_flat.cpu(), got.to(path.model_dtype))),
                         "route": path.route, "buckets": len(path.buckets), "mc": bool(path.G.mc_ptr),
                         "tol": 5e-3 if (nvls and lp is not None and path.G.mc_ptr and accum == 1) else 1e-5}
        opt.close()
        del opt, path, net

    # ---- full-size known-answer test (ResNet-50-sized bucket): exactly representable inputs -> exact expected output ----
    class Big(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(25_557_032))

    for route in ("allreduce", "sharded"):
        net = Big().cuda()
        opt = B200FusedOptimizer(net, torch.optim.SGD, {"lr": 0.0}, engine=eng,
                                 clip=ClipSpec(_lib.CLIP_NORM, max_norm=1.0, norm_type=2.0), lp_dtype=torch.bfloat16, route=route)
        path = opt.path
        base = ((torch.arange(path.n, device="cuda") % 7) - 3).float()           # -3..3
        path.g_flat.copy_((base * (rank + 1)).to(torch.bfloat16))                 # rank r contributes (r+1) * base
        path.after_backward(sync=True, unscale=False)
        expect = base * (world + 1) / 2.0                                         # mean over ranks of (r+1)
        if route == "allreduce":
            kat_ok = bool(torch.equal(path.main_flat, expect))
        else:
            kat_ok = all(bool(torch.equal(path.main_flat[l0: l0 + (g1 - g0)], expect[g0:g1])) for g0, g1, l0, _ in path.segs)
        exp_norm = float(expect.double().pow(2).sum().sqrt())
        got_norm = eng.scaler_get(path.state_id).grad_norm
        zeroed = float(path.g_flat.float().abs().max()) == 0.0
        results["kat_full_size" if route == "allreduce" else "kat_full_size_sharded"] = {
            "exact": kat_ok, "norm_rel_err": abs(got_norm - exp_norm) / exp_norm, "bucket_zeroed": zeroed, "mc": bool(path.G.mc_ptr)}
        eng.step_epilogue()
        opt.close()
        del opt, path, net

    # loss mean / barrier / inf propagation across ranks
    t = torch.tensor(float(rank + 1), device="cuda")
    results["loss_sync"] = eng.loss_sync(t)
    results["loss_sync_expected"] = sum(range(1, world + 1)) / world
    eng.barrier()

    from stoke_b200.fp16 import DeviceGradScaler
    net = OddNet().cuda()
    opt = B200FusedOptimizer(net, torch.optim.Adam, {"lr": 1e-2}, engine=eng,
                             clip=ClipSpec(_lib.CLIP_NORM, max_norm=1.0, norm_type=2.0))
    scaler = DeviceGradScaler(eng, init_scale=2.0**8, growth_interval=1000, state_id=opt.path.state_id)
    before = opt.path.gather_master().clone()
    inject(opt.path, 0, rank, torch.float32, scale=2.0**8)
    if rank == world - 1:
        opt.path.grad_views[0].view(-1)[7] = float("inf")   # only the LAST rank sees the overflow
    opt.path.after_backward(sync=True, unscale=True)
    opt.step()
    st = eng.scaler_get(opt.path.state_id)
    results["inf_skip"] = {"unchanged": bool(torch.equal(before, opt.path.gather_master())), "scale": st.scale,
                           "skipped": st.skipped_steps, "steps": st.opt_steps}
    opt.close()
    eng.comm_check()
    results["caps"] = {"mem_mode": eng.mem_mode, "multicast": eng.multicast, "k1_algo": os.environ.get("STK_K1_ALGO")}
    results["stoke_api"] = stoke_api_section(rank, world, local, os.path.dirname(out_path))
    eng.comm_check()
    torch.distributed.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(results, f)
    torch.distributed.destroy_process_group()


def stoke_api_section(rank, world, local, tmpdir):
    """The public API under DDP / OSS / SDDP on real GPUs: replicas stay bit-identical, BatchNorm buffers follow rank 0
    (broadcast_buffers), the synced loss is the mean over ranks, and save -> load -> continue reproduces the uninterrupted run."""
    import stoke_b200 as sb

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(21)
            self.fc1 = torch.nn.Linear(64, 96)
            self.bn = torch.nn.BatchNorm1d(96)
            self.fc2 = torch.nn.Linear(96, 1)
            with torch.no_grad():
                for p in self.parameters():
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1)

        def forward(self, x):
            return self.fc2(torch.relu(self.bn(self.fc1(x))))

    def batch(step, r):
        g = torch.Generator().manual_seed(77 * step + r)
        return torch.randn(32, 64, generator=g).cuda(), (torch.rand(32, 1, generator=g) > 0.5).float().cuda()

    def all_equal(t):
        ref = t.clone()
        torch.distributed.broadcast(ref, src=0)
        flag = torch.tensor([1.0 if torch.equal(ref, t) else 0.0], device="cuda")
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        return bool(flag.item() == 1.0)

    out = {}
    for name, flags in (("ddp", {}), ("oss", {"fairscale_oss": True}), ("oss_sddp", {"fairscale_oss": True, "fairscale_sddp": True})):
        def make():
            return sb.Stoke(model=Net(), optimizer=sb.StokeOptimizer(optimizer=torch.optim.AdamW, optimizer_kwargs={"lr": 2e-3}),
                            loss=torch.nn.BCEWithLogitsLoss(), batch_size_per_device=32, grad_accum_steps=2,
                            grad_clip=sb.ClipGradNormConfig(max_norm=0.5, norm_type=2.0), gpu=True, fp16="bf16",
                            distributed="ddp", configs=[sb.DDPConfig(local_rank=local)], verbose=False, **flags)

        def run(s, steps):
            synced = []
            for st in steps:
                x, y = batch(st, rank)
                s.backward(s.loss(s.model(x), y))
                synced.append(s.step_loss)
                s.step()
            return synced

        a = make()
        la = run(a, range(0, 6))
        path, tag = a.save(tmpdir, name=f"ck_{name}")
        la += run(a, range(6, 12))
        wa = a.optimizer.path.gather_master().clone()
        # like torch DDP, rank 0's buffers are broadcast at the START of a forward; ranks then update their BatchNorm
        # statistics locally.  An eval-mode forward (no statistics update) must therefore leave every rank with rank 0's.
        a.model_access.eval()
        a.model(batch(99, rank)[0])
        a.model_access.train()
        bufs = torch.cat([b.detach().float().reshape(-1) for b in a.model_access.buffers()])
        # every rank reports the same (mean) loss
        lt = torch.tensor(la, device="cuda", dtype=torch.float64)
        b = make()
        b.load(path, tag)
        lb = run(b, range(6, 12))
        wb = b.optimizer.path.gather_master()
        out[name] = {
            "replicas_identical": all_equal(a.optimizer.path.p_flat.float()),
            "buffers_identical": all_equal(bufs),
            "loss_identical_across_ranks": all_equal(lt),
            "resume_bit_identical": bool(torch.equal(wa, wb)) and la[6:] == lb,
            "opt_steps": a._optimizer_steps, "sharded": a.optimizer.path.sharded, "buckets": len(a.optimizer.path.buckets),
            "overlap": a.optimizer.path.overlap,
            "loss_first_last": [la[0], la[-1]],
        }
        a.close()
        b.close()
        del a, b
    return out


if __name__ == "__main__":
    main(sys.argv[1])
