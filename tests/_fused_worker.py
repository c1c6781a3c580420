"""Helper process for the multi-GPU fused allreduce+SGD tests (one rank per GPU)."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from distributedmnist_b200.parallel.context import init_context, shutdown_context  # noqa: E402
from distributedmnist_b200.parallel.fused import FusedBackend  # noqa: E402


def fp(t):
    return hashlib.sha1(t.detach().cpu().numpy().tobytes()).hexdigest()


def main():
    out_json = sys.argv[1].replace("RANK", os.environ["RANK"])
    numel = int(sys.argv[2])
    ctx = init_context(None, want_gpu=True)
    n, r = ctx.world_size, ctx.rank
    be = FusedBackend(ctx, timeout_ms=10000.0)
    params, grads = be.allocate(numel), be.allocate(numel)
    be.attach_shadow(params)
    g = torch.Generator(device="cpu").manual_seed(1234)
    params.copy_(torch.randn(numel, generator=g).to(ctx.device))
    res = {"rank": r, "full": [], "kofn": [], "bw": []}
    be.barrier()

    # ---- full participation: K == N, compare with NCCL allreduce -> /N -> SGD -------------------
    ref = params.clone()
    for step in range(6):
        gg = torch.Generator(device="cpu").manual_seed(1000 * step + r)
        grads.copy_(torch.randn(numel, generator=gg).to(ctx.device))
        mean = grads.clone()
        dist.all_reduce(mean)
        mean /= n
        lr = 0.1 / (step + 1)
        ref -= lr * mean
        info = be.sync_step(params, grads, lr, step, n)
        torch.cuda.synchronize()
        err = (params - ref).abs().max().item()
        sh_err = (be.shadow.float() - params).abs().max().item()
        res["full"].append({"step": info.global_step, "mask": info.mask, "count": info.count,
                            "accepted": info.accepted, "err": err, "shadow_err": sh_err, "fp": fp(params)})
        ref.copy_(params)   # keep following the kernel's own (reduction-order) result
    be.barrier()

    # ---- K of N with a device-side straggler on the last rank ----------------------------------------
    k = max(n - 1, 1)
    if n > 1:
        for step in range(6, 12):
            gg = torch.Generator(device="cpu").manual_seed(1000 * step + r)
            grads.copy_(torch.randn(numel, generator=gg).to(ctx.device))
            all_g = [torch.empty_like(grads) for _ in range(n)]
            dist.all_gather(all_g, grads)
            before = params.clone()
            torch.cuda.synchronize()
            be.barrier()
            info = be.sync_step(params, grads, 0.05, step, k, delay_s=0.02 if r == n - 1 else 0.0)
            torch.cuda.synchronize()
            be.barrier()   # late rank returns early by design; make sure all pushes are observed
            members = [q for q in range(n) if (info.mask >> q) & 1]
            expect = before - 0.05 * torch.stack([all_g[q] for q in members]).sum(0) / max(len(members), 1)
            res["kofn"].append({"step": info.global_step, "mask": info.mask, "count": info.count,
                                "accepted": info.accepted, "err": (params - expect).abs().max().item(),
                                "fp": fp(params)})

    # ---- bandwidth sweep (device-timed) ----------------------------------------------------------------------
    for nb in [int(x) for x in os.environ.get("DM_BW_SIZES", "").split(",") if x]:
        ne = nb // 4
        if ne > numel:
            continue
        p2, g2 = params[:ne], grads[:ne]
        be._by_ptr[p2.data_ptr()] = be._by_ptr[params.data_ptr()]
        be._by_ptr[g2.data_ptr()] = be._by_ptr[grads.data_ptr()]
        for _ in range(3):
            be.enqueue(p2, g2, n, lr0=0.0)
        torch.cuda.synchronize(); be.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            be.enqueue(p2, g2, n, lr0=0.0)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=ctx.device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res["bw"].append({"bytes": nb, "ms": ms.item()})
    be.check_error()
    with open(out_json, "w") as f:
        json.dump(res, f)
    be.close()
    shutdown_context(ctx)


if __name__ == "__main__":
    main()
