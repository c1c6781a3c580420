"""Helper process: the two bucket kernels of csrc/fused_bucket.cu driven directly (no model), checked against NCCL.

argv: <out json with RANK placeholder>.  DMNIST_NVLS selects multimem (1) or peer loads/stores (0)."""
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
    out_json = sys.argv[1].replace("RANK", os.environ.get("RANK", "0"))
    ctx = init_context(None, want_gpu=True)
    n, r = ctx.world_size, ctx.rank
    be = FusedBackend(ctx, timeout_ms=10000.0)
    numel, e0, e1 = 1665024, 52160, 52160 + 1605632          # the LeNet arena: fc1 weights = early bucket
    n_late = numel - (e1 - e0)
    params, grads = be.allocate(numel), be.allocate(numel)
    be.attach_shadow(params)
    g16 = be.allocate_buffer((e1 - e0) * 2)
    inbox = be.allocate_buffer(2 * n * n_late * 8) if n > 1 else None
    g16v = g16.view(torch.bfloat16, 0, e1 - e0)
    gen = torch.Generator(device="cpu").manual_seed(4321)
    params.copy_(torch.randn(numel, generator=gen).to(ctx.device))
    be.refresh_shadow(params)
    late_idx = torch.cat([torch.arange(0, e0), torch.arange(e1, numel)]).to(ctx.device)
    rows = []
    be.barrier()
    for step in range(5):
        gg = torch.Generator(device="cpu").manual_seed(77 * step + r)
        full = torch.randn(numel, generator=gg).to(ctx.device)
        grads.copy_(full)
        g16v.copy_(full[e0:e1].to(torch.bfloat16))
        # ---- oracle: NCCL all-reduce (fp32 sums); fc1: sum of the bf16-rounded gradients, rounded to bf16 once more ----
        late_sum = full[late_idx].clone()
        if be.late_ll and be.late_bf16:
            late_sum = late_sum.to(torch.bfloat16).float()        # the late bucket crosses the wire as bf16 LL lines
        fc1_sum = g16v.float().clone()
        if n > 1:
            dist.all_reduce(late_sum)
            dist.all_reduce(fc1_sum)
        lr = 0.1 * (0.5 ** (step // 2))
        ref = params.clone()
        ref[late_idx] -= (lr / n) * late_sum
        fc1_ref_lo = params[e0:e1] - (lr / n) * fc1_sum          # before the bf16 rounding of the sum
        torch.cuda.synchronize()
        be.barrier()
        be.enqueue_bucket_v2(params, grads, g16, inbox, 1, e0, e1, 0.1, 0.5, 2)
        be.enqueue_bucket_v2(params, grads, g16, inbox, 2, e0, e1, 0.1, 0.5, 2)
        torch.cuda.synchronize()
        be.check_error()
        info = be.last_step_info()
        late_err = (params[late_idx] - ref[late_idx]).abs().max().item()
        # fc1: the applied sum is bf16(sum): |error| <= lr/n * 2^-8 * |sum| (+ fp32 noise)
        fc1_tol = (lr / n) * fc1_sum.abs() * 2.0 ** -7 + 2e-6       # one bf16 ulp of the sum (round-to-nearest is half of it)
        fc1_bad = ((params[e0:e1] - fc1_ref_lo).abs() > fc1_tol).sum().item()
        red = g16v.float()                                        # the reduced gradient, identical on every rank
        sum_err = ((red - fc1_sum).abs() / (fc1_sum.abs() + 1e-3)).max().item()
        rows.append({"step": info.global_step, "mask": info.mask, "count": info.count, "late_err": late_err,
                     "fc1_bad": fc1_bad, "sum_rel_err": sum_err, "fp": fp(params), "fp_g16": fp(g16v.view(torch.int16)),
                     "shadow_err": (be.shadow.float() - params).abs().max().item(), "pmax": params.abs().max().item()})
    json.dump({"rank": r, "nvls": bool(be.use_nvls and g16.multicast_ptr), "mc_ptr": int(g16.multicast_ptr),
               "alloc": type(g16).__name__, "rows": rows,
               "phases": be.read_phases(), "phases_early": be.read_phases_early()}, open(out_json, "w"))
    be.close()
    shutdown_context(ctx)


if __name__ == "__main__":
    main()
