"""Helper process: device-side interval mode (csrc/fused_interval.cu) driven directly, no model.

Every replica free-runs `iters` iterations of [adopt -> (delay) -> gate -> accumulate -> close -> apply]; the gradient is the
same constant tensor on every replica and iteration, so the mean of ANY subset of accumulated gradients is that tensor and
after T committed ticks the weights must be  w0 - sum_t lr(t) * g  exactly (up to fp32 rounding), whoever contributed.
The last rank is delayed on the device so that it misses ticks.

argv: <out json with RANK placeholder> <interval_ms> <iters> <delay_us of the last rank>"""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from distributedmnist_b200.parallel.context import init_context, shutdown_context  # noqa: E402
from distributedmnist_b200.parallel.fused import FusedBackend  # noqa: E402


def main():
    out_json = sys.argv[1].replace("RANK", os.environ.get("RANK", "0"))
    interval_ms, iters, delay_us = float(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    ctx = init_context(None, want_gpu=True)
    n, r = ctx.world_size, ctx.rank
    be = FusedBackend(ctx, timeout_ms=10000.0)
    numel = 1 << 18
    params, grads, acc = be.allocate(numel), be.allocate(numel), be.allocate(numel)
    be.attach_shadow(params)
    gen = torch.Generator(device="cpu").manual_seed(99)
    w0 = torch.randn(numel, generator=gen).to(ctx.device)
    g = torch.randn(numel, generator=gen).to(ctx.device)
    params.copy_(w0)
    grads.copy_(g)
    be.refresh_shadow(params)
    lr = 0.01
    be.interval_arm(interval_ms)
    rows = []
    slow = n > 1 and r == n - 1
    if slow:
        iters = max(4, int(iters * 250.0 / max(delay_us, 250.0)))      # about the same wall time as the fast replicas
    for it in range(iters):
        be.enqueue_interval_begin(params)
        if slow:
            be.enqueue_straggler_delay(1.0, delay_us)
        else:
            be.enqueue_straggler_delay(1.0, 200.0)           # "compute" of an ordinary replica: 0.2 ms per iteration
        be.enqueue_interval_end(params, grads, acc, lr)
        st = be.read_status()
        rows.append([st["epoch"], st["last_late"], st["last_mask"], st["last_count"]])
    torch.cuda.synchronize()
    be.check_error()
    # quiesce: wait until every replica has stopped committing, then adopt whatever was pushed last
    be.barrier()
    time.sleep(0.05)
    be.enqueue_interval_begin(params)
    torch.cuda.synchronize()
    be.barrier()
    st = be.read_status()
    steps = st["epoch"]
    expect = w0 - steps * lr * g
    err = (params - expect).abs().max().item()
    shadow_err = (be.shadow.float() - params).abs().max().item()
    json.dump({"rank": r, "steps": steps, "err": err, "shadow_err": shadow_err, "pmax": params.abs().max().item(),
               "fp": hashlib.sha1(params.cpu().numpy().tobytes()).hexdigest(), "rows": rows,
               "accepted": st["accepted_steps"], "dropped": st["dropped_steps"],
               "ticks_committed": be._read_u32("iv_ticks_committed")}, open(out_json, "w"))
    be.close()
    shutdown_context(ctx)


if __name__ == "__main__":
    main()
