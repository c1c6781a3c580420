"""BASELINE ONLY -- not the product path.

The same algorithm the way a stock framework expresses it: the LeNet-like
convnet in ``torch.nn.functional`` (cuDNN convolutions, cuBLAS GEMMs, bf16 autocast), gradients averaged
with one flat-bucket ``ncclAllReduce``, then ``/N`` and a torch SGD update -- i.e. NCCL for the
reduction plus separate elementwise kernels.  ``bench.py --impl torch_ddp`` measures it with the same
timing rules as our engine so BASELINE.md §4 can quote both on the same box.

Two modes: eager (one Python-dispatched launch per op: launch-bound at this model size) and ``--graphed``: the whole
step -- forward, backward, the NCCL all-reduce, scale and SGD -- captured ONCE into a CUDA graph over static buffers and
replayed (channels_last activations, in-graph Philox dropout), i.e. the strongest configuration of the stock stack.
"""
from __future__ import annotations

import json
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F


def _forward(p, x, keep):
    y = F.max_pool2d(F.relu(F.conv2d(x, p["c1w"], p["c1b"], padding=2)), 2, 2)
    y = F.max_pool2d(F.relu(F.conv2d(y, p["c2w"], p["c2b"], padding=2)), 2, 2)
    y = y.permute(0, 2, 3, 1).reshape(y.shape[0], -1)
    h = F.relu(y @ p["f1w"] + p["f1b"]) * keep * 2.0
    return h @ p["f2w"] + p["f2b"]


def _run_graphed(args, ctx, flat, views, imgs, lbls, lr) -> int:
    """Whole step as one CUDA graph (static input slot; fwd + bwd + ncclAllReduce + scale + SGD inside the capture)."""
    dev, n, rank, B = ctx.device, ctx.world_size, ctx.rank, args.batch
    pool = imgs.shape[0]
    # Everything that takes part in the captured autograd graph is created ON the capture stream: autograd ties a leaf's
    # gradient accumulation to the stream the leaf was created / first viewed on, and the legacy default stream must not
    # become dependent on a capturing stream (cudaErrorStreamCaptureImplicit).
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    shapes = {k: tuple(v.shape) for k, v in views.items()}
    with torch.cuda.stream(side):
        flat = flat.detach().clone().requires_grad_(True)
        views, off = {}, 0
        for k, shp in shapes.items():
            m = 1
            for d in shp:
                m *= d
            views[k] = flat[off:off + m].view(shp)
            off += m
        sx = imgs[0].clone().contiguous(memory_format=torch.channels_last)
        sy = lbls[0].clone()
        loss_out = torch.zeros((), device=dev)

    def body():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            keep = (torch.rand(B, 512, device=dev) < 0.5).float()       # in-graph Philox dropout mask
            logits = _forward(views, sx, keep)
        loss = F.cross_entropy(logits.float(), sy)
        (grad,) = torch.autograd.grad(loss, flat)
        if n > 1:
            dist.all_reduce(grad)
            grad /= n
        with torch.no_grad():
            flat.add_(grad, alpha=-lr)
            loss_out.copy_(loss)

    with torch.cuda.stream(side):
        for _ in range(3):
            body()
    torch.cuda.synchronize()
    if n > 1:
        dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
        body()
    torch.cuda.synchronize()

    def step(i):
        with torch.cuda.stream(side):
            sx.copy_(imgs[i % pool], non_blocking=True)
            sy.copy_(lbls[i % pool], non_blocking=True)
            graph.replay()

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    if n > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    for i in range(args.steps):
        step(i)
    e1.record(side)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if n > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "MNIST images/sec (whole box, device-timed, max over ranks)",
                          "impl": "torch_ddp_baseline_cudagraph", "value": n * B * args.steps / (ms.item() / 1e3),
                          "unit": "images/s", "n_gpus": n, "steps": args.steps, "ms_per_step": ms.item() / args.steps,
                          "dtype": "bf16 autocast", "data": "synthetic", "last_loss": float(loss_out.item()),
                          "config": {"model": "LeNet-like", "global_batch": n * B, "cuda_graph": True,
                                     "parallelism": "dp%d NCCL allreduce inside the captured graph" % n}}))
    sys.stdout.flush()
    torch.cuda.synchronize()
    if n > 1:
        dist.barrier()
    # tearing down a process group whose collectives live inside an instantiated CUDA graph can hang in
    # destroy_process_group (observed: 400 s until the launcher's timeout); the measurement is done -- leave without it
    import os
    os._exit(0)


def run_baseline(args) -> int:
    from distributedmnist_b200.flags import FLAGS
    from distributedmnist_b200.parallel.context import init_context, shutdown_context
    ctx = init_context(FLAGS, want_gpu=True)
    dev, n, rank, B = ctx.device, ctx.world_size, ctx.rank, args.batch
    g = torch.Generator().manual_seed(0)
    shapes = {"c1w": (32, 1, 5, 5), "c1b": (32,), "c2w": (64, 32, 5, 5), "c2b": (64,), "f1w": (3136, 512),
              "f1b": (512,), "f2w": (512, 10), "f2b": (10,)}
    numel = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    flat = (torch.randn(numel, generator=g) * 0.05).to(dev).requires_grad_(True)
    views, off = {}, 0
    for k, s in shapes.items():
        m = int(torch.tensor(s).prod())
        views[k] = flat[off:off + m].view(s)
        off += m
    pool = 64
    imgs = (torch.rand(pool, B, 1, 28, 28, device=dev) - 0.5).contiguous(memory_format=torch.channels_last_3d) \
        if False else (torch.rand(pool, B, 1, 28, 28, device=dev) - 0.5)
    lbls = torch.randint(0, 10, (pool, B), device=dev)
    lr = 0.01

    graphed = bool(getattr(args, "graphed", False))
    if graphed:
        return _run_graphed(args, ctx, flat, views, imgs, lbls, lr)

    def step(i):
        x, y = imgs[i % pool], lbls[i % pool]
        keep = (torch.rand(B, 512, device=dev) < 0.5).float()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = _forward(views, x, keep)
        loss = F.cross_entropy(logits.float(), y)
        (grad,) = torch.autograd.grad(loss, flat)
        if n > 1:
            dist.all_reduce(grad)          # flat bucket, ncclAllReduce
            grad /= n                      # separate scale kernel
        with torch.no_grad():
            flat.add_(grad, alpha=-lr)     # separate SGD kernel
        return loss

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    if n > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if n > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "MNIST images/sec (whole box, device-timed, max over ranks)",
                          "impl": "torch_ddp_baseline", "value": n * B * args.steps / (ms.item() / 1e3),
                          "unit": "images/s", "n_gpus": n, "steps": args.steps, "ms_per_step": ms.item() / args.steps,
                          "dtype": "bf16 autocast", "data": "synthetic",
                          "config": {"model": "LeNet-like", "global_batch": n * B, "parallelism": "dp%d NCCL allreduce" % n}}))
    sys.stdout.flush()
    shutdown_context(ctx)
    return 0
