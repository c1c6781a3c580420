"""Helper process (CPU, gloo): pass file descriptors between ranks with parallel/symm_mem.py::_FdChannel."""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch.distributed as dist  # noqa: E402

from distributedmnist_b200.parallel.context import init_context, shutdown_context  # noqa: E402
from distributedmnist_b200.parallel.symm_mem import _all_ok, _FdChannel  # noqa: E402


def main():
    out_json = sys.argv[1].replace("RANK", os.environ["RANK"])
    ctx = init_context(None, want_gpu=False)
    n, r = ctx.world_size, ctx.rank
    chan = _FdChannel.get(r, n)
    f = tempfile.TemporaryFile()
    f.write(("hello from rank %d" % r).encode())
    f.flush()
    texts = []
    for rnd in range(2):                                   # two exchanges: tags keep them apart
        fds = chan.all_to_all(f.fileno())
        row = []
        for q, fd in enumerate(fds):
            if q == r:
                row.append("self")
                continue
            row.append(os.pread(fd, 100, 0).decode())       # (a passed fd shares its file OFFSET with every other holder)
            os.close(fd)
        texts.append(row)
    b = chan.broadcast(f.fileno() if r == 0 else None, 0)
    bc = os.pread(b, 100, 0).decode()
    ok = _all_ok(r != 1) if n > 1 else True                # rank 1 "fails" -> everybody must see False
    json.dump({"rank": r, "texts": texts, "bcast": bc, "all_ok": ok}, open(out_json, "w"))
    dist.barrier()
    shutdown_context(ctx)


if __name__ == "__main__":
    main()
