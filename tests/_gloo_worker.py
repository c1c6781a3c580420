"""Helper process for the CPU/gloo distributed tests: runs the real training driver
and dumps a per-step fingerprint of the replica's parameters."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from distributedmnist_b200 import data as mnist_data  # noqa: E402
from distributedmnist_b200.flags import FLAGS  # noqa: E402
from distributedmnist_b200.parallel import backends as B  # noqa: E402
from distributedmnist_b200.parallel.context import init_context, shutdown_context  # noqa: E402
from distributedmnist_b200.train import train  # noqa: E402


def main():
    out_json = sys.argv[1]
    FLAGS.parse(sys.argv[2:])
    torch.set_num_threads(1)
    ctx = init_context(FLAGS, want_gpu=False)
    fingerprints, infos = [], []
    orig = B._CollectiveBackend.sync_step

    def spy(self, params, grads, lr, local_step, k, delay_s=0.0):
        info = orig(self, params, grads, lr, local_step, k, delay_s)
        fingerprints.append(hashlib.sha1(params.numpy().tobytes()).hexdigest())
        infos.append([info.global_step, info.accepted, info.mask, info.count])
        return info
    B._CollectiveBackend.sync_step = spy
    ds = mnist_data.load_mnist(FLAGS.data_dir, worker_id=ctx.rank, n_workers=ctx.world_size, seed=FLAGS.seed,
                               synthetic=True, synthetic_sizes=(512, 128))
    res = train(ctx, ds.train, ds.validation, FLAGS)
    with open(out_json, "w") as f:
        json.dump({"rank": ctx.rank, "fingerprints": fingerprints, "infos": infos,
                   "final_step": res["final_step"], "accepted": res["accepted"], "dropped": res["dropped"],
                   "losses": res["losses"],
                   "final_fp": hashlib.sha1(res["params"].numpy().tobytes()).hexdigest()}, f)
    shutdown_context(ctx)


if __name__ == "__main__":
    main()
