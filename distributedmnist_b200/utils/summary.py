"""Scalar summaries: a JSON-lines event file per directory.

reference: ``tf.summary.FileWriter(eval_dir)`` with the two evaluator scalars
``Validation Accuracy`` / ``Validation Loss`` (src/nn_eval.py:107-110,133-134) and
the chief's (empty) merged training summary (src/distributed_train.py:225,382-390).
TensorBoard's event-file protobuf is not reproduced; tags, steps and wall times
are kept so tools/benchmark.py-style plots can be drawn from the file.
"""
from __future__ import annotations

import json
import os
import time
from typing import Dict


class SummaryWriter:
    def __init__(self, logdir: str, filename: str = "events.out.dmnist.jsonl"):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, filename)
        self._f = open(self.path, "a", buffering=1)

    def add_scalars(self, scalars: Dict[str, float], global_step: int) -> None:
        self._f.write(json.dumps({"wall_time": time.time(), "step": int(global_step),
                                  "scalars": {k: float(v) for k, v in scalars.items()}}) + "\n")

    def close(self) -> None:
        self._f.close()


def read_events(path: str):
    with open(path) as f:
        return [json.loads(line) for line in f if line.strip()]
