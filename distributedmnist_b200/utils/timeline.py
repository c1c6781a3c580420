"""Per-step chrome-trace timeline (``--timeline_logging``).

reference: ``RunOptions.trace_level=FULL_TRACE`` + ``timeline.Timeline(step_stats)
.generate_chrome_trace_format()`` written to
``train_dir/worker=<id>_timeline_iter=<step>.json`` (src/distributed_train.py:317-319,
354-358).  Here spans come from host timestamps and, on the GPU path, from CUDA
events recorded around each kernel of the step; the file name pattern and the
chrome ``traceEvents`` JSON format are the same, so chrome://tracing / Perfetto
open them unchanged.
"""
from __future__ import annotations

import json
import os
import time
from contextlib import contextmanager
from typing import Dict, List, Optional


class Timeline:
    def __init__(self, pid: int = 0):
        self.pid = pid
        self._events: List[Dict] = []
        self._t0 = time.perf_counter()

    def add_span(self, name: str, start_us: float, dur_us: float, tid: int = 0, cat: str = "Op",
                 args: Optional[Dict] = None) -> None:
        self._events.append({"name": name, "cat": cat, "ph": "X", "ts": float(start_us), "dur": float(dur_us),
                             "pid": self.pid, "tid": tid, "args": args or {}})

    @contextmanager
    def span(self, name: str, tid: int = 0, cat: str = "Op"):
        t = time.perf_counter()
        try:
            yield
        finally:
            t1 = time.perf_counter()
            self.add_span(name, (t - self._t0) * 1e6, (t1 - t) * 1e6, tid, cat)

    def add_cuda_spans(self, names: List[str], events: List, base_us: float = 0.0, tid: int = 1) -> None:
        """``events`` = [e0, e1, ... eN] recorded between kernels; span i = (e_i, e_{i+1})."""
        t = base_us
        for i, name in enumerate(names):
            dur = events[i].elapsed_time(events[i + 1]) * 1e3
            self.add_span(name, t, dur, tid=tid, cat="Kernel")
            t += dur

    def generate_chrome_trace_format(self) -> str:
        meta = [{"name": "process_name", "ph": "M", "pid": self.pid, "args": {"name": "worker %d" % self.pid}},
                {"name": "thread_name", "ph": "M", "pid": self.pid, "tid": 0, "args": {"name": "host"}},
                {"name": "thread_name", "ph": "M", "pid": self.pid, "tid": 1, "args": {"name": "cuda stream"}}]
        return json.dumps({"traceEvents": meta + self._events}, indent=1)


def timeline_path(train_dir: str, worker_id: int, step: int) -> str:
    return os.path.join(train_dir, "worker=%d_timeline_iter=%d.json" % (worker_id, step))
