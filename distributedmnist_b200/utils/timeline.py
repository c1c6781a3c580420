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

    def add_device_spans(self, spans: List[Dict], base_us: float = 0.0) -> None:
        """``spans`` = [{"name", "ts", "dur", "stream"}] (microseconds, ``ts`` relative to the first one): one span per kernel /
        copy of the step, one chrome "thread" per CUDA stream -- the per-op view of the reference's FULL_TRACE timeline."""
        streams = sorted({sp.get("stream", 0) for sp in spans}, key=str)
        for sp in spans:
            self.add_span(sp["name"], base_us + sp["ts"], sp["dur"], tid=100 + streams.index(sp.get("stream", 0)), cat="Kernel",
                          args={"stream": sp.get("stream", 0)})
        self._streams = streams

    def generate_chrome_trace_format(self) -> str:
        meta = [{"name": "process_name", "ph": "M", "pid": self.pid, "args": {"name": "worker %d" % self.pid}},
                {"name": "thread_name", "ph": "M", "pid": self.pid, "tid": 0, "args": {"name": "host"}},
                {"name": "thread_name", "ph": "M", "pid": self.pid, "tid": 1, "args": {"name": "cuda stream"}}]
        for i, st in enumerate(getattr(self, "_streams", [])):
            meta.append({"name": "thread_name", "ph": "M", "pid": self.pid, "tid": 100 + i, "args": {"name": "cuda stream %s" % st}})
        return json.dumps({"traceEvents": meta + self._events}, indent=1)


def cupti_device_spans(fn) -> List[Dict]:
    """Run ``fn`` (enqueue work, e.g. one graph-replayed training step) under CUPTI (``torch.profiler``) and return every
    kernel / memcpy / memset it executed on the device as ``{"name", "ts", "dur", "stream"}`` (us, relative to the first)."""
    import tempfile

    import torch
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    with tempfile.NamedTemporaryFile(suffix=".json", delete=False) as f:
        path = f.name
    try:
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"]
              if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    finally:
        try:
            os.remove(path)
        except OSError:
            pass
    ev.sort(key=lambda e: e["ts"])
    if not ev:
        return []
    t0 = ev[0]["ts"]
    return [{"name": e["name"].split("(")[0].replace("void ", ""), "ts": e["ts"] - t0, "dur": e["dur"],
             "stream": e.get("args", {}).get("stream", 0)} for e in ev]


def timeline_path(train_dir: str, worker_id: int, step: int) -> str:
    return os.path.join(train_dir, "worker=%d_timeline_iter=%d.json" % (worker_id, step))
