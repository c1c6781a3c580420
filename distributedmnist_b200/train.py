"""Training driver: the per-replica hot loop.

reference: ``distributed_train.train(target, dataset, dataset_test, cluster_spec)``
(src/distributed_train.py:109-408).  Behaviour kept: K defaults to the number of
workers (:118-121); replica 0 is the chief (:130); staircase LR decay with the /K
rule (:143-156); mode selection from ``--interval_method`` /
``--worker_times_cdf_method`` (:176-188); optional drop-connect in the default mode
only (:194-203); chief restores the latest checkpoint at start (Supervisor,
:244-262) and saves every ``--save_interval_secs`` plus once at the end (:252,
405-408); start-up barrier (:277); the log-line formats (:296,325,341-342,367-371,
398-399); ``worker<id>_time_acc.npy`` every ``--save_results_period`` global steps
(:373-379); per-step chrome timelines (:354-358); loop exit when
``step > max_steps`` checked after the work (:360).

Dropped on purpose: the second forward pass per iteration that the
reference runs only to log loss/accuracy (:334), ``sess.kill()/reset_kill()``
(:391-396, not in stock TF), ``np.set_printoptions(threshold=np.nan)`` (:31).
"""
from __future__ import annotations

import os
import threading
import time
from datetime import datetime
from typing import Dict, Optional

import numpy as np
import torch

from .checkpoint import Saver
from .engine import make_engine
from .flags import FLAGS
from .parallel.aggregators import (SyncReplicasOptimizer, TimeoutReplicasOptimizer,
                                   parse_straggler_spec)
from .parallel.backends import make_backend
from .parallel.context import ReplicaContext
from .parallel.timeout_manager import launch_manager
from .schedule import LearningRateSchedule
from .utils.logging import get_logger
from .utils.summary import SummaryWriter
from .utils.timeline import Timeline, timeline_path

log = get_logger()

LOG_FORMAT = ("Worker %d: %s: step %d, loss = %f, train_acc = %f, test_acc = %f"
              "(%.1f examples/sec; %.3f  sec/batch)")


class _AsyncCheckpointer:
    """Chief-side saver: snapshot on the training thread, file IO on a helper thread."""

    def __init__(self, train_dir: str, spec, interval_s: float):
        self.train_dir, self.spec, self.interval_s = train_dir, spec, interval_s
        self.saver = Saver()
        self._next = time.time() + interval_s
        self._thread: Optional[threading.Thread] = None

    def maybe_save(self, params: torch.Tensor, global_step: int) -> None:
        if time.time() >= self._next:
            self._next = time.time() + self.interval_s
            self.save(params, global_step, blocking=False)

    def save(self, params: torch.Tensor, global_step: int, blocking: bool = True) -> None:
        state = self.spec.to_state_dict(params)  # consistent snapshot (device -> host copy)
        if self._thread is not None:
            self._thread.join()
        self._thread = threading.Thread(target=self.saver.save, args=(self.train_dir, state, global_step))
        self._thread.start()
        if blocking:
            self._thread.join()

    def close(self) -> None:
        if self._thread is not None:
            self._thread.join()


def train(ctx: ReplicaContext, dataset, dataset_test=None, flags=FLAGS) -> Dict:
    num_workers = ctx.world_size
    # reference distributed_train.py:118-121
    if flags.num_replicas_to_aggregate == -1:
        num_replicas_to_aggregate = num_workers
    else:
        num_replicas_to_aggregate = flags.num_replicas_to_aggregate
    assert num_workers > 0
    is_chief = ctx.is_chief

    backend = make_backend(ctx, flags.backend)
    engine = make_engine(flags, ctx, backend)
    spec = engine.spec
    lr_schedule = LearningRateSchedule.from_flags(flags, dataset.num_examples, num_replicas_to_aggregate)
    straggler = parse_straggler_spec(flags.inject_straggler)

    if flags.interval_method or flags.worker_times_cdf_method:
        # reference :179-183 -- interval wins when both are set, as there.
        mode = "interval" if flags.interval_method else "cdf"
        opt = TimeoutReplicasOptimizer(backend, lr_schedule, total_num_replicas=num_workers, mode=mode,
                                       interval_ms=flags.interval_ms, straggler=straggler, seed=flags.seed)
    else:
        opt = SyncReplicasOptimizer(
            backend, lr_schedule, replicas_to_aggregate=num_replicas_to_aggregate,
            total_num_replicas=num_workers,
            drop_connect_probability=flags.drop_connect_probability if flags.drop_connect else None,
            straggler=straggler, seed=flags.seed)
    fused_step = hasattr(engine, "attach_optimizer") and not flags.interval_method
    if fused_step:
        engine.attach_optimizer(opt)   # GPU path: compute + fused allreduce/SGD replay as one CUDA graph

    # Supervisor semantics: chief restores the newest checkpoint, everyone gets it.
    restored_step = 0
    if is_chief:
        os.makedirs(flags.train_dir, exist_ok=True)
        latest = Saver.latest(flags.train_dir)
        if latest is not None:
            state, restored_step = Saver.restore(latest)
            engine.params.copy_(spec.from_state_dict(state).to(engine.params.device))
            log.info("Restored model from %s at step=%d" % (latest, restored_step))
    if num_workers > 1:
        restored_step = backend.all_gather_object(restored_step)[0]
        if restored_step > 0:
            backend.broadcast_(engine.params, 0)
    if restored_step > 0:
        opt.local_step = restored_step
        if hasattr(backend, "set_global_step"):
            backend.set_global_step(restored_step)
    if hasattr(engine, "params_updated"):
        engine.params_updated()
    log.info("%s Supervisor" % datetime.now())

    ckpt = _AsyncCheckpointer(flags.train_dir, spec, flags.save_interval_secs) if is_chief else None
    summary = SummaryWriter(flags.train_dir) if (is_chief and flags.should_summarize) else None

    # Even if not using timeout, we want to wait until all machines are ready (reference :275-277).
    timeout_client, timeout_server = launch_manager(backend, flags)

    next_summary_time = time.time() + flags.save_summaries_secs
    begin_time = time.time()
    cur_iteration = -1
    if flags.interval_method:
        opt.start_interval_updates()   # reference :287-288 (chief timer) -> shared absolute deadlines

    time_acc_list = []
    step = restored_step
    loss_value = train_acc_value = float("nan")
    results = {"steps": [], "losses": [], "accepted": 0, "dropped": 0}

    while True:
        log.info("A new iteration...")
        cur_iteration += 1

        if flags.worker_times_cdf_method:
            t_deq = opt.wait_op()
            timeout_client.broadcast_worker_dequeued_token(cur_iteration, t_deq)

        tl = Timeline(pid=ctx.rank) if flags.timeline_logging else None
        start_time = time.time()
        images, labels = dataset.next_batch(flags.batch_size)
        if tl:
            tl.add_span("next_batch", 0.0, (time.time() - start_time) * 1e6)
        log.info("RUNNING SESSION... %f" % time.time())

        t0 = time.perf_counter()
        engine.load_batch(images, labels)
        t1 = time.perf_counter()
        if fused_step:
            engine.train_step()
            t2 = time.perf_counter()
            info = opt._account(engine.step_info())   # device -> host read of the step's outcome
        else:
            engine.forward_backward(opt.local_step)
            t2 = time.perf_counter()
            info = opt.apply_gradients(engine.params, engine.grads) if not isinstance(opt, TimeoutReplicasOptimizer) \
                else opt.apply_gradients(engine.params, engine.grads, ctx.rank, flags.worker_times_cdf_method)
            if info.applied and hasattr(engine, "params_updated") and not hasattr(engine, "train_step"):
                engine.params_updated()
        do_log = (cur_iteration % max(flags.log_every, 1) == 0)
        if do_log or flags.timeline_logging:
            loss_value, train_acc_value = engine.loss_acc()
        t3 = time.perf_counter()
        step = info.global_step
        log.info("Global step attained: %d" % step)
        log.info("DONE RUNNING SESSION...")

        if flags.worker_times_cdf_method:
            timeout_client.broadcast_worker_finished_computing_gradients(cur_iteration, opt.mark_finished())

        finish_time = time.time()

        if tl:
            base = (t0 - tl._t0) * 1e6
            tl.add_span("load_batch(H2D)", base, (t1 - t0) * 1e6)
            tl.add_span("forward_backward", base + (t1 - t0) * 1e6, (t2 - t1) * 1e6)
            tl.add_span("aggregate+apply", base + (t2 - t0) * 1e6, (t3 - t2) * 1e6)
            if hasattr(engine, "kernel_spans"):
                engine.kernel_spans(tl)
            with open(timeline_path(flags.train_dir, ctx.rank, step), "w") as f:
                f.write(tl.generate_chrome_trace_format())

        if step > flags.max_steps:
            break

        test_acc_value = 0.0   # hard-wired in the reference too (:363)
        duration = finish_time - start_time
        examples_per_sec = flags.batch_size / float(duration)
        if do_log:
            log.info(LOG_FORMAT % (ctx.rank, datetime.now(), step, loss_value, train_acc_value,
                                   test_acc_value, examples_per_sec, duration))
        time_acc_list.append((finish_time, train_acc_value, test_acc_value, loss_value))
        results["steps"].append(step)
        results["losses"].append(loss_value)

        if step % flags.save_results_period == 0:
            np.save(os.path.join(flags.train_dir, "worker%d_time_acc.npy" % ctx.rank), np.array(time_acc_list))

        if is_chief:
            ckpt.maybe_save(engine.params, step)
            if summary is not None and next_summary_time < time.time():
                log.info("Running Summary operation on the chief.")
                summary.add_scalars({"loss": loss_value, "train_acc": train_acc_value,
                                     "learning_rate": lr_schedule(step)}, step)
                log.info("Finished running Summary operation.")
                next_summary_time += flags.save_summaries_secs

    if is_chief:
        log.info("Elapsed Time: %f" % (time.time() - begin_time))
    if flags.worker_times_cdf_method:
        timeout_client.flush()

    # Save after the training ends (reference :405-408).
    if is_chief:
        ckpt.save(engine.params, step, blocking=True)
        ckpt.close()
    if summary is not None:
        summary.close()
    results.update(final_step=step, accepted=opt.accepted_steps, dropped=opt.dropped_steps,
                   params=engine.params, elapsed=time.time() - begin_time,
                   timeout_server=timeout_server, engine=engine, optimizer=opt)
    return results
