"""FusedBackend: gradient aggregation by the single sm_100a kernel of ``csrc/fused_sync.cu``.

The product path for SURVEY §2.5 X1-X8: the parameter arena, the gradient arena and a
control block live in symmetric memory; one kernel per step publishes arrival, obtains
the commit mask (all-to-all flags for K == N, chief-owned bitmap + CAS for K < N), sums
the contributors' gradients straight out of peer HBM over NVLink, scales by
``lr / popcount(mask)``, applies SGD on its shard and pushes the new weights into every
rank's arena, then refreshes the local bf16 shadow used by the tensor-core kernels.
No NCCL collective and no separate elementwise kernel is launched.

Two ways to drive it:

* ``sync_step`` -- eager, host-visible result (tests, the reference-style training loop);
* ``enqueue`` -- stream-ordered launch with the LR schedule evaluated on the device from
  the device-resident step counter, so the whole training step replays as a CUDA graph
  with no per-step host input.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

import torch

from ..ops.lib import check, load, require_blackwell, stream_ptr
from .backends import Backend, StepInfo
from .context import ReplicaContext
from .symm_mem import SymmetricBuffer, allocate_symmetric


class FusedBackend(Backend):
    name = "fused"

    def __init__(self, ctx: ReplicaContext, ctas: int = 0, timeout_ms: float = 30000.0):
        super().__init__(ctx)
        assert ctx.on_gpu, "FusedBackend needs a GPU"
        self.lib = load()
        require_blackwell(ctx.device)
        check(self.lib.dm_set_device(ctx.device.index or 0), "dm_set_device")
        # programmatic dependent launch between the kernels of a step (DMNIST_PDL=0 disables)
        self.lib.dm_set_pdl(0 if os.environ.get("DMNIST_PDL", "1") == "0" else 1)
        self.ctas = ctas or int(os.environ.get("DMNIST_SYNC_CTAS", "148"))
        # NVLS (in-switch reduction + multicast store) for the arenas when the fabric offers it; DMNIST_NVLS=0 -> P2P only
        #   DMNIST_NVLS unset: allocate multicast-capable arenas, use multimem from 4 replicas up (at 2 the peer path is
        #   faster: measured 106 vs 115 us/step, profiles/bench_r1_call23_2gpu.txt); =1 always; =0 never (CUDA-IPC arenas)
        nv = os.environ.get("DMNIST_NVLS", "")
        self.want_nvls = nv != "0"
        self.use_nvls = nv == "1" or (nv == "" and ctx.world_size >= 4)
        self.timeout_ms = timeout_ms
        self._buffers: List[SymmetricBuffer] = []
        self._by_ptr: Dict[int, SymmetricBuffer] = {}
        nbytes = int(self.lib.dm_sync_ctrl_bytes())
        self.ctrl = SymmetricBuffer(nbytes, ctx.rank, ctx.world_size, ctx.device)
        self._ctrl_bytes = self.ctrl.view(torch.uint8)
        self._off = {f: int(self.lib.dm_sync_ctrl_offset(f.encode())) for f in
                     ("epoch", "error", "accepted_steps", "dropped_steps", "last_mask", "last_count", "last_late",
                      "global_step", "t_arrive", "t_start", "cta_counter", "t_phase", "t_phase_e")}
        self.shadow: Optional[torch.Tensor] = None     # bf16 copy of the parameter arena
        self.drop_keep = 0.0
        self.drop_seed = 0

    # ---- memory ---------------------------------------------------------------------------
    def allocate(self, numel: int) -> torch.Tensor:
        buf = allocate_symmetric(numel * 4, self.ctx.rank, self.ctx.world_size, self.ctx.device, self.want_nvls)
        self._buffers.append(buf)
        t = buf.view(torch.float32, 0, numel)
        self._by_ptr[t.data_ptr()] = buf
        return t

    def attach_shadow(self, params: torch.Tensor) -> torch.Tensor:
        """Create the local bf16 shadow arena of ``params`` and fill it."""
        self.shadow = torch.empty(params.numel(), dtype=torch.bfloat16, device=params.device)
        self.refresh_shadow(params)
        return self.shadow

    def refresh_shadow(self, params: torch.Tensor) -> None:
        if self.shadow is not None:
            check(self.lib.dm_f32_to_bf16(ctypes.c_void_p(params.data_ptr()), ctypes.c_void_p(self.shadow.data_ptr()),
                                          ctypes.c_longlong(params.numel()), stream_ptr()), "dm_f32_to_bf16")

    # ---- control-block access ---------------------------------------------------------------
    def _read_u32(self, field: str, index: int = 0) -> int:
        off = self._off[field] + 4 * index
        return int(self._ctrl_bytes[off:off + 4].view(torch.int32).item()) & 0xFFFFFFFF

    def _write_u32(self, field: str, value: int) -> None:
        off = self._off[field]
        self._ctrl_bytes[off:off + 4].view(torch.int32).fill_(value if value < 2 ** 31 else value - 2 ** 32)

    @property
    def device_epoch(self) -> int:
        return self._read_u32("epoch")

    def set_global_step(self, step: int) -> None:
        """Checkpoint restore: every rank starts from ``step`` (also the chief's commit counter)."""
        self._write_u32("epoch", step)
        self._write_u32("global_step", step)
        torch.cuda.synchronize()
        self.barrier()

    def check_error(self) -> None:
        e = self._read_u32("error")
        if e:
            raise RuntimeError("fused sync watchdog fired on rank %d: %s timeout"
                               % (self.ctx.rank, {1: "arrival", 2: "push-complete"}.get(e, str(e))))

    def read_phases(self):
        """%globaltimer stamps of the last launch: start, decided, reduced, pushed, landed, end (ns, relative)."""
        t = self._ctrl_bytes[self._off["t_phase"]:self._off["t_phase"] + 48].view(torch.int64).cpu().tolist()
        return [x - t[0] for x in t]

    def read_phases_early(self):
        """Early-bucket kernel of the last step: start, arrived, reduced+pushed (CTA 0), all pushes out (ns, relative)."""
        t = self._ctrl_bytes[self._off["t_phase_e"]:self._off["t_phase_e"] + 32].view(torch.int64).cpu().tolist()
        return [x - t[0] for x in t]

    def read_timing(self, first_step: int, last_step: int):
        """(start_ns, arrive_ns) %globaltimer stamps of local steps [first, last] (cdf telemetry)."""
        ring = 1024
        ta = self._ctrl_bytes[self._off["t_arrive"]:self._off["t_arrive"] + 8 * ring].view(torch.int64).cpu()
        ts = self._ctrl_bytes[self._off["t_start"]:self._off["t_start"] + 8 * ring].view(torch.int64).cpu()
        return [(int(ts[s % ring]), int(ta[s % ring])) for s in range(first_step, last_step + 1)]

    # ---- launches -------------------------------------------------------------------------------
    def enqueue(self, params: torch.Tensor, grads: torch.Tensor, k: int, lr0: float, decay_rate: float = 1.0,
                decay_steps: int = 1, stream: Optional[torch.cuda.Stream] = None) -> None:
        pb, gb = self._by_ptr[params.data_ptr()], self._by_ptr[grads.data_ptr()]
        rc = self.lib.dm_fused_sync_sgd(
            self.ctrl.ptr_table(), pb.ptr_table(), gb.ptr_table(), self.ctx.rank, self.ctx.world_size, int(k),
            ctypes.c_longlong(params.numel()), ctypes.c_float(lr0), ctypes.c_float(decay_rate), int(decay_steps),
            ctypes.c_float(self.drop_keep), ctypes.c_uint(self.drop_seed), ctypes.c_double(self.timeout_ms),
            ctypes.c_void_p(0 if self.shadow is None else self.shadow.data_ptr()), int(self.ctas), stream_ptr(stream),
            *self._mc_ptrs(pb, gb))
        check(rc, "dm_fused_sync_sgd")

    def _mc_ptrs(self, pb, gb):
        on = self.use_nvls and pb.multicast_ptr and gb.multicast_ptr
        return ctypes.c_void_p(gb.multicast_ptr if on else 0), ctypes.c_void_p(pb.multicast_ptr if on else 0)

    def enqueue_bucket(self, params: torch.Tensor, grads: torch.Tensor, phase: int, begin: int, end: int,
                       early_begin: int, early_end: int, lr0: float, decay_rate: float = 1.0, decay_steps: int = 1,
                       ctas: int = 0, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Bucketed aggregation (K == N, world_size > 1): ``phase`` 1 = early bucket ``[begin, end)`` (floats) on a few
        CTAs next to the remaining backward kernels, 2 = late bucket + completion of the step (csrc/fused_sync.cu)."""
        pb, gb = self._by_ptr[params.data_ptr()], self._by_ptr[grads.data_ptr()]
        rc = self.lib.dm_fused_sync_bucket(
            self.ctrl.ptr_table(), pb.ptr_table(), gb.ptr_table(), self.ctx.rank, self.ctx.world_size, int(phase),
            ctypes.c_longlong(begin), ctypes.c_longlong(end), ctypes.c_longlong(early_begin), ctypes.c_longlong(early_end),
            ctypes.c_float(lr0), ctypes.c_float(decay_rate), int(decay_steps), ctypes.c_double(self.timeout_ms),
            ctypes.c_void_p(0 if self.shadow is None else self.shadow.data_ptr()), int(ctas), stream_ptr(stream),
            *self._mc_ptrs(pb, gb))
        check(rc, "dm_fused_sync_bucket")

    @property
    def nvls_active(self) -> bool:
        return self.use_nvls and bool(self._buffers) and all(b.multicast_ptr for b in self._buffers[:2])

    def enqueue_straggler_delay(self, prob: float, usec: float, seed: int = 12345,
                                stream: Optional[torch.cuda.Stream] = None) -> None:
        check(self.lib.dm_straggler_delay(ctypes.c_void_p(self.ctrl.local_ptr), ctypes.c_float(prob),
                                          ctypes.c_uint(int(usec)), ctypes.c_uint(seed), stream_ptr(stream)),
              "dm_straggler_delay")

    def enqueue_stamp_start(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        check(self.lib.dm_stamp_start(ctypes.c_void_p(self.ctrl.local_ptr), stream_ptr(stream)), "dm_stamp_start")

    def drop_connect_(self, grads: torch.Tensor, keep_prob: float, step: int) -> None:
        # Applied inside the fused kernel's load stage (per contributor); just arm it.
        self.drop_keep = float(keep_prob)
        self.drop_seed = 0x2545F491

    def last_step_info(self) -> StepInfo:
        late = self._read_u32("last_late")
        return StepInfo(global_step=self._read_u32("epoch"), accepted=not late, mask=self._read_u32("last_mask"),
                        count=self._read_u32("last_count"), stale=bool(late))

    def sync_step(self, params, grads, lr, local_step, k, delay_s: float = 0.0) -> StepInfo:
        if delay_s > 0:
            self.enqueue_straggler_delay(1.0, delay_s * 1e6)
        self.enqueue(params, grads, k, lr0=float(lr))
        info = self.last_step_info()   # device -> host read (synchronises)
        self.check_error()
        return info

    def interval_tick(self, params, acc, count, lr, tick) -> StepInfo:
        """Mode C tick: every rank joins; ranks with an empty accumulator contribute zeros and the
        divisor is the total number of accumulated gradients (mean of whatever arrived)."""
        import torch.distributed as dist
        counts = self.all_gather_object(int(count))
        total = sum(counts)
        mask = sum(1 << i for i, c in enumerate(counts) if c > 0)
        if total == 0:
            return StepInfo(self.device_epoch, False, 0, 0, applied=False)
        # lr/total instead of lr/popcount: fold the ratio into lr0 (kernel divides by N contributors)
        self.enqueue(params, acc, self.ctx.world_size, lr0=float(lr) * self.ctx.world_size / total)
        info = self.last_step_info()
        self.check_error()
        return StepInfo(info.global_step, count > 0, mask, total)

    def close(self) -> None:
        torch.cuda.synchronize()
        self.barrier()
        for b in self._buffers + [self.ctrl]:
            b.close()
