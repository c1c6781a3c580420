"""Symmetric memory: identical buffers on every rank, each mapped into every peer.

Two allocators, same interface (``local_ptr``, ``peer_ptrs``, ``multicast_ptr``, ``view``, ``ptr_table``):

* :class:`SymmetricBuffer` -- the kernel library's own (``csrc/symm_mem.cu``: ``cudaMalloc`` + CUDA IPC, handles
  exchanged over the process group).  Always available; peer (P2P) addresses only.
* :class:`MulticastBuffer` -- CUDA VMM allocations bound to an **NVLS multicast object** (NVSwitch), obtained through
  ``torch.distributed._symmetric_memory`` (allocation/rendezvous plumbing only: no collective of that module is used).
  Besides the peer addresses it yields ``multicast_ptr``: one virtual address whose loads are reduced *inside the
  switch* (``multimem.ld_reduce``) and whose stores are replicated to every rank (``multimem.st``) -- what the fused
  aggregation kernel uses when every replica contributes.

Both are exposed as ordinary ``torch`` tensors (zero-copy) plus the table of peer device pointers the kernels take.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from ..ops.lib import check, load

_TYPESTR = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4", torch.bfloat16: None, torch.int64: "<i8",
            torch.uint16: "<u2"}


class _CudaArray:
    """Minimal ``__cuda_array_interface__`` carrier."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self._owner = owner


class SymmetricBuffer:
    """One symmetric allocation of ``nbytes`` (zero-initialised) on every rank."""

    def __init__(self, nbytes: int, rank: int, world_size: int, device: torch.device, group=None):
        self.lib = load()
        self.nbytes = (nbytes + 255) // 256 * 256
        self.rank, self.world_size, self.device = rank, world_size, device
        p = ctypes.c_void_p()
        check(self.lib.dm_symm_alloc(ctypes.c_ulonglong(self.nbytes), ctypes.byref(p)), "dm_symm_alloc")
        self.local_ptr = int(p.value)
        self.peer_ptrs: List[int] = [0] * world_size
        self.peer_ptrs[rank] = self.local_ptr
        self._opened: List[int] = []
        if world_size > 1:
            h = ctypes.create_string_buffer(64)
            check(self.lib.dm_symm_ipc_handle(ctypes.c_void_p(self.local_ptr), h), "dm_symm_ipc_handle")
            handles: List[Optional[bytes]] = [None] * world_size
            dist.all_gather_object(handles, bytes(h.raw), group=group)
            for r, hb in enumerate(handles):
                if r == rank:
                    continue
                q = ctypes.c_void_p()
                check(self.lib.dm_symm_ipc_open(ctypes.create_string_buffer(hb, 64), ctypes.byref(q)),
                      "dm_symm_ipc_open(rank %d)" % r)
                self.peer_ptrs[r] = int(q.value)
                self._opened.append(int(q.value))
        self.multicast_ptr = 0
        self._bytes = torch.as_tensor(_CudaArray(self.local_ptr, self.nbytes, self), device=device)

    def view(self, dtype: torch.dtype, offset_bytes: int = 0, numel: Optional[int] = None) -> torch.Tensor:
        item = torch.empty((), dtype=dtype).element_size()
        if numel is None:
            numel = (self.nbytes - offset_bytes) // item
        return self._bytes[offset_bytes:offset_bytes + numel * item].view(dtype)

    def ptr_table(self, offset_bytes: int = 0):
        """``(c_void_p * world)`` of peer addresses (index = rank) for the kernel launchers."""
        arr = (ctypes.c_void_p * self.world_size)()
        for r, p in enumerate(self.peer_ptrs):
            arr[r] = p + offset_bytes
        return arr

    def close(self) -> None:
        for p in self._opened:
            self.lib.dm_symm_ipc_close(ctypes.c_void_p(p))
        self._opened = []


class MulticastBuffer:
    """Symmetric allocation with an NVLS multicast mapping (``multicast_ptr`` is 0 when the fabric has none)."""

    def __init__(self, nbytes: int, rank: int, world_size: int, device: torch.device, group=None):
        import warnings

        import torch.distributed._symmetric_memory as tsm
        self.nbytes = (nbytes + 255) // 256 * 256
        self.rank, self.world_size, self.device = rank, world_size, device
        group = group if group is not None else dist.group.WORLD
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                tsm.enable_symm_mem_for_group(group.group_name)       # no-op on current PyTorch
            except Exception:  # noqa: BLE001
                pass
        self._t = tsm.empty(self.nbytes, dtype=torch.uint8, device=device)
        self._t.zero_()
        torch.cuda.synchronize(device)
        self._hdl = tsm.rendezvous(self._t, group)
        self.local_ptr = int(self._t.data_ptr())
        self.peer_ptrs: List[int] = [int(p) for p in self._hdl.buffer_ptrs]
        assert self.peer_ptrs[rank] == self.local_ptr, "symmetric memory: local pointer mismatch"
        mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0)
        self.multicast_ptr = mc
        self._bytes = self._t

    view = SymmetricBuffer.view
    ptr_table = SymmetricBuffer.ptr_table

    def close(self) -> None:
        self._hdl = None


def allocate_symmetric(nbytes: int, rank: int, world_size: int, device: torch.device, want_multicast: bool):
    """Arena allocator used by the fused backend: NVLS-capable when asked for and available, IPC otherwise."""
    if want_multicast and world_size > 1:
        try:
            buf = MulticastBuffer(nbytes, rank, world_size, device)
            return buf
        except Exception as e:  # noqa: BLE001 -- no NVSwitch / VMM: P2P still works
            import sys
            print("[dmnist] multicast symmetric memory unavailable (%s: %s); using CUDA-IPC peer mappings"
                  % (type(e).__name__, str(e).splitlines()[0][:200]), file=sys.stderr)
    return SymmetricBuffer(nbytes, rank, world_size, device)
