"""Peer-mapped gradient exchange between the GPUs of one NVSwitch box (C ABI: ``grpo_peer_*`` / ``grpo_ipc_*``).

What it replaces: the fp32 gradient averaging FSDP performs for the reference's actor (``mp_reduce_dtype``,
verl/workers/actor/config.py:58; verl/workers/fsdp_workers.py:242-280) and the passes of ``_optimizer_step`` that follow it
(verl/workers/actor/dp_actor.py:155-167) - for the replicated lm_head weight whose fp32 ``dW`` the fused head accumulates.

One process per GPU. ``torch.distributed`` only carries the 64-byte CUDA IPC handles at set-up (``all_gather_object``);
after that every exchange is this library's own kernels loading from / storing to the other GPUs' memory over NVLink,
ordered by flag barriers on the caller's CUDA stream (no host synchronisation, no NCCL call):

    barrier -> reduce-scatter (+ sum of squares of the slab) -> barrier -> norm / clip coefficient (caller)
            -> scale + cast to bf16 + all-gather + zero of the accumulator -> barrier

An all-reduce moves 2 (W-1)/W of the buffer per direction per GPU however it is issued; this sequence sends the second
half in the dtype the optimizer consumes (bf16, the parameter's dtype): 0.75x the wire bytes, and the norm / clip / cast /
zero passes over HBM ride along. Results are bit-identical on every rank and from run to run.

There is no CPU path here: ``PeerGroup`` needs the CUDA library and peer-accessible GPUs and raises ``PeerUnavailable``
otherwise (the caller may then keep NCCL's all-reduce).
"""
from __future__ import annotations

import ctypes
import os
import socket
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

MAX_PEERS = 8
_FLAGS_OFFSET, _PARTIALS_OFFSET, _CONTROL_BYTES = 0, 64, 4096
DEFAULT_TIMEOUT_MS = 120000  # a rank that is this late is gone (NCCL's own watchdog waits ten minutes)


class PeerUnavailable(RuntimeError):
    """The ranks of the group cannot map each other's memory (several hosts, no peer access, IPC refused ...)."""


class PeerBuffer:
    """A tensor of this rank plus the mapped addresses of the same tensor on every rank (rank order)."""

    def __init__(self, tensor: torch.Tensor, ptrs: List[int], opened: List[Tuple[int, int]]):
        self.tensor = tensor
        self.ptrs = ptrs
        self.table = (ctypes.c_void_p * MAX_PEERS)(*(ptrs + [0] * (MAX_PEERS - len(ptrs))))
        self._opened = opened  # (mapped pointer, offset) of the peers' copies, to close


def slab_bounds(n: int, rank: int, world: int, unit: int = 8) -> Tuple[int, int]:
    """Element range [e0, e1) of the slab rank ``rank`` reduces: units of ``unit`` elements, ceil(units / world) per rank
    (the partition the kernels use; host-side mirror for tests and tools)."""
    units = n // unit
    per = -(-units // world)
    u0 = min(per * rank, units)
    u1 = min(u0 + per, units)
    return u0 * unit, u1 * unit


class PeerGroup:
    """Control block (barrier flags + per-rank partial sums) shared by the ranks of ``group``, and the registry of
    peer-mapped buffers. Construction and ``register`` / ``release`` are COLLECTIVE over the group."""

    def __init__(self, group: Optional["dist.ProcessGroup"] = None, device: Optional[torch.device] = None,
                 timeout_ms: int = DEFAULT_TIMEOUT_MS):
        if not (dist.is_available() and dist.is_initialized()):
            raise PeerUnavailable("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.timeout_ms = int(os.environ.get("GRPO_PEER_TIMEOUT_MS", timeout_ms))
        self.lib = _lib.load()
        self.epoch = 0
        self.buffers: List[PeerBuffer] = []
        # every rank must be on this host and able to reach every other rank's GPU (two ranks on ONE GPU also work - CUDA
        # IPC maps between processes, the barrier then relies on the driver's time-slicing; tests do that). The verdict is
        # agreed over the group so that either every rank gets a PeerGroup or every rank raises
        here = (socket.gethostname(), self.device.index)
        places: List[Optional[tuple]] = [None] * self.world
        dist.all_gather_object(places, here, group=group)
        error = None
        if self.world > MAX_PEERS:
            error = f"{self.world} ranks: the peer path covers one box of up to {MAX_PEERS} GPUs"
        elif len({p[0] for p in places}) != 1:
            error = f"ranks are spread over several hosts: {places}"
        else:
            for _, index in places:
                if index != self.device.index and not torch.cuda.can_device_access_peer(self.device.index, index):
                    error = f"GPU {self.device.index} cannot access GPU {index}"
        errors: List[Optional[str]] = [None] * self.world
        dist.all_gather_object(errors, error, group=group)
        if any(errors):
            raise PeerUnavailable("; ".join(e for e in errors if e))
        self.control = torch.zeros(_CONTROL_BYTES, dtype=torch.uint8, device=self.device)
        self.scratch = torch.zeros(_lib.GRAD_SCRATCH_DOUBLES, dtype=torch.float64, device=self.device)
        torch.cuda.synchronize(self.device)  # the flags read zero before anybody can announce an epoch
        self._control = self.register(self.control)
        base = self._control.ptrs
        self._flags = (ctypes.c_void_p * MAX_PEERS)(*([p + _FLAGS_OFFSET for p in base] + [0] * (MAX_PEERS - self.world)))
        self._partials = (ctypes.c_void_p * MAX_PEERS)(*([p + _PARTIALS_OFFSET for p in base] + [0] * (MAX_PEERS - self.world)))
        # this rank's double[world] of slab sums of squares, written by the peers
        self.partials = self.control[_PARTIALS_OFFSET:_PARTIALS_OFFSET + 8 * self.world].view(torch.float64)

    # ------------------------------------------------------------------------------------------------------------
    def register(self, tensor: torch.Tensor) -> PeerBuffer:
        """Map ``tensor`` (contiguous, on this rank's device; same shape and dtype on every rank) into every rank.
        Collective. If any rank fails, every rank raises ``PeerUnavailable`` and nothing stays mapped."""
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.device == self.device
        handle = ctypes.create_string_buffer(64)
        offset = ctypes.c_int64(0)
        error = None
        try:
            _lib.check(self.lib.grpo_ipc_export(_lib.ptr(tensor), handle, ctypes.byref(offset)), "grpo_ipc_export")
        except Exception as exc:  # e.g. expandable_segments: the allocation has no IPC handle
            error = f"rank {self.rank}: {exc}"
        mine = (bytes(handle.raw), int(offset.value), tuple(tensor.shape), str(tensor.dtype), error)
        every: List[Optional[tuple]] = [None] * self.world
        dist.all_gather_object(every, mine, group=self.group)
        ptrs: List[int] = []
        opened: List[Tuple[int, int]] = []
        if error is None:
            try:
                for q, (h, off, shape, dtype, err) in enumerate(every):
                    if err is not None:
                        raise PeerUnavailable(err)
                    if shape != mine[2] or dtype != mine[3]:
                        raise PeerUnavailable(f"rank {q} registers {shape} {dtype}, rank {self.rank} {mine[2]} {mine[3]}")
                    if q == self.rank:
                        ptrs.append(_lib.ptr(tensor))
                        continue
                    out = ctypes.c_void_p(0)
                    _lib.check(self.lib.grpo_ipc_open(h, ctypes.c_int64(off), ctypes.byref(out)), "grpo_ipc_open")
                    ptrs.append(int(out.value))
                    opened.append((int(out.value), off))
            except Exception as exc:
                error = f"rank {self.rank}: {exc}"
        outcomes: List[Optional[str]] = [None] * self.world
        dist.all_gather_object(outcomes, error, group=self.group)  # also: nobody proceeds before everybody has mapped
        failed = [o for o in outcomes if o is not None]
        if failed:
            for p, off in opened:
                self.lib.grpo_ipc_close(ctypes.c_void_p(p), ctypes.c_int64(off))
            raise PeerUnavailable("; ".join(failed))
        buf = PeerBuffer(tensor, ptrs, opened)
        self.buffers.append(buf)
        return buf

    def release(self, buf: PeerBuffer) -> None:
        """Unmap the peers' copies of ``buf`` (collective: nobody unmaps while a kernel of another rank may still touch
        the memory)."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for p, off in buf._opened:
            _lib.check(self.lib.grpo_ipc_close(ctypes.c_void_p(p), ctypes.c_int64(off)), "grpo_ipc_close")
        buf._opened = []
        if buf in self.buffers:
            self.buffers.remove(buf)
        dist.barrier(group=self.group)  # the owner may free the memory only after every importer has closed it

    # ------------------------------------------------------------------------------------------------------------
    def barrier(self) -> None:
        """Stream-ordered barrier over the ranks (one 32-thread kernel): everything the ranks' streams did before it -
        stores into peer memory included - is visible to whatever any rank's stream does after it. Epochs are compared
        as signed 32-bit differences, so the counter may wrap."""
        self.epoch += 1
        _lib.check(self.lib.grpo_peer_barrier(self._flags, self.rank, self.world, ctypes.c_uint(self.epoch & 0xFFFFFFFF),
                                              self.timeout_ms, _lib.stream_ptr(self.device)), "grpo_peer_barrier")

    def reduce_scatter_sumsq(self, buf: PeerBuffer) -> torch.Tensor:
        """Mean over the ranks of slab ``rank`` of the fp32 buffer, left in this rank's copy; returns the sum of squares of
        the WHOLE averaged buffer (double[1], identical bits on every rank). The other slabs of the local copy keep this
        rank's own contribution until ``scale_cast_allgather`` zeroes them."""
        t = buf.tensor
        assert t.dtype == torch.float32 and t.numel() % 8 == 0
        self.barrier()  # every rank's accumulator is complete
        _lib.check(self.lib.grpo_peer_reduce_scatter_sumsq(buf.table, self._partials, self.rank, self.world, t.numel(),
                                                           _lib.ptr(self.scratch), _lib.stream_ptr(self.device)),
                   "grpo_peer_reduce_scatter_sumsq")
        self.barrier()  # every slab is reduced, every partial has landed, nobody reads the accumulators any more
        return self.partials.sum(dim=0, keepdim=True)

    def scale_cast_allgather(self, buf: PeerBuffer, out: PeerBuffer, scale: Optional[torch.Tensor] = None,
                             zero_after: bool = True) -> torch.Tensor:
        """``out`` (bf16, same element count) on EVERY rank <- bf16(scale * averaged buffer); ``zero_after`` clears the
        whole local accumulator in the same pass. Follows ``reduce_scatter_sumsq`` of the same ``buf``."""
        t, o = buf.tensor, out.tensor
        assert o.dtype == torch.bfloat16 and o.numel() == t.numel()
        if scale is not None:
            assert scale.is_cuda and scale.dtype == torch.float32 and scale.numel() == 1
        _lib.check(self.lib.grpo_peer_scale_cast_allgather(_lib.ptr(t), out.table, self.rank, self.world, t.numel(),
                                                           _lib.ptr(scale) if scale is not None else None, 1.0,
                                                           1 if zero_after else 0, _lib.stream_ptr(self.device)),
                   "grpo_peer_scale_cast_allgather")
        self.barrier()  # every rank's slab has landed in every gradient buffer
        return o

    def allreduce_mean_(self, buf: PeerBuffer) -> torch.Tensor:
        """General in-place fp32 mean all-reduce of a registered buffer (same result on every rank, rank-ordered sums)."""
        t = buf.tensor
        assert t.dtype == torch.float32 and t.numel() % 4 == 0
        self.barrier()
        _lib.check(self.lib.grpo_peer_allreduce_mean(buf.table, self.rank, self.world, t.numel(),
                                                     _lib.stream_ptr(self.device)), "grpo_peer_allreduce_mean")
        self.barrier()
        return t


_GROUPS: Dict[tuple, PeerGroup] = {}


def get_group(group: Optional["dist.ProcessGroup"] = None, device: Optional[torch.device] = None) -> PeerGroup:
    """The process-wide ``PeerGroup`` of (``group``, ``device``): one control block however many actors use it.
    Collective on first use."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    key = (id(group) if group is not None else None, dev.index)
    if key not in _GROUPS:
        _GROUPS[key] = PeerGroup(group, dev)
    return _GROUPS[key]
