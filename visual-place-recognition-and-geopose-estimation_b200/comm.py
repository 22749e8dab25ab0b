"""Peer-memory communicator for the gradient all-reduce of data-parallel fine-tuning.

``PeerComm`` owns one symmetric fp32 allocation per rank -- ``[ data | staging | flags ]`` -- whose addresses every
rank of the node knows, and exposes ``allreduce_()``: the library's own two-shot sum all-reduce kernel over NVLink
(``vpr_allreduce_sum_f32``, csrc/comm.cu) instead of ``ncclAllReduce``.  The gradient buffers of the fine-tuning step
are *views of the data area*, so the kernels that produce the gradients write straight into memory the peers can
reach and nothing is copied before or after the exchange.

Plumbing (PyTorch, not the product): the peer mappings come from ``torch.distributed._symmetric_memory`` when it is
available (it also binds an NVLS multicast address when the fabric supports one) and otherwise from CUDA IPC handles
exchanged through the process group (``UntypedStorage._share_cuda_``, what ``torch.multiprocessing`` uses).
``torch.distributed`` is used for the rendezvous only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib as L


class PeerComm:
    """``PeerComm(numel)``: a symmetric fp32 buffer of ``numel`` elements (``.data``, a local tensor to carve gradient
    buffers from) that can be sum-all-reduced in place across the ranks of ``group`` (one node, <= 8 GPUs)."""

    def __init__(self, numel: int, group=None, device=None, backend: str | None = None):
        if not dist.is_initialized():
            raise RuntimeError("PeerComm needs an initialised process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > L.COMM_MAX_WORLD:
            raise RuntimeError(f"PeerComm supports up to {L.COMM_MAX_WORLD} ranks on one node")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        lib = L.lib()
        self.numel = (int(numel) + 3) // 4 * 4
        self.staging_elems = int(lib.vpr_comm_staging_elems(self.numel, self.world))
        self.flag_elems = int(lib.vpr_comm_flag_bytes()) // 4
        # every area starts on a 256-byte boundary
        self._off_staging = (self.numel + 63) // 64 * 64
        self._off_flags = self._off_staging + (self.staging_elems + 63) // 64 * 64
        total = self._off_flags + self.flag_elems
        backend = backend or os.environ.get("VPR_COMM_BACKEND", "auto")
        self.backend = None
        self.mc_ptr = 0
        errors = []

        def agreed(flag: bool) -> bool:             # every rank must take the same path (the set-up calls are collective)
            t = torch.tensor([1 if flag else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(t.item())

        if backend in ("auto", "symm_mem"):
            try:
                self._setup_symm_mem(total)
                ok = True
            except Exception as e:                                  # noqa: BLE001
                errors.append(f"symm_mem: {type(e).__name__}: {e}")
                ok = False
            if agreed(ok):
                self.backend = "symm_mem"
            elif backend == "symm_mem":
                raise RuntimeError("PeerComm: symmetric memory set-up failed (" + "; ".join(errors) + ")")
        if self.backend is None:
            try:
                self._setup_ipc(total)
                ok = True
            except Exception as e:                                  # noqa: BLE001
                errors.append(f"ipc: {type(e).__name__}: {e}")
                ok = False
            if not agreed(ok):
                raise RuntimeError("PeerComm: no peer-memory mapping available on every rank (" + "; ".join(errors) + ")")
            self.backend = "ipc"
            self.mc_ptr = 0
        if not agreed(bool(self.mc_ptr)) or os.environ.get("VPR_COMM_MULTIMEM", "1") == "0":
            self.mc_ptr = 0
        self.setup_errors = errors
        self.data = self._local[: self.numel]
        self._local.zero_()                         # flags must start at zero; the rest is for tidy padding
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        peers = L.CommPeers()
        peers.rank, peers.world = self.rank, self.world
        for r in range(self.world):
            base = self._peer_ptrs[r]
            peers.data[r] = base
            peers.staging[r] = base + 4 * self._off_staging
            peers.flags[r] = base + 4 * self._off_flags
        peers.mc_data = self.mc_ptr or None
        peers.staging_elems = self.staging_elems
        self._peers = peers

    # ---- plumbing -----------------------------------------------------------------------------------------------
    def _setup_symm_mem(self, total: int) -> None:
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        t = symm_mem.empty(total, dtype=torch.float32, device=self.device)
        try:
            hdl = symm_mem.rendezvous(t, group)
        except TypeError:
            hdl = symm_mem.rendezvous(t, group.group_name)
        off = int(getattr(hdl, "offset", 0) or 0)       # the tensor's byte offset inside the symmetric allocation
        ptrs = [int(p) + off for p in hdl.buffer_ptrs]
        if ptrs[self.rank] != t.data_ptr():
            raise RuntimeError("symmetric memory handle does not describe the local buffer")
        self._local, self._hdl, self._peer_ptrs = t, hdl, ptrs
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        self.mc_ptr = mc + off if mc else 0

    def _setup_ipc(self, total: int) -> None:
        t = torch.zeros(total, dtype=torch.float32, device=self.device)
        st = t.untyped_storage()
        handle = st._share_cuda_()                  # (device, ipc handle, size, offset, ref counter ..., event ...)
        infos = [None] * self.world
        dist.all_gather_object(infos, (handle, t.storage_offset() * 4), group=self.group)
        self._local = t
        self._peer_storages, self._peer_views, ptrs = [], [], []
        for r, (h, off) in enumerate(infos):
            if r == self.rank:
                ptrs.append(t.data_ptr())
                continue
            h = list(h)
            h[0] = self.device.index                # map the peer's allocation into THIS process's device context
            ps = torch.UntypedStorage._new_shared_cuda(*h)
            view = torch.empty(0, dtype=torch.float32, device=self.device).set_(ps, 0, (ps.nbytes() // 4,))
            self._peer_storages.append(ps)
            self._peer_views.append(view)
            ptrs.append(view.data_ptr() + off)
        self._peer_ptrs = ptrs

    def use_multimem(self, enabled: bool) -> None:
        """Switch between the switch-side reduction (multimem, needs the multicast mapping) and the peer-store path.
        Collective in effect: every rank must choose the same."""
        self._peers.mc_data = (self.mc_ptr or None) if enabled else None

    # ---- the collective -----------------------------------------------------------------------------------------
    def allreduce_(self, offset: int = 0, count: int | None = None, ctas: int = 16) -> None:
        """In-place sum over ranks of ``data[offset : offset + count]`` on the current stream (no host sync; safe to
        capture in a CUDA graph).  ``offset`` and ``count`` must be multiples of 4; every rank must make the same calls
        in the same order."""
        count = self.numel - offset if count is None else int(count)
        with torch.cuda.device(self.device):
            L.check(L.lib().vpr_allreduce_sum_f32(C.byref(self._peers), int(offset), count, int(ctas),
                                                  L.stream_ptr(self.device)), "vpr_allreduce_sum_f32")
