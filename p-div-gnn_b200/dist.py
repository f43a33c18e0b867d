"""Data-parallel training over independent mesh graphs (one process per GPU).

The reference is single-device.  The path shards naturally over graphs:
every rank runs the whole fwd+bwd on its own batch of meshes (graph-LayerNorm statistics
are per rank = the reference run with batch B_local) and the only exchange is ONE
all-reduce of the flat 167 299-float gradient buffer ``pdg_backward`` writes -- it is
averaged in place inside the autograd backward, before torch sees the per-parameter views.

The exchange itself is ``pdg_allreduce_mean`` (csrc/pdg_peer.cu): one kernel over NVLink peer memory (every rank
pushes its 669 KB buffer into a slot of each peer's IPC-mapped exchange area, raises a flag there and sums the slots it
holds in rank order).
``torch.distributed`` (NCCL) is the plumbing -- rendezvous, exchange of the IPC handles, parameter broadcast -- and the
fallback when peer mapping is not possible (more than 16 ranks, several nodes, ``PDG_P2P_ALLREDUCE=0``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun-style initialisation; returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_indices(n_items: int, rank: int, world: int):
    """Round-robin graph ownership: item i belongs to rank i % world."""
    return list(range(rank, n_items, world))


def allreduce_flat_(flat: torch.Tensor, group=None, peer=None) -> torch.Tensor:
    """In-place mean over ranks of a flat gradient buffer.

    ``peer``: a :class:`PeerAllreduce` (one kernel over NVLink peer memory); without it:

    NCCL: ONE collective with ``ReduceOp.AVG`` (the 1/G scale happens inside the reduction kernel -- no separate
    elementwise launch between the backward and the optimizer).  Other backends (gloo in the CPU tests): sum, then
    1/G (exact for G = 2^k)."""
    if peer is not None:
        return peer(flat)
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            if flat.is_cuda and dist.get_backend(group) == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
                flat.mul_(1.0 / world)
    return flat


class PeerAllreduce:
    """In-place mean of a flat CUDA fp32 buffer over the ranks of ``group`` through NVLink peer memory.

    Construction is collective: every rank allocates its exchange area in the library, the 64-byte IPC handles travel
    through ``all_gather_object`` and each rank maps the others'.  ``self_test`` runs one exchange and compares it with
    NCCL's result; :func:`enable_data_parallel` keeps NCCL when construction or the test fails."""

    def __init__(self, n_floats: int, device, group=None):
        from . import _lib
        L = _lib.lib()
        self.group, self.device, self.n = group, torch.device(device), int(n_floats)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise RuntimeError("PeerAllreduce handles up to 16 ranks of one node")
        self.seq = 0
        self._mine = C.c_void_p()
        self._opened = []
        # every rank runs the SAME sequence of collectives whatever fails locally, and all ranks raise together
        handle, err = (C.c_char * 64)(), None
        try:
            with torch.cuda.device(self.device):
                _lib.check(L.pdg_peer_alloc(self.n, self.world, C.byref(self._mine), handle), "pdg_peer_alloc")
        except Exception as ex:
            err = repr(ex)[:200]
        every = [None] * self.world
        dist.all_gather_object(every, (None if err else bytes(handle), os.uname().nodename), group=group)
        if err is None and any(h[0] is None for h in every):
            err = "a peer could not allocate its exchange area"
        if err is None and any(h[1] != every[0][1] for h in every):
            err = "ranks on several nodes (CUDA IPC is node-local)"
        self._ptrs = (C.c_void_p * self.world)()
        if err is None:
            try:
                with torch.cuda.device(self.device):
                    for r, (h, _) in enumerate(every):
                        if r == self.rank:
                            self._ptrs[r] = self._mine.value
                        else:
                            p = C.c_void_p()
                            _lib.check(L.pdg_peer_open((C.c_char * 64).from_buffer_copy(h), C.byref(p)), "pdg_peer_open")
                            self._ptrs[r] = p.value
                            self._opened.append(p.value)
            except Exception as ex:
                err = repr(ex)[:200]
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        ok = torch.tensor([0 if err else 1], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also the barrier: every area is mapped everywhere
        if not bool(ok.item()):
            self.close()
            raise RuntimeError(f"PeerAllreduce unavailable: {err or 'a peer failed to map the exchange areas'}")

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        from . import _lib
        if not flat.is_cuda or flat.dtype != torch.float32 or not flat.is_contiguous() or flat.numel() != self.n:
            raise ValueError("PeerAllreduce: expected a contiguous CUDA float32 buffer of the size it was built for")
        self.seq += 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pdg_allreduce_mean(_lib.ptr(flat), self.n, self._ptrs, self.rank, self.world, self.seq,
                                                     _lib.ptr(self.status), _lib.stream_ptr(self.device)), "pdg_allreduce_mean")
        return flat

    def timed_out(self) -> bool:
        """True when a peer did not arrive in some exchange (device -> host read)."""
        return bool(self.status.item())

    def self_test(self) -> bool:
        g = torch.Generator(device="cpu").manual_seed(1234 + self.rank)
        x = torch.randn(self.n, generator=g).to(self.device)
        ref = x.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM, group=self.group)
        ref /= self.world
        y = self(x.clone())
        ok = (not self.timed_out()) and torch.allclose(y, ref, rtol=1e-5, atol=1e-6)
        flag = torch.tensor([1 if ok else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)  # all ranks take the same decision
        return bool(flag.item())

    def close(self):
        from . import _lib
        L = _lib.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                L.pdg_peer_close(C.c_void_p(p))
            self._opened = []
            if self._mine:
                L.pdg_peer_free(self._mine)
                self._mine = C.c_void_p()


def enable_data_parallel(model, group=None, peer: bool | None = None):
    """Average ``pdg_backward``'s flat gradient across ranks inside backward.

    ``peer`` (default: env ``PDG_P2P_ALLREDUCE``, on): use the one-kernel NVLink exchange when the process group is NCCL
    on one node and its self-test agrees with NCCL; otherwise (and on any failure) the NCCL all-reduce stays."""
    model._pdg_dp = (True, group)
    model._pdg_peer = None
    if peer is None:
        peer = os.environ.get("PDG_P2P_ALLREDUCE", "1") not in ("", "0")
    if (peer and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
            and dist.get_backend(group) == "nccl"):
        from . import _lib
        dev = next(model.parameters()).device
        pa, ok = None, 0
        try:
            pa = PeerAllreduce(_lib.PDG_PARAM_ELEMS, dev, group)
            ok = 1
        except Exception as ex:  # no IPC / no peer access: keep NCCL (decided collectively below)
            model._pdg_peer_error = repr(ex)[:300]
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if bool(flag.item()) and pa.self_test():
            model._pdg_peer = pa
        elif pa is not None:
            try:
                pa.close()
            except Exception:
                pass
    return model


def broadcast_parameters(model, src: int = 0, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=group)
    return model
