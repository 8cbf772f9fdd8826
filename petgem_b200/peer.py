"""Peer-memory transport of the multi-GPU Krylov iteration (host side of csrc/pg_comm.cu).

The reference reaches its two per-iteration collectives through PETSc: the VecScatter inside
``MatMult`` and the all-reduce inside ``VecDot``/``VecNorm`` of ``ksp.solve`` (petgem/solver.py:584-590) on
the objects of ``createParallelMatrix``/``createParallelVector`` (petgem/parallel.py:150-203).  Here the
ranks of one NVSwitch box map each other's buffers once (CUDA IPC; the 64-byte handles travel through the
``torch.distributed`` group, whatever its backend) and from then on both collectives are plain kernels:
``pg_comm_push``/``pg_comm_wait``/``pg_comm_ack`` for the halo, ``pg_comm_allreduce`` for the scalars.
``torch.distributed`` is the plumbing of the set-up only.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib, ptr, stream_ptr


class _RawCuda:
    """Device memory owned by the library, exposed to torch through the CUDA array interface."""

    def __init__(self, address, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(address), False),
                                         "version": 3, "strides": None}


class SymmetricBuffer:
    """One allocation per rank, mapped by every rank: `local` (uint8 tensor over this rank's memory) and
    `addr[r]` = address of rank r's allocation in THIS process."""

    def __init__(self, comm, nbytes):
        L = lib()
        self.comm = comm
        self.nbytes = int(max(nbytes, 16))
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        check(L.pg_ipc_alloc(self.nbytes, C.byref(own), handle), "pg_ipc_alloc")
        self.own = own.value
        handles = comm.all_gather_object(bytes(handle))
        self.addr = []
        self._opened = []
        for r, h in enumerate(handles):
            if r == comm.rank:
                self.addr.append(self.own)
                continue
            p = C.c_void_p()
            hb = (C.c_ubyte * 64).from_buffer_copy(h)
            check(L.pg_ipc_open(hb, C.byref(p)), "pg_ipc_open (CUDA IPC between the GPUs of the box)")
            self.addr.append(p.value)
            self._opened.append(p.value)
        self._raw = _RawCuda(self.own, self.nbytes)
        self.local = torch.as_tensor(self._raw, device=comm.device)

    def view(self, dtype, count=None, offset_bytes=0):
        item = torch.empty((), dtype=dtype).element_size()
        count = (self.nbytes - offset_bytes) // item if count is None else count
        return self.local[offset_bytes: offset_bytes + count * item].view(dtype)

    def close(self):
        """Collective: every rank unmaps before anybody frees."""
        if self.own is None:
            return
        L = lib()
        torch.cuda.synchronize()
        for p in self._opened:
            L.pg_ipc_close(C.c_void_p(p))
        self._opened = []
        self.comm.barrier()
        self.local = None
        L.pg_ipc_free(C.c_void_p(self.own))
        self.own = None


class PeerComm:
    """The ranks of one box with each other's control blocks mapped."""

    def __init__(self, dist, group=None, device=None, timeout_s=20.0):
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        L = lib()
        if self.world > 16:
            raise ValueError("peer transport: at most 16 ranks (one NVSwitch box)")
        self.ctrl = SymmetricBuffer(self, L.pg_comm_ctrl_bytes())
        table = (C.c_void_p * self.world)(*[C.c_void_p(a) for a in self.ctrl.addr])
        h = C.c_void_p()
        check(L.pg_comm_create(self.rank, self.world, table, float(timeout_s), C.byref(h)), "pg_comm_create")
        self.handle = h
        self._next_channel = 0
        self.barrier()  # nobody raises a flag before every control block is mapped

    # -- set-up plumbing over torch.distributed --------------------------------------------------
    def all_gather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        torch.cuda.synchronize()
        self.all_gather_object(0)

    def channel(self):
        """A fresh halo channel (same sequence of calls on every rank -> same id everywhere)."""
        c = self._next_channel
        if c >= 64:
            raise RuntimeError("peer transport: out of halo channels (64 per communicator)")
        self._next_channel += 1
        return c

    # -- the collectives ----------------------------------------------------------------------------
    def allreduce(self, t: torch.Tensor):
        """In-place sum over ranks of a complex128 device tensor of <= 64 scalars (rank order, bit-identical
        on every rank)."""
        k = t.numel()
        check(lib().pg_comm_allreduce(self.handle, k, ptr(t), ptr(t), stream_ptr()), "pg_comm_allreduce")

    def status(self):
        check(lib().pg_comm_status(self.handle, stream_ptr()), "pg_comm_status")

    def close(self):
        if self.handle is not None:
            torch.cuda.synchronize()
            self.barrier()
            lib().pg_comm_destroy(self.handle)
            self.handle = None
            self.ctrl.close()


class PeerExchange:
    """A halo pattern on one channel: which of my entries go to which rank (send_idx grouped by destination)
    and which ranks send to me.  `target(...)` binds it to a destination buffer on every rank."""

    def __init__(self, comm: PeerComm, send_idx: torch.Tensor, send_splits, recv_splits):
        self.comm = comm
        self.chan = comm.channel()
        self.send_idx = send_idx.to(torch.int32).contiguous()
        self.send_splits = [int(s) for s in send_splits]
        self.recv_splits = [int(s) for s in recv_splits]
        seg = [0]
        for s in self.send_splits:
            seg.append(seg[-1] + s)
        self.seg = (C.c_int64 * (comm.world + 1))(*seg)
        self.from_mask = sum(1 << r for r, s in enumerate(self.recv_splits) if s > 0)
        self.n_recv = sum(self.recv_splits)
        # where my segment starts inside the receive area of each destination: after the segments of lower ranks
        allrecv = comm.all_gather_object(self.recv_splits)
        self.dst_entry = [sum(allrecv[d][: comm.rank]) for d in range(comm.world)]

    def target(self, buf: SymmetricBuffer, recv_offset_bytes_by_rank, k):
        """Destination table for pushes of k interleaved right-hand sides into `buf`, whose receive area starts
        recv_offset_bytes_by_rank[d] bytes into rank d's allocation."""
        w = self.comm.world
        dst = [(buf.addr[d] + int(recv_offset_bytes_by_rank[d]) + self.dst_entry[d] * k * 16) if self.send_splits[d]
               else 0 for d in range(w)]
        return (C.c_void_p * w)(*[C.c_void_p(a) for a in dst])

    def push(self, x: torch.Tensor, k, dst_table):
        check(lib().pg_comm_push(self.comm.handle, self.chan, int(k), ptr(x), ptr(self.send_idx), self.seg, dst_table,
                                 stream_ptr()), "pg_comm_push")

    def wait(self):
        check(lib().pg_comm_wait(self.comm.handle, self.chan, self.from_mask, stream_ptr()), "pg_comm_wait")

    def ack(self):
        check(lib().pg_comm_ack(self.comm.handle, self.chan, self.from_mask, stream_ptr()), "pg_comm_ack")
