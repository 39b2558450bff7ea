"""Peer-memory exchange for a sharded generation (csrc/des_comm.cu): one process per GPU on one node.

torch.distributed is used once, to hand every rank the 64-byte cudaIpc handles of the others; after that the two
exchange steps of a generation (fitness all-gather, partial all-reduce) are kernels of this library storing into the
peers' memory over NVLink — no NCCL call on the hot path, and the whole generation can be captured in a CUDA graph.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class _DevArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it (no copy, no ownership)."""

    def __init__(self, ptr, n, owner):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}
        self._owner = owner


class PeerComm:
    def __init__(self, N, P, device, process_group=None):
        self.lib = _lib.load()
        self.pg = process_group
        self.rank = dist.get_rank(process_group)
        self.world = dist.get_world_size(process_group)
        self.device = torch.device(device)
        self.N, self.P = int(N), int(P)
        self._h = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.des_comm_create(C.byref(self._h), self.rank, self.world, self.N, self.P, handle), 'des_comm_create')
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
            every = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(every, mine, group=process_group)
            blob = b''.join(bytes(t.cpu().numpy().tobytes()) for t in every)
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            _lib.check(self.lib.des_comm_connect(self._h, buf), 'des_comm_connect')
            ptr = self.lib.des_comm_fitness_all_dev(self._h)
            self.fitness_all = torch.as_tensor(_DevArray(ptr, self.N, self), device=self.device)
            dist.barrier(group=process_group)            # every block is mapped before anyone stores into a peer

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def allgather_fitness(self, member_offset, n_local):
        _lib.check(self.lib.des_comm_allgather_fitness(self._h, int(member_offset), int(n_local), self._stream()),
                   'des_comm_allgather_fitness')

    def allreduce_partial(self, partial_local, out):
        assert partial_local.is_cuda and out.is_cuda and partial_local.dtype == torch.float32 and out.dtype == torch.float32
        _lib.check(self.lib.des_comm_allreduce_partial(self._h, C.c_void_p(out.data_ptr()), C.c_void_p(partial_local.data_ptr()),
                                                       self.P, self._stream()), 'des_comm_allreduce_partial')

    def close(self):
        if self._h:
            self.fitness_all = None
            self.lib.des_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
