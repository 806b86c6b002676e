"""One process per GPU: torch.distributed is the plumbing (rendezvous, barrier, max-over-ranks
timing, broadcasting the NCCL id); the DSS halo itself is ncclSend/ncclRecv issued by the library
on the compute stream (role of ClimaComms ``MPICommsContext`` + ``graph_context`` in the reference,
src/simulation/grids.jl:44,75; docs/src/gpu_and_mpi.md:82-93)."""
from __future__ import annotations

import ctypes as C
import os


class DistributedComms:
    def __init__(self, backend=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.nranks = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.nranks > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group(backend=backend, device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(backend=backend)
        elif torch.cuda.is_available():
            torch.cuda.set_device(self.local_rank)

    def nccl_unique_id(self) -> bytes:
        """Rank 0 asks the library (ncclGetUniqueId) and broadcasts the 128 bytes."""
        from . import capi

        obj = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            capi.check(capi.load().b200_nccl_unique_id(C.cast(buf, C.c_void_p)), "b200_nccl_unique_id")
            obj[0] = buf.raw
        if self.nranks > 1:
            self.dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def barrier(self):
        if self.nranks > 1:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.nranks == 1:
            return x
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        t = self.torch.tensor([x], dtype=self.torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, flag: bool) -> bool:
        if self.nranks == 1:
            return bool(flag)
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        t = self.torch.tensor([1 if flag else 0], dtype=self.torch.int32, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def finalize(self):
        if self.nranks > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()
