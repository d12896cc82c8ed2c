"""Batch-sharded data parallelism (new; the reference is single-device): one process per GPU,
full replicas of G and D, per-rank batches, and ONE sum all-reduce of each network's flat fp32 gradient buffer
per optimizer step (NCCL over NVLink/NVSwitch on B200; gloo for the CPU tests).  The 1/world scale is folded
into the fused Adam kernel.  BatchNorm statistics in G stay per rank (reference semantics at per-GPU batch)."""
import os

import torch
import torch.distributed as dist


class Comm:
    def __init__(self, backend=None):
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if self.world_size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if self.backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group("nccl", rank=self.rank, world_size=self.world_size,
                                        device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(self.backend, rank=self.rank, world_size=self.world_size)

    def all_reduce_sum_(self, flat):
        if self.world_size > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return flat

    def all_reduce_max_(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def broadcast_module(self, module):
        """Rank 0's parameters and buffers become everyone's (identical replicas at step 0)."""
        if self.world_size > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0)

    def barrier(self):
        if self.world_size > 1:
            dist.barrier()

    def shard(self, n_global):
        """[lo, hi) slice of a global batch owned by this rank."""
        per = n_global // self.world_size
        return self.rank * per, (self.rank + 1) * per

    def close(self):
        if self.world_size > 1 and dist.is_initialized():
            dist.destroy_process_group()
