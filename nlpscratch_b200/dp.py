"""Data-parallel training over one NVLink/NVSwitch box: one process per GPU, batch sharded by rows.

The reference has no distributed code; what must hold is the identity checked on the
oracle: a step on R equal shards with gradients averaged over ranks equals the
reference's single step on the concatenated batch, because loss and dlogits are means over rows
(transformer.py:542,547).  Each rank therefore scales its dlogits by n_local/n_global (1/R for
equal shards), the flat gradient buffer is summed with NCCL, and clip + Adam then run identically
on every rank.

Overlap: the C library records one CUDA event per gradient bucket ([vocab] [layer L-1] ... [layer 0]
[We], contiguous ranges of the flat buffer) as backward completes it, and -- default backend ``"native"`` --
owns the exchange itself: ``nlps_comm_init`` creates an NCCL communicator inside ``libnlps_b200.so``, and
the undeferred fused step then sums each bucket group with ``ncclAllReduce`` on a communication stream
while earlier layers are still in backward, followed by clip + Adam -- ONE library call, ONE CUDA graph
after its second sighting.  ``torch.distributed`` is plumbing only: rendezvous, shipping the 128-byte
NCCL id, barriers.  ``backend="torch"`` keeps the round-1 path (``dist.all_reduce`` per bucket from
Python on a torch side stream, clip + Adam eager behind it) for A/B runs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of this rank's rows; the first ``n_rows % world`` ranks take one extra row."""
    base, extra = divmod(int(n_rows), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_weight(n_rows: int, rank: int, world: int, seq_len: int = 1) -> float:
    """Factor for this rank's dlogits so that SUM over ranks of local mean-gradients equals the
    gradient of the mean over the global batch (ragged last batch: utils/train.py:104-110)."""
    b, e = shard_rows(n_rows, rank, world)
    return float((e - b) * seq_len) / float(n_rows * seq_len) if n_rows else 0.0


def bucket_plan(sizes: Sequence[int], min_bucket: int) -> List[Tuple[int, int]]:
    """Coalesce consecutive (first, count) gradient ranges given as sizes in completion order into
    launches of at least ``min_bucket`` floats -- NVSwitch collectives are latency-, not link-bound,
    so tiny buckets (biases, LayerNorm) ride with their layer."""
    plan, start, acc = [], 0, 0
    for i, n in enumerate(sizes):
        acc += int(n)
        if acc >= min_bucket or i == len(sizes) - 1:
            plan.append((start, i + 1))
            start, acc = i + 1, 0
    return plan


class _CudaBuffer:
    """Expose a raw device range to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, n_floats: int, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def allreduce_sum_host(arrays: Sequence[np.ndarray], group=None):
    """CPU (gloo) twin of the bucketed all-reduce, used by the world_size-2 tests: sums each
    float32 array over ranks in place."""
    import torch
    import torch.distributed as dist
    works = []
    tensors = [torch.from_numpy(a) for a in arrays]
    for t in tensors:
        works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()


class DataParallel:
    """Wraps a ``nlpscratch_b200.Transformer`` for one rank of a data-parallel job."""

    def __init__(self, model, group=None, min_bucket_floats: int = 1 << 20, overlap: bool = True,
                 backend: Optional[str] = None):
        import ctypes as C
        import os
        import torch
        import torch.distributed as dist
        from . import _lib
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before DataParallel")
        self.model, self.group, self.overlap = model, group, overlap
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._torch, self._dist, self._lib, self._C = torch, dist, _lib, C
        self.backend = backend or os.environ.get("NLPS_DP_BACKEND", "native")
        if "NLPS_DP_BUCKET" in os.environ:                       # A/B: floats per coalesced exchange
            min_bucket_floats = int(os.environ["NLPS_DP_BUCKET"])
        if self.backend not in ("native", "torch"):
            raise ValueError("backend must be 'native' or 'torch'")
        if self.backend == "native":
            # rank 0 makes the NCCL id, torch.distributed's store ships its 128 bytes, every rank joins
            ident = [None]
            if self.rank == 0:
                buf = (C.c_char * 128)()
                _lib.call("nlps_comm_unique_id", buf)
                ident[0] = bytes(buf.raw)
            dist.broadcast_object_list(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            _lib.call("nlps_comm_init", model._h, C.c_char_p(ident[0]), C.c_int(self.rank), C.c_int(self.world),
                      C.c_int64(int(min_bucket_floats)))
        n = C.c_int()
        _lib.call("nlps_num_buckets", model._h, C.byref(n))
        self.ranges = []
        for b in range(n.value):
            first, cnt = C.c_int64(), C.c_int64()
            _lib.call("nlps_bucket_range", model._h, C.c_int(b), C.byref(first), C.byref(cnt))
            self.ranges.append((first.value, cnt.value))
        self.plan = bucket_plan([c for _, c in self.ranges], min_bucket_floats)
        flat = model.flat("grad")
        self._grad_t = torch.as_tensor(_CudaBuffer(flat.ptr, flat.size, model), device="cuda")
        pflat = model.flat("param")
        self._param_t = torch.as_tensor(_CudaBuffer(pflat.ptr, pflat.size, model), device="cuda")
        self.comm_stream = torch.cuda.Stream()
        self.launches = 0

    def broadcast_parameters(self, src: int = 0):
        """Identical replicas: rank ``src``'s weights everywhere (the reference has one process)."""
        if self.backend == "native":
            self._lib.call("nlps_comm_broadcast_params", self.model._h, self._C.c_int(int(src)))
            self.model.synchronize()
            return
        self.model.synchronize()
        self._dist.broadcast(self._param_t, src=src, group=self.group)
        self._torch.cuda.current_stream().synchronize()

    def allreduce_gradients(self):
        """Sum the flat gradient buffer over ranks (after an unfused backward or a deferred step): bucket by bucket as
        backward finishes them on the torch backend, one library call on the native one."""
        if self.backend == "native":
            self._lib.call("nlps_comm_allreduce_grads", self.model._h)
            return
        torch, dist, C = self._torch, self._dist, self._C
        m = self.model
        cs = self.comm_stream
        with torch.cuda.stream(cs):
            for lo, hi in self.plan:
                # the last bucket of the group is the last one to complete
                self._lib.call("nlps_stream_wait_bucket", m._h, C.c_int(hi - 1), C.c_void_p(cs.cuda_stream))
                # ranges of one group are adjacent only within a layer block; reduce each contiguous run
                for first, cnt in _merge_adjacent(self.ranges[lo:hi]):
                    dist.all_reduce(self._grad_t[first:first + cnt], op=dist.ReduceOp.SUM, group=self.group)
                    self.launches += 1
        # the model's stream continues (clip + Adam) only after the reductions
        self._lib.call("nlps_wait_stream", m._h, C.c_void_p(cs.cuda_stream))

    def train_step(self, x_local, y_local, learning_rate=1e-4, clip_grad_norm: Optional[float] = 1.0,
                   weight: Optional[float] = None):
        """One data-parallel step; returns this rank's local mean loss (device scalar)."""
        w = (1.0 / self.world) if weight is None else float(weight)
        if self.backend == "native":       # forward, loss, backward || bucketed ncclAllReduce, clip, Adam: one call, one graph
            return self.model.train_step(x_local, y_local, learning_rate, clip_grad_norm, grad_scale=w)
        loss = self.model.train_step(x_local, y_local, learning_rate, clip_grad_norm, grad_scale=w,
                                     defer_update=True)
        if self.world > 1:
            self.allreduce_gradients()
        self.model.apply_update(learning_rate, clip_grad_norm)
        return loss


    def close(self):
        if self.backend == "native" and self.model.__dict__.get("_h") is not None:
            self._lib.call("nlps_comm_destroy", self.model._h)


def save_checkpoint_rank0(dp: "DataParallel", filename: str):
    """Replicas are identical after every step, so rank 0 alone writes the resume checkpoint; everyone waits for it."""
    if dp.rank == 0:
        dp.model.save_checkpoint(filename)
    dp._dist.barrier(group=dp.group)


def _merge_adjacent(ranges):
    out = []
    for first, cnt in sorted(ranges):
        if out and out[-1][0] + out[-1][1] == first:
            out[-1] = (out[-1][0], out[-1][1] + cnt)
        else:
            out.append((first, cnt))
    return out
