"""Data-parallel plumbing for the training step (training/train_encoder.py:105-118,185,284-311).

The reference wraps the model in ``DistributedDataParallel`` and — because its accumulation loop never enters
``no_sync()`` — all-reduces every gradient after every micro-batch. Here gradients live in ONE flat
bf16 buffer (every ``param.grad`` is a view into it, so the wgrad GEMM epilogues accumulate in place) that is split
into a few buckets in backward order; each bucket is summed across ranks once per optimizer step, on a side stream,
as soon as the last micro-batch's backward has produced it (NCCL over NVLink/NVSwitch). The 1/world_size of DDP's
mean is folded into the fused optimizer (``grad_scale``).  Everything here is device-agnostic so the host logic is
covered by 2-process gloo tests on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_batch(global_batch: int, world_size: int, rank: int) -> tuple[int, int]:
    """[start, end) rows of the global batch owned by ``rank`` (train_encoder.py:115-118: equal split)."""
    assert global_batch % world_size == 0, "Batch size must be divisible by the number of processes."
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


class FlatGradBuckets:
    """Flat gradient storage + bucketed all-reduce.

    ``bucket_param_groups``: list of lists of parameters, in the order their gradients become final during backward
    (head first, embedding last). Every parameter's ``.grad`` becomes a view into the flat buffer.
    """

    def __init__(self, bucket_param_groups, process_group=None, comm_stream=None):
        params = [p for g in bucket_param_groups for p in g]
        assert params, "no parameters"
        dev, dtype = params[0].device, params[0].dtype
        align = 64  # elements: keeps every view 128-byte aligned for vector / TMA access
        offsets, total = [], 0
        self.bucket_ranges = []
        for group in bucket_param_groups:
            start = total
            for p in group:
                offsets.append(total)
                total += (p.numel() + align - 1) // align * align
            self.bucket_ranges.append((start, total))
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        for p, off in zip(params, offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        self.params = params
        self.param_bucket = {}
        for bi, group in enumerate(bucket_param_groups):
            for p in group:
                self.param_bucket[id(p)] = bi
        self.bucket_sizes = [len(g) for g in bucket_param_groups]
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.comm_stream = comm_stream
        self._pending = []
        self._ready_count = [0] * len(bucket_param_groups)
        self._armed = False
        # CUDA events around the compute stream's wait for the side stream in finish(): the part of the all-reduce that
        # the backward did not hide (read with exposed_ms(); recording costs nothing on the device)
        self._wait_events = None

    # ---- overlap protocol: arm() before the last micro-batch's backward, notify() as gradients become final -----
    def arm(self):
        self._armed = self.world > 1
        self._ready_count = [0] * len(self.bucket_sizes)

    def notify(self, params):
        """Called by the autograd functions right after they finished writing the gradients of ``params``."""
        if not self._armed:
            return
        for p in params:
            bi = self.param_bucket.get(id(p))
            if bi is None:
                continue
            self._ready_count[bi] += 1
            if self._ready_count[bi] == self.bucket_sizes[bi]:
                self._launch(bi)

    def _launch(self, bi):
        s, e = self.bucket_ranges[bi]
        chunk = self.flat[s:e]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                work = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        else:
            work = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self._pending.append(work)

    def finish(self):
        """Reduce whatever was not launched during backward and make the compute stream wait for all of it."""
        if self.world > 1:
            if self._armed:
                for bi, n in enumerate(self._ready_count):
                    if n != self.bucket_sizes[bi]:
                        self._launch(bi)
            else:
                for bi in range(len(self.bucket_sizes)):
                    self._launch(bi)
            timed = self.comm_stream is not None and self.flat.is_cuda
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            for w in self._pending:
                w.wait()
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            if timed:
                e1.record()
                self._wait_events = (e0, e1)
        self._pending = []
        self._armed = False

    def exposed_ms(self) -> float:
        """Milliseconds the compute stream spent waiting for the gradient all-reduce at the end of the latest step
        (synchronises; 0 without a process group)."""
        if self._wait_events is None:
            return 0.0
        e0, e1 = self._wait_events
        e1.synchronize()
        return float(e0.elapsed_time(e1))

    def zero_(self):
        self.flat.zero_()


def model_buckets(model, blocks_per_bucket: int = 1):
    """Backward-order buckets for OmniBioTA: [ln_f + lm_head], blocks from last to first, [wte].
    One block per bucket (25 MB for the small shape, 101 MB for the large one): the all-reduce of everything but the
    embedding is in flight long before the backward ends, and the last block's bucket — the one whose transfer can only
    start when the backward is nearly over — stays small."""
    buckets = [[model.transformer.ln_f.weight, model.lm_head.weight]]
    blocks = list(model.transformer.h)[::-1]
    for i in range(0, len(blocks), blocks_per_bucket):
        buckets.append([p for blk in blocks[i:i + blocks_per_bucket] for p in blk.parameters()])
    buckets.append([model.transformer.wte.weight])
    return buckets
