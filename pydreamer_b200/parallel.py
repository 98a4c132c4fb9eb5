"""Data parallelism over the batch dimension: one process per GPU, parameters replicated,
ONE collective per step — an all-reduce(SUM) of the flat fp32 gradient arena (all four optimizer groups,
~48.6 M elements at the Atari shape) over NCCL/NVLink, scaled by 1/world BEFORE grad_clip so clip thresholds and
the grad_norm* metrics equal the single-process global-batch values (dreamer.py:73-87).

The reference has no distributed path at all; every loss is a mean over the batch, so with equal
per-rank batches the global gradient is the mean of the rank gradients.  No other exchange is needed (LayerNorm is
per-row, there is no BatchNorm).  Works with any torch.distributed backend (tests use gloo on CPU)."""
import torch
import torch.distributed as dist


class GradAllReduce:
    def __init__(self, world_size=None, group=None):
        self.group = group
        self.world = world_size or dist.get_world_size(group)
        self._scale = None

    def broadcast_params(self, model, src=0):
        model._ensure_arena()
        dist.broadcast(model._arena, src, group=self.group)
        model._weights_dirty = True

    def allreduce_grads(self, model):
        g = model._garena
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        model.ops.scale_by(g, None, 1.0 / self.world)        # exact fp32 scaling (no operand rounding)
