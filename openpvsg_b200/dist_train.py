"""Data-parallel training (SURVEY.md 8e for the training row): clips are sharded over the ranks, weights replicated, and
the one exchange step is the gradient all-reduce after backward -- what torch DistributedDataParallel does for the
reference (tools/train.py --launcher pytorch -> mmdet train_detector -> MMDistributedDataParallel).

One flat fp32 bucket (the detector has 44 M parameters = 176 MB: a single NCCL all-reduce over NVLink / NVSwitch, where
in-switch reduction applies, instead of per-tensor calls), averaged over the ranks, copied back into the .grad tensors.
Parameters without a gradient on this rank contribute zeros, so every rank issues the same collective.  gloo on CPU in
the tests (tests/test_tubes_cpu.py)."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_gradients(params, bucket=None):
    """Average the gradients of ``params`` over all ranks, in place.  ``bucket``: optional reusable flat buffer.
    Returns the flat buffer (so a caller can keep it between steps)."""
    params = [p for p in params if p.requires_grad]
    if world() == 1 or not params:
        return bucket
    n = sum(p.numel() for p in params)
    dev = params[0].device
    if bucket is None or bucket.numel() != n or bucket.device != dev:
        bucket = torch.empty(n, dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        k = p.numel()
        if p.grad is None:
            bucket[off:off + k].zero_()
        else:
            bucket[off:off + k].copy_(p.grad.reshape(-1))
        off += k
    dist.all_reduce(bucket)
    bucket.div_(world())
    off = 0
    for p in params:
        k = p.numel()
        if p.grad is None:
            p.grad = bucket[off:off + k].view_as(p).clone()
        else:
            p.grad.copy_(bucket[off:off + k].view_as(p))
        off += k
    return bucket


def broadcast_parameters(model, src=0):
    """Make every rank start from rank ``src``'s weights (DDP does this at construction)."""
    if world() == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src)
