"""Data-parallel plumbing of the path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The reference is single-device; the path shards naturally because every op is per-sample
(InstanceNorm statistics per (n,c), Grams per image, losses are batch means), so the only exchanges are
  C1  one all-reduce(avg) of the 1,712,771 TransformerNet gradients per step           (train_cnn.py:333-334)
  C2  one all-reduce(sum) of the per-rank 'smartaverage' feature/Gram sums per artist  (train_cnn.py:234-243)
"""
import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_range(n_items, rank, world_size):
    """Contiguous block of `n_items` owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class GradBucket:
    """Flat fp32 bucket: grads are packed, all-reduced ONCE (NVLS/NVLink under NCCL) and unpacked."""

    def __init__(self):
        self.flat = None

    def allreduce_mean(self, params, group=None):
        ws, _ = world(group)
        if ws == 1:
            return
        grads = [p.grad for p in params]
        sizes = [g.numel() for g in grads]
        if self.flat is None or self.flat.numel() != sum(sizes) or self.flat.device != grads[0].device:
            self.flat = torch.empty(sum(sizes), dtype=torch.float32, device=grads[0].device)
        chunks = list(self.flat.split(sizes))
        torch._foreach_copy_(chunks, [g.reshape(-1) for g in grads])
        if grads[0].is_cuda:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:                                   # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat /= ws
        # one multi-tensor copy back (a per-parameter loop was 62 tiny kernels per step)
        torch._foreach_copy_(grads, [c.view_as(g) for g, c in zip(grads, chunks)])


def allreduce_mean_flat(flat, group=None):
    """In-place mean all-reduce of ONE flat fp32 buffer (the TransformerNet gradient arena: no pack / unpack copies)."""
    ws, _ = world(group)
    if ws == 1:
        return
    if flat.is_cuda:
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:                                       # gloo has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= ws


def broadcast_parameters(params, group=None, src=0):
    """Make every rank start from rank `src`'s parameters (one flat broadcast).  Data-parallel training only averages
    gradients; without this a differing seed or a checkpoint loaded on one rank gives silently diverging replicas."""
    ws, _ = world(group)
    if ws == 1:
        return
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        dist.broadcast(flat, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
        off = 0
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n


def allreduce_sums(tensors, group=None):
    """In-place SUM all-reduce of a list of tensors (smartaverage feature / Gram sums and the painting count)."""
    ws, _ = world(group)
    if ws == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
