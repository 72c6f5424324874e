"""Multi-GPU plumbing: independent sequences shard across ranks (one process per GPU); the only communication is
a final gather of per-frame poses and masks (the reference itself has no collective)."""
import torch
import torch.distributed as dist


def my_sequences(n_sequences, rank, world):
    """Sequence ids owned by ``rank``: seq_id % world == rank (config 4: 64 sequences over 1/2/4/8 GPUs)."""
    return [s for s in range(n_sequences) if s % world == rank]


def gather_results(odom, mask, group=None):
    """All ranks contribute odom f64 [..., 7] and mask u8 [..., N]; every rank returns the rank-major concatenation
    (NCCL all_gather over NVLink on GPUs, gloo on CPU).  ~8 KB per frame: latency-bound, done once per job."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return odom.unsqueeze(0), mask.unsqueeze(0)
    world = dist.get_world_size(group)
    od = [torch.empty_like(odom) for _ in range(world)]
    mk = [torch.empty_like(mask) for _ in range(world)]
    dist.all_gather(od, odom.contiguous(), group=group)
    dist.all_gather(mk, mask.contiguous(), group=group)
    return torch.stack(od), torch.stack(mk)
