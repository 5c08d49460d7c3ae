"""Multi-GPU plumbing of the batched estimator: streams are independent, so rank r of `world` owns the
contiguous stream ids [r*batch, (r+1)*batch) and there is NO data-path collective.  The only exchange is the gather of the packed
window states [batch][W+1][16] (P3, Q4 xyzw, V3, Ba3, Bg3) after a solve -- one all-gather of batch*(W+1)*128 bytes per rank."""
import torch
import torch.distributed as dist


def stream_ids_for_rank(rank: int, world: int, batch_per_rank: int):
    return list(range(rank * batch_per_rank, (rank + 1) * batch_per_rank))


def rank_of_stream(stream_id: int, batch_per_rank: int) -> int:
    return stream_id // batch_per_rank


def gather_states(local: torch.Tensor, group=None) -> torch.Tensor:
    """local: (batch, W+1, 16) float64 on this rank's device (cuda with the nccl backend, cpu with gloo).
    Returns (world*batch, W+1, 16) ordered by global stream id on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local.clone()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out
