"""Pair-list sharding and result gather for the one-process-per-GPU deployment.

The reference processes pairs serially on one GPU (NCT/main.cu:471-540).  Pairs share nothing but the read-only VGG
weights, so the path shards at pair granularity with no data-path collective: rank r takes the lines i of pairs.txt
with i % world == r (the same rule nct_run_pairs applies in C++); the only communication is the gather of the result
images on rank 0 (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def shard_indices(n_pairs: int, rank: int, world: int):
    """indices of the pairs rank `rank` of `world` processes"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_pairs, world))


def gather_results(local_results, n_pairs: int, rank: int, world: int, dist=None):
    """local_results: list of equally shaped uint8 tensors, one per index of shard_indices(n_pairs, rank, world), on the
    device the process group communicates on.  Returns on rank 0 the list of all n_pairs results in pair order (None on
    the other ranks).  Ranks with fewer pairs pad with a zero tensor so every gather call has the same shape."""
    import torch

    if world == 1:
        return list(local_results)
    if dist is None:
        import torch.distributed as dist  # noqa: F811
    rounds = -(-n_pairs // world)
    proto = local_results[0] if local_results else None
    out = [None] * n_pairs if rank == 0 else None
    for k in range(rounds):
        if k < len(local_results):
            send = local_results[k]
        else:
            send = torch.zeros_like(proto) if proto is not None else None
        if send is None:
            raise ValueError("a rank without any pair cannot take part in the gather (n_pairs < world)")
        bucket = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
        dist.gather(send, bucket, dst=0)
        if rank == 0:
            for r in range(world):
                idx = k * world + r
                if idx < n_pairs:
                    out[idx] = bucket[r]
    return out
