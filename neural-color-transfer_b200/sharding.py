"""Pair-list sharding and result gather for the one-process-per-GPU deployment.

The reference processes pairs serially on one GPU (NCT/main.cu:471-540).  Pairs share nothing but the read-only VGG
weights, so the path shards at pair granularity with no data-path collective: rank r takes the lines i of pairs.txt
with i % world == r (the same rule nct_run_pairs applies in C++); the only communication is the gather of the result
images on rank 0 (NCCL over NVLink on the GPU box, gloo in the CPU tests).

NCCL has no gather primitive; the gather is point-to-point: rank 0 posts the receives of a
round up front as ONE group (`batch_isend_irecv` = ncclGroupStart/End, so they progress concurrently), every other rank
sends a pair's result as soon as that pair is finished, on the process group's own stream, overlapped with the next
pair's compute.  Result shapes differ from pair to pair (pairs.txt images), so the shapes travel first in one small
`all_gather_object`; ranks without any pair (n_pairs < world) take part in that collective and then have nothing to send.
"""
from __future__ import annotations


def shard_indices(n_pairs: int, rank: int, world: int):
    """indices of the pairs rank `rank` of `world` processes"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_pairs, world))


class StreamedGather:
    """Per-pair streamed gather of result tensors on rank 0.

    `shapes` = the shapes of this rank's results, in the order of shard_indices(n_pairs, rank, world) (known before the
    compute: a result has the content image's shape).  Protocol, identical on every rank:
        g = StreamedGather(shapes, n_pairs, rank, world, device, dtype, dist)   # collective: all_gather_object
        g.post()                      # rank 0: receives for every remote pair, one group per round
        g.send(k, tensor)             # k-th local result is ready (on the current stream): isend to rank 0 / keep
        out = g.wait()                # rank 0: list of n_pairs tensors in pair order; others: None
    """

    def __init__(self, shapes, n_pairs: int, rank: int, world: int, device, dtype, dist=None):
        import torch

        if dist is None and world > 1:
            import torch.distributed as dist  # noqa: F811
        self.torch, self.dist = torch, dist
        self.n_pairs, self.rank, self.world, self.device, self.dtype = n_pairs, rank, world, device, dtype
        self.mine = shard_indices(n_pairs, rank, world)
        shapes = [tuple(int(v) for v in s) for s in shapes]
        if len(shapes) != len(self.mine):
            raise ValueError(f"rank {rank} owns {len(self.mine)} pairs but gave {len(shapes)} shapes")
        if world > 1:
            every = [None] * world
            dist.all_gather_object(every, shapes)  # every rank enters, also those without pairs
        else:
            every = [shapes]
        self.all_shapes = every
        self.out = [None] * n_pairs if rank == 0 else None
        self.pending = []

    def post(self):
        if self.rank != 0 or self.world == 1:
            return
        torch, dist = self.torch, self.dist
        rounds = -(-self.n_pairs // self.world)
        for k in range(rounds):
            ops = []
            for r in range(1, self.world):
                idx = k * self.world + r
                if idx < self.n_pairs:
                    buf = torch.empty(self.all_shapes[r][k], dtype=self.dtype, device=self.device)
                    self.out[idx] = buf
                    ops.append(dist.P2POp(dist.irecv, buf, r))
            if ops:
                self.pending += dist.batch_isend_irecv(ops)

    def send(self, k: int, tensor):
        if tuple(tensor.shape) != self.all_shapes[self.rank][k]:
            raise ValueError(f"result {k} has shape {tuple(tensor.shape)}, announced {self.all_shapes[self.rank][k]}")
        if self.rank == 0:
            self.out[self.mine[k]] = tensor
        else:
            self.pending.append(self.dist.isend(tensor, 0))

    def wait(self):
        for w in self.pending:
            w.wait()
        self.pending = []
        return self.out


def gather_results(local_results, n_pairs: int, rank: int, world: int, dist=None):
    """local_results: list of uint8 tensors (any shapes), one per index of shard_indices(n_pairs, rank, world), on the
    device the process group communicates on.  Returns on rank 0 the list of all n_pairs results in pair order (None on
    the other ranks).  Safe for n_pairs < world (ranks without pairs only join the shape exchange)."""
    import torch

    if world == 1:
        return list(local_results)
    dev = local_results[0].device if local_results else torch.device("cpu")
    if local_results:
        dtype = local_results[0].dtype
    else:
        dtype = torch.uint8
    g = StreamedGather([t.shape for t in local_results], n_pairs, rank, world, dev, dtype, dist)
    g.post()
    for k, t in enumerate(local_results):
        g.send(k, t)
    return g.wait()
