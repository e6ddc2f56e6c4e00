"""Multi-GPU plumbing: utterances shard embarrassingly (one process per GPU, like the reference's
`nj` jobs, extract_xvectors_for_pytorch.sh:125-136); the only exchange on the path is the
all-gather of embeddings before all-pairs scoring (BASELINE config 4).  NCCL over NVLink on the GPU
box, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_indices(n, rank, world):
    """Utterance i goes to rank i % world (after the caller's length sort this balances frames)."""
    return list(range(rank, n, world))


def all_gather_embeddings(local, n_total, rank, world):
    """local: (n_local, D) rows for shard_indices(n_total, rank, world) -> (n_total, D) on every
    rank, in original utterance order.  Shards are padded to the common maximum so that one
    all_gather_into_tensor suffices."""
    d = local.shape[1]
    n_max = (n_total + world - 1) // world
    pad = torch.zeros(n_max, d, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    gathered = torch.empty(world * n_max, d, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad)
    # row r*n_max + j holds utterance j*world + r
    out = gathered.view(world, n_max, d).transpose(0, 1).reshape(world * n_max, d)
    return out[:n_total].contiguous()


def all_gather_blocks(local, out=None):
    """Contiguous sharding (rank r owns rows [r*n, (r+1)*n), every rank the same n): ONE all_gather_into_tensor
    straight into the (world*n, D) table, no padding, no reordering pass, no second buffer -- the layout the
    1 M-utterance job uses (bench.py, BASELINE configs[3]: 2.05 GB at 1 M x 512)."""
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out
