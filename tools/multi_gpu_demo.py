#!/usr/bin/env python
"""BASELINE config 4 in miniature, one process per GPU under torchrun: shard utterances i % world,
extract on each GPU, ONE NCCL all-gather of the (n/world, 512) embedding shards, every rank scores
its own row block of the all-pairs cosine matrix.  Rank 0 checks the gathered table and its score
block against a single-GPU run of the whole set."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200 import ops  # noqa: E402
from asv_subtools_b200.model.xvector import Xvector  # noqa: E402
from asv_subtools_b200.parallel import all_gather_embeddings, shard_indices  # noqa: E402
from oracle import nnet as onn  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, T, F = 1000, 200, 80
    m = Xvector(F, 10, training=False)
    m.load_state_dict(onn.make_state_dict(onn.xvector_spec(F), 102), strict=True)
    m.cuda().eval()
    feats = onn.synthetic_feats(n, T, F, 7)                 # every rank builds the same synthetic set
    idx = shard_indices(n, rank, world)
    local_emb = m.extract_embedding_batch(feats[idx])
    full = all_gather_embeddings(local_emb, n, rank, world)
    full = ops.center_length_norm(full, ops.column_mean(full))
    rows = full[idx].contiguous()
    block = ops.cosine_matrix(rows, full)                   # (n/world, n) row block of the all-pairs matrix
    # all-pairs EER without a score matrix: every rank bins its 256-row units, one all-reduce of counters
    from asv_subtools_b200.score import trial_histogram as th
    spk = torch.arange(n, device="cuda", dtype=torch.int32) % 50
    res = th.zoom_eer(full, spk, passes=3, rank=rank, world=world)
    ok = True
    if rank == 0:
        one = th.zoom_eer(full, spk, passes=3, group=False)
        ok = ok and one["eer"] == res["eer"] and int(res["hist"].sum()) == n * (n - 1) // 2
        print("multi_gpu_demo: all-pairs EER sharded {:.6f} == single {:.6f}: {}".format(res["eer"], one["eer"], ok))
    if rank == 0:
        ref = m.extract_embedding_batch(feats)
        ref = ops.center_length_norm(ref, ops.column_mean(ref))
        ok = ok and bool(torch.equal(ref, full)) and bool(torch.allclose(ops.cosine_matrix(ref, ref)[idx], block, atol=1e-6))
        print("multi_gpu_demo: world={} n={} gathered==single-GPU: {} block {}".format(world, n, ok, tuple(block.shape)))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
