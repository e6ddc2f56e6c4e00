#!/usr/bin/env python
"""torchrun check of the peer-stored embedding table (csrc/peer.cu, parallel.PeerTable) against NCCL's all-gather:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_table_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200.model.xvector import Xvector  # noqa: E402
from asv_subtools_b200.parallel import PeerTable, all_gather_blocks  # noqa: E402
from oracle import nnet as onn  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    m = Xvector(80, 10, training=False, extracted_embedding="far")
    m.load_state_dict(onn.make_state_dict(onn.xvector_spec(80), 102), strict=True)
    m.cuda().eval()
    ex = m.extractor()
    n, t = 1000, 61                                    # ragged last batch: 1000 = 3 x 256 + 232
    feats = torch.from_numpy(onn.synthetic_feats(n, t, 80, 4000 + rank)).to(dev)
    plain = ex.extract_shard(feats, 256).clone()
    want = all_gather_blocks(plain)
    table = PeerTable(n, 512)
    table.attach(ex)
    for rep in range(3):
        table.tensor.zero_()
        table.barrier()
        got_local = ex.extract_shard(feats, 256)
        table.barrier()
        ok = torch.equal(table.tensor, want) and torch.equal(got_local, plain)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("peer_table_check: rep %d world %d table %s == all-gather: %s" % (rep, world, tuple(table.tensor.shape), bool(flag.item())))
        assert flag.item() == 1
    host = torch.empty(n, t, 80, dtype=torch.float32, pin_memory=True)
    host.copy_(feats)
    out = torch.empty(n, 512, dtype=torch.float32, pin_memory=True)
    table.tensor.zero_()
    table.barrier()
    ex.extract_shard_host(host.data_ptr(), n, t, out.data_ptr(), 256)
    table.barrier()
    ok = torch.equal(table.tensor, want) and torch.equal(out, plain.cpu())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("peer_table_check: host-buffer shard call: %s" % bool(flag.item()))
    assert flag.item() == 1
    table.detach(ex)
    table.barrier()
    table.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
