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


class PeerTable:
    """The (world x n, D) embedding table of BASELINE configs[3] on every GPU of one node, filled by peer stores over
    NVLink instead of an all-gather after extraction (csrc/peer.cu): every rank cudaMallocs its copy, the 64-byte CUDA
    IPC handles travel through the process group, every rank maps the others' copies, and the shard calls then store
    each batch's embeddings into all copies while the next batches run (`attach(extractor)`).  `tensor` is this rank's
    copy as a torch view; it is complete on every rank after `barrier()`.

    Same-node ranks only (CUDA IPC + peer access).  Raises RuntimeError if the driver refuses the mapping -- the caller
    falls back to `all_gather_blocks`."""

    def __init__(self, rows_per_rank, dim, group=None, device=None, total_rows=None):
        import ctypes as C

        from ._lib import check, lib
        self._C, self._lib, self._check = C, lib, check
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise ValueError("PeerTable maps at most 16 peers (XVB_MAX_PEERS)")
        self.n, self.dim = int(rows_per_rank), int(dim)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rows = int(total_rows) if total_rows is not None else self.world * self.n
        nbytes = self.rows * self.dim * 4
        base = C.c_void_p()
        check(lib.xvb_ipc_alloc(C.byref(base), nbytes), "xvb_ipc_alloc")
        self._base = base.value
        self._opened = []
        handle = (C.c_uint8 * 64)()
        check(lib.xvb_ipc_export(C.c_void_p(self._base), handle), "xvb_ipc_export")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        everyone = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        handles = everyone.cpu().numpy().reshape(self.world, 64)
        self.pointers = (C.c_void_p * self.world)()
        ok = torch.ones(1, dtype=torch.int32, device=self.device)
        try:
            for r in range(self.world):
                if r == self.rank:
                    self.pointers[r] = self._base
                    continue
                p = C.c_void_p()
                buf = (C.c_uint8 * 64)(*handles[r].tolist())
                check(lib.xvb_ipc_open(buf, C.byref(p)), "xvb_ipc_open")
                self._opened.append(p.value)
                self.pointers[r] = p.value
        except RuntimeError:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)       # all or nobody: a half-mapped table would hang the step
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError("CUDA IPC peer mapping is not available between these ranks")
        self.tensor = self._view(self._base, (self.rows, self.dim))

    def _view(self, ptr, shape):
        holder = type("_CudaBuffer", (), {})()
        holder.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                           "strides": None}
        t = torch.as_tensor(holder, device=self.device)
        t._xvb_keepalive = self          # the allocation lives as long as the table object
        return t

    def attach(self, extractor, row0=None):
        """Make `extractor`'s shard calls (x-vector `ops.Extractor` or the ECAPA extractor) store into every copy, this
        rank's rows starting at `row0` (default rank * rows_per_rank; unequal shards pass their own offsets)."""
        extractor.set_gather(self.pointers, self.world, self.rank * self.n if row0 is None else int(row0), self.dim)

    @staticmethod
    def detach(extractor):
        extractor.set_gather(None, 0, 0, 0)

    def barrier(self):
        """All ranks have issued their stores and finished them: order this rank's stream, then meet the others."""
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)

    def close(self):
        C, lib = self._C, self._lib
        for p in self._opened:
            lib.xvb_ipc_close(C.c_void_p(p))
        self._opened = []
        if self._base:
            torch.cuda.synchronize()
            lib.xvb_ipc_free(C.c_void_p(self._base))
            self._base = None
