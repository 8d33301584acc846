"""Multi-GPU plumbing of the path: one process per GPU, whole games / positions shard with NO data-path
collective; the only exchange is the weight-blob broadcast at (re)load.

The reference has no collective at all: every GPU worker re-uploads ~40 tensors from host memory
(/root/reference/src/neural/cuda/cuda_common.cc:228-243, cuda_forward_pipe.cc:440-552).  Here the rank that
parsed the weight file packs them once and `torch.distributed.broadcast` (NCCL over NVLink/NVSwitch on GPUs, gloo
in the CPU tests) ships the packed blob to every replica.
"""
import torch


def shard_counts(n_units, world):
    """Units (games or positions) per rank: sizes differ by at most one, rank order = unit order."""
    base, extra = divmod(int(n_units), int(world))
    return [base + (1 if r < extra else 0) for r in range(world)]


def shard_range(n_units, rank, world):
    counts = shard_counts(n_units, world)
    start = sum(counts[:rank])
    return start, start + counts[rank]


def replicate_weights(pipe, dist, rank, device, src=0, gpu=0):
    """Broadcast rank `src`'s packed weight blob into every rank's replica and verify the copies.

    `pipe` needs weights_blob(gpu) -> (ptr, nbytes), weights_export(ptr, nbytes, gpu), weights_import(ptr, nbytes,
    gpu) and weights_checksum(gpu) (B200ForwardPipe, or a host-memory stand-in in the gloo tests).
    Returns the common checksum."""
    _, nbytes = pipe.weights_blob(gpu)
    sizes = torch.tensor([nbytes], dtype=torch.int64, device=device)
    lo, hi = sizes.clone(), sizes.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if int(lo) != int(hi):
        raise RuntimeError("weight blob sizes differ across ranks (%d..%d): different architectures?" % (int(lo), int(hi)))
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    if rank == src:
        pipe.weights_export(buf.data_ptr(), nbytes, gpu)
    if buf.is_cuda:
        torch.cuda.synchronize()
    dist.broadcast(buf, src=src)
    if buf.is_cuda:
        torch.cuda.synchronize()
    if rank != src:
        pipe.weights_import(buf.data_ptr(), nbytes, gpu)
    cs = torch.tensor([pipe.weights_checksum(gpu) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)
    lo, hi = cs.clone(), cs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if int(lo) != int(hi):
        raise RuntimeError("weight replicas differ after the broadcast")
    return int(lo)


def max_over_ranks(value, dist, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
