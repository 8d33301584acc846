"""N>1 host logic on CPU: world_size-2 gloo processes run the weight-blob replication protocol of
sayuri_b200/dist.py against a host-memory stand-in for the engine, and the weak-scaling sharding helpers."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class HostBlobPipe:
    """Same four methods as B200ForwardPipe's blob interface, over a numpy buffer."""

    def __init__(self, nbytes, fill_seed=None):
        self.blob = np.zeros(nbytes, dtype=np.uint8)
        if fill_seed is not None:
            self.blob[:] = np.random.default_rng(fill_seed).integers(0, 256, nbytes, dtype=np.uint8)

    def weights_blob(self, gpu=0):
        return self.blob.ctypes.data, self.blob.nbytes

    def weights_export(self, ptr, nbytes, gpu=0):
        ctypes.memmove(ptr, self.blob.ctypes.data, nbytes)

    def weights_import(self, ptr, nbytes, gpu=0):
        ctypes.memmove(self.blob.ctypes.data, ptr, nbytes)

    def weights_checksum(self, gpu=0):
        h = 1469598103934665603
        for b in self.blob[:: max(1, self.blob.size // 4096)].tolist():   # sampled FNV-1a, enough for the test
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nbytes_by_rank, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sayuri_b200.dist import max_over_ranks, replicate_weights, shard_range
        pipe = HostBlobPipe(nbytes_by_rank[rank], fill_seed=123 if rank == 0 else None)
        try:
            cs = replicate_weights(pipe, dist, rank, torch.device("cpu"))
            want = HostBlobPipe(nbytes_by_rank[0], fill_seed=123)
            ok = bool(np.array_equal(pipe.blob, want.blob)) and cs == (want.weights_checksum() & 0x7FFFFFFFFFFFFFFF)
            results[rank] = ("ok" if ok else "mismatch", shard_range(513, rank, world), max_over_ranks(rank + 1.5, dist, "cpu"))
        except RuntimeError as e:
            results[rank] = ("error: %s" % e, None, None)
    finally:
        dist.destroy_process_group()


def _run(world, nbytes_by_rank):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), nbytes_by_rank, results), nprocs=world, join=True)
    return dict(results)


def test_blob_broadcast_world2_gloo():
    res = _run(2, [1 << 20, 1 << 20])
    assert res[0][0] == "ok" and res[1][0] == "ok"
    assert res[0][1] == (0, 257) and res[1][1] == (257, 513)       # weak-scaling shards cover the units exactly once
    assert res[0][2] == res[1][2] == 2.5                            # max over ranks (timing reduction)


def test_blob_size_mismatch_is_detected():
    res = _run(2, [4096, 8192])
    assert all(r[0].startswith("error") and "sizes differ" in r[0] for r in res.values())


def test_shard_counts():
    from sayuri_b200.dist import shard_counts, shard_range
    assert shard_counts(512, 8) == [64] * 8
    assert shard_counts(10, 4) == [3, 3, 2, 2]
    covered = []
    for r in range(4):
        a, b = shard_range(10, r, 4)
        covered += list(range(a, b))
    assert covered == list(range(10))
