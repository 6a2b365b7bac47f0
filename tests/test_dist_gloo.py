"""N > 1 path on CPU: world_size-2 gloo process group, page sharding and the text gather that
bench.py runs over NCCL on the GPUs."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from ocrs_b200.dist import shard_pages, pack_texts, unpack_texts


def test_shard_pages_cover_everything():
    for n in (0, 1, 7, 8, 16, 17):
        for world in (1, 2, 3, 8):
            got = [p for r in range(world) for p in shard_pages(n, world, r)]
            assert got == list(range(n))
            sizes = [len(shard_pages(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip():
    texts = ["line one\nline two", "", "café 100% [b]"]
    assert unpack_texts(pack_texts(texts)) == texts


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from ocrs_b200.dist import gather_texts
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pages = list(shard_pages(5, world, rank))
    texts = [f"page {p}\nrank {rank} ✓" * (p + 1) for p in pages]
    out = gather_texts(texts)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_texts_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = q.get(timeout=120)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert len(out) == 2
    flat = [t for r in out for t in r]
    assert flat == [f"page {p}\nrank {0 if p < 3 else 1} ✓" * (p + 1) for p in range(5)]
