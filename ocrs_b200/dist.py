"""Multi-GPU plumbing: pages shard across ranks with no data-path collective; the only exchange is
the gather of the recognised text to rank 0.  `torch.distributed` is used for
the process group only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence


def shard_pages(n_pages: int, world: int, rank: int) -> range:
    """Contiguous page shards, sizes differing by at most one."""
    base, extra = divmod(n_pages, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def pack_texts(texts: Sequence[str]) -> bytes:
    """Per-page texts -> one byte string (form-feed separated; OCR output never contains \\f)."""
    return "\f".join(texts).encode("utf-8")


def unpack_texts(blob: bytes) -> List[str]:
    return blob.decode("utf-8").split("\f") if blob else []


def gather_texts(texts: Sequence[str], device=None) -> Optional[List[List[str]]]:
    """Gathers every rank's page texts on rank 0 (two phases: byte counts, then padded payload).
    Returns [rank][page] on rank 0 and None elsewhere.  Works with any initialised backend."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(texts)]
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = pack_texts(texts)
    n = torch.tensor([len(payload)], dtype=torch.int64, device=device)
    lens = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(lens, n)
    mx = max(1, max(int(x.item()) for x in lens))
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    if payload:
        buf[: len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return [unpack_texts(bytes(out[r][: int(lens[r].item())].cpu().numpy().tobytes())) for r in range(world)]
